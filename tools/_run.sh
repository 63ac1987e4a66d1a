mkdir -p gpurun_out
L=$PWD/x264_b200/csrc
for v in "" _r40 _r24; do
 X264CU_LIB=$L/libx264_b200$v.so timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('lib$v', round(d['value'],1), round(d['e2e']['value'],1), round(r['ms_per_launch'],2), round(r['share_of_step'],2), r['traffic'], round(r['ms_per_launch_alone_28_searches'],2))"
done
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_mbtree.py tests/test_gpu_mc.py -x -q -k "not cfg6 and not 3840 and not 352" 2>&1 | tail -4
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_lookahead.py -x -q -k "cfg0 or cfg4" 2>&1 | tail -4
echo "== synccheck"; timeout 300 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_lookahead.py -x -q -k "cfg0" 2>&1 | tail -3
