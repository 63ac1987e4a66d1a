L=$PWD/x264_b200/csrc
timeout 600 python -m pytest tests/test_gpu_aq.py -x -q 2>&1 | tail -2
for v in "" _nw10 _nw12; do
 X264CU_LIB=$L/libx264_b200$v.so timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('lib$v', round(d['value'],1), round(d['e2e']['value'],1), round(r['ms_per_launch'],2), round(r['share_of_step'],2))"
done
for v in _nw10 _nw12; do X264CU_LIB=$L/libx264_b200$v.so timeout 600 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_slicetype.py -x -q 2>&1 | tail -1; done
