mkdir -p gpurun_out
L=$PWD/x264_b200/csrc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 400 gpurun_out/b_default.err
python -c "import json; d=json.load(open('gpurun_out/b_default.json')); print('default', d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['ms_per_launch_28_searches'])"
echo "== stats base"; X264CU_STATS=1 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>&1 | grep -E "x264cu" | cut -c1-300
echo "== phase profile"; X264CU_LIB=$L/libx264_b200_prof.so timeout 300 python tools/la_phase_profile.py 2>&1 | tail -9
echo "== parity of variants"
for v in c3 r32; do X264CU_LIB=$L/libx264_b200_$v.so timeout 600 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_slicetype.py -x -q 2>&1 | tail -1; done
echo "== variants"
for v in "" _c3 _r32; do
 for cfg in "8 4" "12 6" "16 8"; do
  set -- $cfg
  X264CU_LIB=$L/libx264_b200$v.so X264CU_RUN_AHEAD=$1 X264CU_PREFETCH_GROUP=$2 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib$v ra/pg $cfg', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['ms_per_launch'], d['roofline']['ms_per_launch_28_searches'])"
 done
done
