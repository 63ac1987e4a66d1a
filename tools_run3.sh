mkdir -p gpurun_out
L=$PWD/x264_b200/csrc
timeout 900 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_slicetype.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['ms_per_launch'], d['roofline']['ms_per_launch_28_searches'])"; done
echo "== phase profile"; X264CU_LIB=$L/libx264_b200_prof.so timeout 300 python tools/la_phase_profile.py 2>&1 | tail -9
