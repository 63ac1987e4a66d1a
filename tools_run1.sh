mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 600 gpurun_out/b_default.err
timeout 300 python bench.py --workload satd > gpurun_out/b_satd.json 2> gpurun_out/b_satd.err
timeout 400 python bench.py --workload lookahead --weightp 1 > gpurun_out/b_la_w.json 2> gpurun_out/b_la_w.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/la_launches.csv python bench.py --workload lookahead --steps 1 --warmup 3 --quick > gpurun_out/ncu_la.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 6 -c 1 -f -o gpurun_out/search_full python bench.py --workload lookahead --steps 1 --warmup 3 --quick > gpurun_out/ncu_search.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/satd_launches.csv python bench.py --workload satd --steps 2 --warmup 3 --quick > gpurun_out/ncu_satd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mvfield_kernel -s 3 -c 1 -f -o gpurun_out/satd_full python bench.py --workload satd --steps 2 --warmup 3 --quick > gpurun_out/ncu_satd_full.log 2>&1
cat gpurun_out/b_default.json gpurun_out/b_satd.json gpurun_out/b_la_w.json gpurun_out/b_ref.json
