python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; grep -c "smoke\[" gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log | cut -c1-300
python bench.py --workload rgb_shared --no-comparators > gpurun_out/bench_rgbs.json 2> gpurun_out/bench_rgbs.err; tail -c 1500 gpurun_out/bench_rgbs.json
ncu --set full --clock-control none --import-source on -k regex:conv3x3_f16x2 -c 4 -o gpurun_out/r02_conv_f16x2 -f python tools/conv_bench.py --modes f16x2 --reps 1 > gpurun_out/ncu_f16x2.log 2>&1
ncu -i gpurun_out/r02_conv_f16x2.ncu-rep --page raw --csv > gpurun_out/r02_conv_f16x2_raw.csv 2>/dev/null; wc -c gpurun_out/r02_conv_f16x2_raw.csv
