python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; grep -c "lossless round trip OK" gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
python bench.py --workload rgb_shared --no-comparators > gpurun_out/bench_rgbs.json 2> gpurun_out/bench_rgbs.err; echo rgbs rc=$?
python bench.py --workload crops --no-comparators --no-cpu-baseline > gpurun_out/bench_crops.json 2> gpurun_out/bench_crops.err; echo crops rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; tail -c 400 gpurun_out/bench_ref.json
