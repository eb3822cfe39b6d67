python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_10.log 2>&1; tail -3 gpurun_out/pytest_gpu_10.log
python bench.py > gpurun_out/bench_20.log 2>&1; tail -1 gpurun_out/bench_20.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_20.log 2>&1; tail -1 gpurun_out/bench_ref_20.log | cut -c1-400
