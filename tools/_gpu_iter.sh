python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_11.log 2>&1; tail -2 gpurun_out/pytest_gpu_11.log
python bench.py > gpurun_out/bench_21.log 2>&1; tail -1 gpurun_out/bench_21.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_c.log 2>&1
ncu --set full --clock-control none -k regex:"dwt_fwd_level_kernel|pyr_|gap_fill|spiht_encode_kernel|spiht_decode_kernel|dwt_inv_level_kernel" -c 26 -o gpurun_out/prof_r1_final -f python tools/profile_step.py --batch 256 --steps 1 > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log; ls -la gpurun_out
