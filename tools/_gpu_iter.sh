python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_12.log 2>&1; tail -2 gpurun_out/pytest_gpu_12.log
python bench.py > gpurun_out/bench_22.log 2>&1; tail -1 gpurun_out/bench_22.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_d.log 2>&1
ncu --set full --clock-control none -k regex:"dwt_fwd_level_kernel|pyr_|gap_fill|spiht_encode_kernel|spiht_decode_kernel|dwt_inv_level_kernel" -c 26 -o /tmp/prof_r1_final -f python tools/profile_step.py --batch 256 --steps 1 > gpurun_out/ncu_final.log 2>&1; tail -1 gpurun_out/ncu_final.log
python tools/ncu_summary.py /tmp/prof_r1_final.ncu-rep > gpurun_out/r01_kernels_b256_ncu_full.md; wc -l gpurun_out/r01_kernels_b256_ncu_full.md; ls -la gpurun_out | head
