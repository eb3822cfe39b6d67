python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_13.log 2>&1; tail -2 gpurun_out/pytest_gpu_13.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_23.log 2>&1; tail -1 gpurun_out/bench_23.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_e.log 2>&1
ncu --set full --clock-control none -k regex:"dwt_fwd_level_kernel|pyr_|gap_fill|spiht_encode_kernel" -c 12 -o /tmp/prof_enc_final -f python tools/profile_step.py --batch 256 --steps 1 > gpurun_out/ncu_final2.log 2>&1; tail -1 gpurun_out/ncu_final2.log
python tools/ncu_summary.py /tmp/prof_enc_final.ncu-rep > gpurun_out/r01_kernels_b256_encode.md; wc -l gpurun_out/r01_kernels_b256_encode.md
