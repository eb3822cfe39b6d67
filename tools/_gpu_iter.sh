python -m pytest tests/test_gpu_transform.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-decode --e2e-steps 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'])"
ncu --set full --clock-control none --import-source on -k regex:"dwt_fwd_level_kernel" -c 2 -o gpurun_out/prof_tmp -f python tools/profile_step.py --batch 64 --steps 1 > gpurun_out/ncu_tmp.log 2>&1; tail -1 gpurun_out/ncu_tmp.log
