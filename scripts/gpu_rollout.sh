#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_ustream_gpu.py tests/test_configs_gpu.py tests/test_step_gpu.py tests/test_dropin_gpu.py tests/test_demo_gen_gpu.py -x -q -m gpu 2>&1 | tail -5
python scripts/time_kernels.py 2>&1 | grep -E "rollout|expand" | tee gpurun_out/time_rollout.txt
timeout 600 python scripts/rollout_sweep.py 16,20,22 2>&1 | grep -v config | tee gpurun_out/rollout_sweep.txt
python bench.py --steps 20 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S9 step', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
