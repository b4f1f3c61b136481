#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demo_gen_gpu.py tests/test_configs_gpu.py tests/test_dropin_gpu.py tests/test_step_gpu.py tests/test_rollout_ustream_gpu.py -x -q -m gpu 2>&1 | tail -4
python scripts/time_kernels.py 2>&1 | grep -E "demo_gen|accumulate" | tee gpurun_out/time_demo_sparse.txt
