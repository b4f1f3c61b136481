#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_demo_gen_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/time_kernels.py 2>&1 | grep -E "demo_gen"
for v in 0 1 2 3 4 5; do echo "variant $v"; TG_TUNING=1 TG_DEMO_VARIANT=$v timeout 300 python scripts/time_kernels.py 2>&1 | grep -E "S=9 demo_gen"; done
