#!/bin/bash
# K3: demo tests + timings (sizes as argument, default 4,9)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demo_gen_gpu.py tests/test_dist_gloo.py tests/test_configs_gpu.py -q -x 2>&1 | tail -3
timeout 600 python scripts/time_demo.py ${1:-4,9} 2>&1 | tee gpurun_out/demo_check.txt
