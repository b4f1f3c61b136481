#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demo_gen_gpu.py tests/test_configs_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in 1 0; do TG_DEMO_MMA=$v timeout 300 python scripts/time_demo16.py 17 49; done 2>&1 | tee gpurun_out/time_demo16.txt
TG_DEMO_MMA=1 timeout 300 python scripts/time_demo16.py 17 12 2>&1 | tee -a gpurun_out/time_demo16.txt
