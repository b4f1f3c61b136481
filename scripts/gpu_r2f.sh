#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_int16_gpu.py -q -k "demo_store or wide" 2>&1 | tail -5
timeout 300 python scripts/time_r2.py 2>&1 | grep "S=9 demo_sample"
