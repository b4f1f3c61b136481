#!/bin/bash
# 4x4x4 row-owner sample batcher: tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "sample or getitem or dataset or buffers or dropin or int16 or reference_unchanged or alignment" 2>&1 | tail -3
timeout 600 python scripts/time_r2.py 2>&1 | grep "S=4 demo_sample" | tee gpurun_out/r2u_sample4.txt
