#!/bin/bash
# K4 ticket scheduling: sample tests + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "sample or getitem or dataset or buffers or dropin or int16" 2>&1 | tail -3
timeout 600 python scripts/time_r2.py 2>&1 | grep "S=9 demo_sample" | tee gpurun_out/r2r_sample.txt
