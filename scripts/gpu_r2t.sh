#!/bin/bash
# 4x4x4 change of basis (thread per game): tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "basis or int16 or config" 2>&1 | tail -3
timeout 600 python scripts/time_r2.py 2>&1 | grep "S=4 change_of_basis" | tee gpurun_out/r2t_basis4.txt
