#!/bin/bash
# mma9 rework: basis tests + timings + ncu brief of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "basis or int16" 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-100
timeout 600 python scripts/time_r2.py 2>&1 | grep "change_of_basis" | tee gpurun_out/r2m_basis.txt
