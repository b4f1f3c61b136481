#!/bin/bash
# K2: rollout / replay tests + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "rollout or replay or config or take_actions or dataset or dropin or smoke or host" 2>&1 | tail -3
timeout 600 python scripts/time_kernels.py 2>&1 | grep -i "rollout\|replay" | tee gpurun_out/rollout_check.txt
