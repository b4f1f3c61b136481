#!/bin/bash
# K8 three-stage expand: tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "expand or child or mcts or step or dropin" 2>&1 | tail -3
timeout 600 python scripts/time_keys.py 2>&1 | tee gpurun_out/r2v_expand.txt | tail -12
