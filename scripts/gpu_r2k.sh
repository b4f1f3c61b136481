#!/bin/bash
# product-structured state key: key/expand/MCTS tests, smoke, expand timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "key or expand or mcts or act or child" 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160
timeout 600 python scripts/time_kernels.py 2>&1 | grep -i "expand\|state_key" | tee gpurun_out/r2k_expand.txt
timeout 300 python scripts/time_keys.py 2>&1 | tee -a gpurun_out/r2k_expand.txt
