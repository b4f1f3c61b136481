#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_aux_basis_gpu.py -x -q -m gpu -k "basis" 2>&1 | tail -3
for v in 0 4 5 6 0 5; do TG_BASIS_VARIANT=$v timeout 300 python scripts/time_basis16.py 18 0.03; done 2>&1 | grep variant | tee gpurun_out/time_basis16.txt
