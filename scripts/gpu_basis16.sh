#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_aux_basis_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in 0 1; do TG_BASIS_VARIANT=$v timeout 300 python scripts/time_basis16.py 18 0.03; done 2>&1 | tee gpurun_out/time_basis16.txt
TG_BASIS_VARIANT=0 timeout 300 python scripts/time_basis16.py 18 0.1 2>&1 | tee -a gpurun_out/time_basis16.txt
