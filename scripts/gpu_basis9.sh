#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_aux_basis_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
python scripts/time_kernels.py 2>&1 | grep -E "change_of_basis" | tee gpurun_out/time_basis_setup.txt
