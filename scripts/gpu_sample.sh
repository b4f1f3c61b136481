#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_aux_basis_gpu.py tests/test_dropin_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -4
python scripts/time_sample.py 2>&1 | tee gpurun_out/time_sample.txt
