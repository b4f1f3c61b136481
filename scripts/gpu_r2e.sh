#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_int16_gpu.py tests/test_buffers_gpu.py tests/test_aux_basis_gpu.py tests/test_demo_gen_gpu.py -q 2>&1 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python scripts/time_r2.py 2>&1 | grep "S=9"
