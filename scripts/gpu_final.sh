#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_round_check.sh 2>&1 | tail -8 | cut -c1-300
python scripts/time_kernels.py > gpurun_out/time_kernels.txt 2>&1; tail -3 gpurun_out/time_kernels.txt
python scripts/time_aux.py > gpurun_out/time_aux.txt 2>&1; tail -2 gpurun_out/time_aux.txt
