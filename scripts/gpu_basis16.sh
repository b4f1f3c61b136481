#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_aux_basis_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in 0 2 3 5 1; do TG_BASIS_VARIANT=$v timeout 300 python scripts/time_basis16.py 18 0.03; done 2>&1 | tee gpurun_out/time_basis16.txt
for v in 0 2; do TG_BASIS_VARIANT=$v timeout 300 python scripts/time_basis16.py 18 0.1; done 2>&1 | tee -a gpurun_out/time_basis16.txt
for v in 0; do
TG_BASIS_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:basis_mma16 -s 2 -c 1 -o gpurun_out/prof_basis_mma16_f16 -f python scripts/prof_basis.py 16 > gpurun_out/ncu_mma16_f16.log 2>&1
tail -2 gpurun_out/ncu_mma16_f16.log
done
