#!/bin/bash
# fresh ncu captures (with source) of the 9x9x9 demo generator and change of basis at HEAD
mkdir -p gpurun_out
for w in demo9 basis9; do
  timeout 120 python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'demo_kernel|basis_mma' -s 1 -c 1 -o gpurun_out/r2n_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/r2n_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2n_ncu_$w.log
done
