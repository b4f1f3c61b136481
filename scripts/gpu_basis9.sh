#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:basis_fast -s 2 -c 1 -o gpurun_out/prof_basis_S9c -f python scripts/prof_basis.py 9 > gpurun_out/ncu_basis9c.log 2>&1
tail -1 gpurun_out/ncu_basis9c.log
