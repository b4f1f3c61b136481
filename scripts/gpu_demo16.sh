#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do TG_ACC_VARIANT=$v timeout 300 python scripts/time_demo16.py 17 49 | grep -E "accumulate:" | sed "s/^/acc_variant=$v /"; done 2>&1 | tee gpurun_out/time_acc16_variants.txt
ncu --set full --clock-control none --import-source on -k regex:demo_kernel -s 3 -c 1 -o gpurun_out/prof_demo16_fused -f python scripts/time_demo16.py 15 49 > gpurun_out/ncu_demo16_fused.log 2>&1
tail -1 gpurun_out/ncu_demo16_fused.log
