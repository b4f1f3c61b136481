#!/bin/bash
# ncu captures (with source) of given prof_r2.py workloads: bash scripts/gpu_ncu_capture.sh "demo9 acc9" [kernel regex]
mkdir -p gpurun_out
for w in $1; do
  timeout 120 python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${2:-demo_kernel} -s ${3:-1} -c 1 -o gpurun_out/ncu_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/ncu_log_$w.log 2>&1
  tail -1 gpurun_out/ncu_log_$w.log
done
