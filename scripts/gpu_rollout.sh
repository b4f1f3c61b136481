#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/prof_rollout_lazy -f python scripts/prof_rollout.py > gpurun_out/ncu_rollout_lazy.log 2>&1
tail -1 gpurun_out/ncu_rollout_lazy.log
