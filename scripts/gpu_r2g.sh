#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_demo_gen_gpu.py tests/test_configs_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-200
timeout 300 python scripts/time_kernels.py 2>&1 | grep -E "demo_gen|accumulate"
for w in demo9 demo4; do
  timeout 120 python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'demo_kernel' -s 1 -c 1 -o gpurun_out/r2g_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/r2g_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2g_ncu_$w.log
done
