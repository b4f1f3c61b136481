#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_int16_gpu.py tests/test_buffers_gpu.py -q 2>&1 | tail -30
for w in sample basis9 basis16; do
  python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'demo_sample_dm|basis_mma' -s 1 -c 1 -o gpurun_out/r2_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/r2_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2_ncu_$w.log
done
