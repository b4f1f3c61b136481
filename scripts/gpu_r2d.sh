#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_demo_gen_gpu.py -x -q 2>&1 | tail -8
timeout 300 python -m pytest tests/test_int16_gpu.py tests/test_buffers_gpu.py -q 2>&1 | tail -8
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_int16_gpu.py --deselect tests/test_buffers_gpu.py --deselect tests/test_demo_gen_gpu.py 2>&1 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python scripts/time_r2.py 2>&1 | tail -20
for w in sample demo9 demo4 basis9; do
  timeout 120 python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'demo_sample|stream_kernel|basis_mma' -s 1 -c 1 -o gpurun_out/r2d_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/r2d_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2d_ncu_$w.log
done
