#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:demo_kernel -s 2 -c 1 -o gpurun_out/prof_demo_S4c -f python scripts/prof_demo.py 4 > gpurun_out/ncu_demo4c.log 2>&1
tail -1 gpurun_out/ncu_demo4c.log
