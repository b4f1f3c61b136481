#!/bin/bash
mkdir -p gpurun_out
for v in 0 7 8; do TG_DEMO_VARIANT=$v timeout 300 python scripts/time_demo16.py 17 49 | grep -E "demo_gen|checksum" | sed "s/^/variant=$v /"; done 2>&1 | tee gpurun_out/time_demo16_nt.txt
