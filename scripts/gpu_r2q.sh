#!/bin/bash
# K3 9x9x9 tile-shape variants (sweep build)
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do
  echo "variant $v: $(TG_TUNING=1 TG_DEMO_VARIANT=$v timeout 300 python scripts/time_demo.py 9 2>&1 | grep demo_gen)"
done | tee gpurun_out/r2q_variants.txt
