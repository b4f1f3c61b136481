#!/bin/bash
# Full GPU check: parity tests, smoke, bench line, per-kernel timings.  Run under gpurun.
set -o pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_head.json
python scripts/time_kernels.py 2>&1 | tee gpurun_out/time_kernels.txt
