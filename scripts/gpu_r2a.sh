#!/bin/bash
# round 2, call A: parity suite after the token-bound / loader changes, reference-unchanged test, bench without extras
mkdir -p gpurun_out
ls baseline/_ref | head -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --no-extras > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python scripts/time_kernels.py > gpurun_out/r2a_time_kernels.txt 2>&1; echo "time rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err; head -c 3000 gpurun_out/r2a_bench.json; echo; cat gpurun_out/r2a_ref.json | head -c 1500
