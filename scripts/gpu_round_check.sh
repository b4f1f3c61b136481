#!/bin/bash
# Round check: full GPU parity suite, smoke, the bench line, the launch list of the timed region.  Run under gpurun.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step_S9.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
head -c 600 gpurun_out/bench_r01.json; echo; tail -c 900 gpurun_out/bench_r01.json
