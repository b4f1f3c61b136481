#!/bin/bash
# Round evidence: bench line, launch list of the timed region, full ncu capture of the headline kernel.
# Run under gpurun (one GPU).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step_S9.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/prof_step_S9 -f $SHORT > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
head -c 1500 gpurun_out/bench_r01.json
