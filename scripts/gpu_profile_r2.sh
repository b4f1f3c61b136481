#!/bin/bash
# Round-2 evidence: full parity suite, smoke, the bench line (both arms), the launch list of the timed region and one
# `ncu --set full` capture per kernel family.  Run under gpurun (one GPU).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_bench_S9.json 2> gpurun_out/r02_bench_S9.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --metric demos > gpurun_out/r02_bench_demos_S9.json 2> gpurun_out/r02_bench_demos.err; echo "demos rc=$?"
timeout 600 python bench.py --metric demos --impl reference --steps 2 --warmup 0 > gpurun_out/r02_bench_demos_reference_arm.json 2>> gpurun_out/r02_ref.err; echo "demos ref rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
$SHORT > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step_S9.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
python scripts/time_r2.py > gpurun_out/r02_time_r2.txt 2>&1
python scripts/time_kernels.py > gpurun_out/r02_time_kernels.txt 2>&1
for w in step demo9 demo4 demo16 sample basis9 basis16 basis4 rank rollout expand; do
  timeout 120 python scripts/prof_r2.py $w > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'step_kernel|demo_kernel|demo4_thread|demo_sample|basis_mma|basis_fast|basis4_thread|slice_rank|rollout_kernel|rollout4|expand_kernel|expand4' -s 1 -c 1 -o gpurun_out/r02_prof_$w -f python scripts/prof_r2.py $w > gpurun_out/r02_ncu_$w.log 2>&1
  tail -1 gpurun_out/r02_ncu_$w.log
done
head -c 1200 gpurun_out/r02_bench_S9.json; echo; cat gpurun_out/r02_bench_reference_arm.json | head -c 800
