set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_S9.json 2> gpurun_out/bench_S9.err; tail -3 gpurun_out/bench_S9.err; cat gpurun_out/bench_S9.json
for c in 1 2 3 4 5 6; do python bench.py --steps 20 --no-e2e --no-cpu --ctas-per-sm $c | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S9 ctas', $c, d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done
python bench.py --steps 20 --no-e2e --no-cpu --size 4 --games 8388608 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S4', d['ms_per_step'], d['value'], d['roofline']['frac'])"
python bench.py --steps 20 --no-e2e --no-cpu --size 16 --games 131072 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S16', d['ms_per_step'], d['value'], d['roofline']['frac'])"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_S9.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/prof_step_S9 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
