#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2_bench_n8c.json 2> gpurun_out/r2_bench_n8c.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8c.json'))
print('steps', d['value'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e']['ceiling_gbs'], d['e2e']['achieved_gbs'])
print('demos e2e', d['demos']['e2e'])
PY
