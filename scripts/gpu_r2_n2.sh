#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"
tail -c 1500 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print('steps', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e']['ceiling_gbs'])
print('demos', d['demos']['value'], d['demos']['roofline']['frac'], 'e2e', d['demos']['e2e'])
print(json.dumps(d['extras']['multi_gpu'], indent=1))
print(json.dumps(d['extras']['config5_rollout_16M_x64']))
PY
