#!/bin/bash
# round 2, call B: new kernels (int16 paths, 9x9x9 MMA change of basis, TMA batcher), full suite, smoke, bench with extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_int16_gpu.py tests/test_buffers_gpu.py -x -q 2>&1 | tail -25
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_int16_gpu.py --deselect tests/test_buffers_gpu.py 2>&1 | tail -25
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print('steps', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'])
print('demos', d['demos']['value'], d['demos']['roofline']['frac'], 'demos4', d['demos_4x4x4']['value'], d['demos_4x4x4']['roofline']['frac'])
e=d['extras']
for k in ('rollout','change_of_basis','demo_sample','expand_children','slice_rank','size_4x4x4','size_16x16x16','config5_rollout_16M_x64','rollout_host'):
    print(k, json.dumps(e.get(k))[:700])
PY
