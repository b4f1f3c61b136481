#!/bin/bash
# full GPU suite + smoke + key timings after the state-key contract change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160
timeout 300 python scripts/time_keys.py 2>&1 | tee gpurun_out/r2l_keys.txt
timeout 600 python bench.py --no-extras --no-cpu > gpurun_out/r2l_bench.json 2> gpurun_out/r2l.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('steps', d['value'], d['roofline']['frac'], 'demos', d['demos']['value'], d['demos']['roofline']['frac'])
PY
