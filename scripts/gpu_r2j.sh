#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
timeout 600 python bench.py --no-extras --no-cpu > gpurun_out/r2j_bench.json 2> gpurun_out/r2j.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e']['ceiling_gbs'], 'demos e2e', d['demos']['e2e']['value'], d['demos']['e2e']['frac_of_ceiling'])
PY
