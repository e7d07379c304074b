#!/bin/bash
# round 2: lanes per GPU on the final tree (same box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for l in 4 6 8 4 8; do
timeout 600 python bench.py --lanes $l --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2_lanes_$l.json 2>/dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/r2_lanes_$l.json'))
print('lanes $l:', 'value', round(j['value'],2), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), round(j['e2e']['ms_per_step'],1))
PY
done | tee gpurun_out/r2_lanes.txt
