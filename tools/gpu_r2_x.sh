#!/bin/bash
# round 2, call X: e2e leg with the bulk transfers done by a kernel over the mapped host pointers instead of the copy engines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in dma kernel; do
GSA_BULK_COPY=$m timeout 600 python bench.py --no-files --no-cpu-baseline --no-dp-stress --steps 4 > gpurun_out/r2x_bulk_$m.json 2>/dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/r2x_bulk_$m.json'))
print('bulk copies by $m:', 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'Gbp/s', round(j['e2e']['ms_per_step'],1), 'ms', {k: round(v,2) for k,v in j['e2e']['per_contig_ms'].items()})
PY
done
python __graft_entry__.py smoke 2>&1 | tail -1
