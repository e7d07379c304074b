#!/bin/bash
# round 2, call J: hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) A/B on the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2j_bench_conn$c.json 2> gpurun_out/r2j_bench_conn$c.err
python - <<PY
import json
j=json.load(open('gpurun_out/r2j_bench_conn$c.json'))
print('connections $c:', {k:j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
PY
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 900 python bench.py --lanes 8 --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2j_bench_conn32_l8.json 2> /dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/r2j_bench_conn32_l8.json'))
print('connections 32, 8 lanes:', {k:j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
PY
