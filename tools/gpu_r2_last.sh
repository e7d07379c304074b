#!/bin/bash
# round 2: the default bench line and the reference arm on the final tree (6 contigs in flight by default)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
s=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_C4_n1.json 2> gpurun_out/bench_C4_n1.err
echo "default bench wall: $(( $(date +%s) - s )) s"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'files', j['e2e_files'] and (j['e2e_files']['value'], j['e2e_files']['seconds']), 'cpu', j['cpu_baseline'] and j['cpu_baseline']['value'])
print('roofline', {k: j['roofline'][k] for k in ('achieved','frac','ms_per_launch','share_of_step')}, 'k2', j['roofline_k2']['frac'], j['roofline_k2']['ms_per_contig'], 'k3', j['roofline_k3']['frac'], j['roofline_k3']['gcups'], 'stress', j['roofline_k3_stress']['frac'], j['roofline_k3_stress']['gcups'])
print(j['phases_alone_ms_per_step'], j['config']['parallelism'][:160])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_C4_reference.json 2> /dev/null; cut -c1-200 gpurun_out/bench_C4_reference.json
