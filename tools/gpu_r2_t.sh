#!/bin/bash
# round 2, call T: k_seed at 6 vs 8 CTAs per SM (80 vs 64 registers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in 6 8 6 8; do echo "== GSA_SEED_OCC=$o"; GSA_SEED_OCC=$o timeout 300 python tools/prof_contig.py --reps 4 2>&1 | grep -v "^\[bench" | tail -2; done | tee gpurun_out/r2t_seed_occ.txt
GSA_SEED_OCC=8 timeout 600 python -m pytest tests/test_gpu_seed.py -m gpu -x -q 2>&1 | tail -2
GSA_SEED_OCC=8 timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2t_bench_occ8.json 2> /dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2t_bench_occ8.json'))
print('occ 8:', {k:j[k] for k in ('value','ms_per_step','phases_alone_ms_per_step')}); print(j['e2e']['value'], j['e2e']['ms_per_step'])
PY
