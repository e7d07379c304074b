#!/bin/bash
# round 2, call U: block logic + dedup in a kernel (N4), one host wait per contig in K2 -- tests, contig times, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2u_pytest.txt
tail -8 gpurun_out/r2u_pytest.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tail -2 | tee gpurun_out/r2u_contig.txt
GSA_BLOCK_LOGIC=host timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tail -1 | tee -a gpurun_out/r2u_contig.txt
timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2u_bench_C4_n1.json 2> gpurun_out/r2u_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2u_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','phases_alone_ms_per_step')}); print(j['e2e']['value'], j['e2e']['ms_per_step'])
PY
