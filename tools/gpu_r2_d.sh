#!/bin/bash
# round 2, call D: the chained-scan K2 -- tests, per-contig phase times, launch list, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2d_pytest.txt
tail -4 gpurun_out/r2d_pytest.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2d_contig.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_C4_contig.csv python tools/prof_contig.py --reps 2 > gpurun_out/r2d_ncu1.log 2>&1
timeout 900 python bench.py --no-files --no-cpu-baseline > gpurun_out/r2d_bench_C4_n1.json 2> gpurun_out/r2d_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2d_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','phases_ms_per_step','phases_alone_ms_per_step')}); print(j['e2e']['value'])
PY
