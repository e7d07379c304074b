#!/bin/bash
# round 2, call C: tests, per-contig phase times, bench C4 N=1 with the NVML sampler
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2c_pytest.txt
tail -4 gpurun_out/r2c_pytest.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2c_contig_pack.txt
GSA_NO_PACK=1 timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2c_contig_nopack.txt
timeout 900 python bench.py --no-files > gpurun_out/r2c_bench_C4_n1.json 2> gpurun_out/r2c_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2c_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','phases_ms_per_step','phases_alone_ms_per_step','clocks')}); print(j['e2e'])
PY
timeout 600 python bench.py --no-files --lanes 8 --no-cpu-baseline --no-dp-stress > gpurun_out/r2c_bench_C4_n1_l8.json 2> /dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2c_bench_C4_n1_l8.json'))
print('lanes 8:',{k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'])
PY
