#!/bin/bash
# round 2, call K: prefetched uploads (double buffering) in e2e, full-warp peer sums, S2 pipeline -- tests + bench + K1 capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2k_pytest.txt
tail -4 gpurun_out/r2k_pytest.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tail -2 | tee gpurun_out/r2k_contig.txt
timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2k_bench_C4_n1.json 2> gpurun_out/r2k_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2k_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','phases_alone_ms_per_step')}); print(j['e2e'])
PY
timeout 900 python bench.py --lanes 8 --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2k_bench_C4_n1_l8.json 2> /dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2k_bench_C4_n1_l8.json'))
print('8 lanes', {k:j[k] for k in ('value','ms_per_step')}); print(j['e2e'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k k_seed -s 1 -c 1 -f -o gpurun_out/prof_kseed_C4 python tools/prof_contig.py --reps 2 > gpurun_out/r2k_ncu2.log 2>&1
tail -3 gpurun_out/r2k_ncu2.log
