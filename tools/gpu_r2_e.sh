#!/bin/bash
# round 2, call E: mapped-file emitters -- tests, C4 CLI parity + timing, the full default bench (files + cpu baseline)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2e_pytest.txt
tail -4 gpurun_out/r2e_pytest.txt
timeout 900 tools/c4_parity.sh 1 > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2e_c4_parity.txt
grep -E "timing|wall|identical|MISMATCH|took" gpurun_out/r2e_c4_parity.txt
s=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2e_bench_C4_n1.json 2> gpurun_out/r2e_bench_C4_n1.err
echo "bench default wall: $(( $(date +%s) - s )) s"
tail -3 gpurun_out/r2e_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2e_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','e2e_files','cpu_baseline')}); print(j['e2e']['value'])
PY
