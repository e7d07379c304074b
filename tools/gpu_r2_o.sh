#!/bin/bash
# round 2, call O: compact record form -- the GPU suite, then one bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2o_pytest.txt
tail -6 gpurun_out/r2o_pytest.txt
timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2o_bench_C4_n1.json 2> gpurun_out/r2o_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2o_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','phases_alone_ms_per_step')}); print(j['e2e']['value'], j['e2e']['ms_per_step'])
PY
