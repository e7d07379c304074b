#!/bin/bash
# round 2, call S: window starts by pointer jumping over the candidate list (one cooperative launch) -- tests, contig times, bench, PCIe probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2s_pytest.txt
tail -5 gpurun_out/r2s_pytest.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tail -2 | tee gpurun_out/r2s_contig.txt
timeout 900 python bench.py --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2s_bench_C4_n1.json 2> gpurun_out/r2s_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2s_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','phases_alone_ms_per_step')}); print(j['e2e']['value'], j['e2e']['ms_per_step'])
PY
timeout 200 python tools/pcie_probe.py 2>&1 | tee gpurun_out/r2s_pcie.txt
