#!/bin/bash
# round 2, call G: N3 (device variant records), small transfers off the copy engines -- tests, C4 CLI md5 + timing, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2g_pytest.txt
tail -6 gpurun_out/r2g_pytest.txt
C4_SKIP_REF=1 timeout 900 tools/c4_parity.sh 1 > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2g_c4_cli.txt
grep -E "timing|wall|ours\.|identifies" gpurun_out/r2g_c4_cli.txt
echo "expected: b258ea61e14ee5d35df6d05590e995bc ours.maf / 2be8b88de7192c8a56c60b62fe0322a9 ours.vcf"
D=/tmp/gsa_bench_cache/C4
GSA_VARIANTS=host GSA_TIMING=1 bin/GSAlign -t 16 -i $D/ref -q $D/qry.fa -o $D/hostvar 2>&1 | grep -E "timing" | tee gpurun_out/r2g_c4_cli_hostvar.txt
rm -f $D/hostvar.maf $D/hostvar.vcf
timeout 1200 python bench.py --no-files --no-cpu-baseline > gpurun_out/r2g_bench_C4_n1.json 2> gpurun_out/r2g_bench_C4_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2g_bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','phases_alone_ms_per_step')}); print(j['e2e']); print(j['roofline_k3'])
PY
