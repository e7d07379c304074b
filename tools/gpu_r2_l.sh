#!/bin/bash
# round 2, call L (2 GPUs): outbox race fix -- gather tests, many-contig -gpus 2 test, C4 CLI md5 with -gpus 2; e2e diagnosis on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_cli.py -m gpu -x -q -k "gather or multi_gpu or outbox" 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r2l_pytest_n2.txt
C4_SKIP_REF=1 timeout 900 tools/c4_parity.sh 2 > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2l_c4_cli_n2.txt
grep -E "timing|wall|ours\.|identifies|FatalError" gpurun_out/r2l_c4_cli_n2.txt
echo "expected: b258ea61e14ee5d35df6d05590e995bc ours.maf / 2be8b88de7192c8a56c60b62fe0322a9 ours.vcf"
D=/tmp/gsa_bench_cache/C4
for rep in 1 2; do
  GSA_OUTBOX_RESERVE=0 bin/GSAlign -t 24 -gpus 2 -i $D/ref -q $D/qry.fa -o $D/grow > /dev/null 2>&1; ( cd $D && md5sum grow.maf grow.vcf ); rm -f $D/grow.maf $D/grow.vcf
done
for m in h2d d2h; do
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --e2e-mode $m --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2l_e2e_$m.json 2> /dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/r2l_e2e_$m.json'))
print('e2e mode $m:', j['ms_per_step'], 'ms device-resident;', j['e2e']['ms_per_step'], 'ms', j['e2e']['per_contig_ms'])
PY
done
