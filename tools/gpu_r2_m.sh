#!/bin/bash
# round 2, call M (8 GPUs of one box): C4 sharded over 8 GPUs -- CLI md5 with -gpus 8, strong-scaling bench at N = 8 and 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
C4_SKIP_REF=1 timeout 900 tools/c4_parity.sh 8 > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2m_c4_cli_n8.txt
grep -E "timing|wall|ours\.|identifies|FatalError" gpurun_out/r2m_c4_cli_n8.txt
echo "expected: b258ea61e14ee5d35df6d05590e995bc ours.maf / 2be8b88de7192c8a56c60b62fe0322a9 ours.vcf"
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2m_bench_C4_n$n.json 2> gpurun_out/r2m_bench_C4_n$n.err
  tail -1 gpurun_out/r2m_bench_C4_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['parallelism'][-150:]); print(d.get('phases_alone_ms_per_step'))"
  grep -E "rank 0" gpurun_out/r2m_bench_C4_n$n.err | tail -2
done
