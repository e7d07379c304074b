#!/bin/bash
# round 2, call I (N GPUs of one box): the NCCL record gather in tests, in the CLI (C4 md5) and in the strong-scaling bench
# usage: tools/gpu_r2_i.sh <N>
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_cli.py -m gpu -x -q -k "gather or multi_gpu or outbox" 2>&1 | grep -v "^$" | tail -5 | tee gpurun_out/r2i_pytest_n$N.txt
C4_SKIP_REF=1 timeout 900 tools/c4_parity.sh $N > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2i_c4_cli_n$N.txt
grep -E "timing|wall|ours\.|identifies|FatalError" gpurun_out/r2i_c4_cli_n$N.txt
echo "expected: b258ea61e14ee5d35df6d05590e995bc ours.maf / 2be8b88de7192c8a56c60b62fe0322a9 ours.vcf"
for n in $(echo 1 $N | tr ' ' '\n' | sort -un); do
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2i_bench_C4_n1.json 2> gpurun_out/r2i_bench_C4_n1.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2i_bench_C4_n$n.json 2> gpurun_out/r2i_bench_C4_n$n.err; fi
  tail -1 gpurun_out/r2i_bench_C4_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'][-120:])"
  grep -E "gather|rank 0" gpurun_out/r2i_bench_C4_n$n.err | tail -4
done
