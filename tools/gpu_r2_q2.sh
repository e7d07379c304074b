#!/bin/bash
# round 2, final tree on 8 GPUs: strong-scaling bench at N = 8 and N = 4 (compact records, block logic in the kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/bench_C4_n$n.json 2> gpurun_out/bench_C4_n$n.err
  tail -1 gpurun_out/bench_C4_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['parallelism'][-140:])"
done
