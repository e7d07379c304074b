#!/bin/bash
# round 2, call Q (8 GPUs): strong-scaling bench at N = 8 with the compact record form (and the plain one for comparison)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for form in compact plain; do
  raw=0; [ $form = plain ] && raw=1
  GSA_GATHER_RAW=$raw timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-files --no-cpu-baseline --no-dp-stress > gpurun_out/r2q_bench_C4_n8_$form.json 2> gpurun_out/r2q_bench_C4_n8_$form.err
  tail -1 gpurun_out/r2q_bench_C4_n8_$form.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$form N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['parallelism'][-150:])"
done
