set -x
mkdir -p gpurun_out
python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu-baseline --no-dp-stress > /dev/null 2>&1   # build the workload once
for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload C2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_n$n.json 2> gpurun_out/bench_C2_n$n.err; tail -1 gpurun_out/bench_C2_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
done
