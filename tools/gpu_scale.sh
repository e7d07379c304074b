set -x
mkdir -p gpurun_out
python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1   # build the workload once
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload C2 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_C2_n8$tag.json 2> gpurun_out/bench_C2_n8$tag.err; tail -1 gpurun_out/bench_C2_n8$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'])"; }
run ""
GSA_BENCH_NO_GATHER=1 run _nogather
run _lanes2 --lanes 2
