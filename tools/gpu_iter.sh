set -x
mkdir -p gpurun_out
python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; tail -2 gpurun_out/bench_C2.err | cut -c1-300; cut -c1-600 gpurun_out/bench_C2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_n2.json 2> gpurun_out/bench_C2_n2.err; tail -3 gpurun_out/bench_C2_n2.err | cut -c1-300; cut -c1-900 gpurun_out/bench_C2_n2.json
