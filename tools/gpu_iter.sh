set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload C2 --steps 3 --warmup 3 > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; grep -v Warn gpurun_out/bench_C2.err | tail -4; cat gpurun_out/bench_C2.json
