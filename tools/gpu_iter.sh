set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; cat gpurun_out/bench_C2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
