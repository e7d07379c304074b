set -x
mkdir -p gpurun_out
python tools/bench_dp.py --out gpurun_out/dp_bench.json 2>&1 | tail -12
for w in C3s C5s; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['counts_per_step'])"; done
