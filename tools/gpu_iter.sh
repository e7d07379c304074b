set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in C2 C3s C3 C5s; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; grep -v Warn gpurun_out/bench_$w.err | tail -2 | cut -c1-300; cat gpurun_out/bench_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['roofline']['frac'], d['roofline']['note'][-330:-200])"; done
