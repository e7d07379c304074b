set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in C5s C2; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['roofline']['frac'])"; done
D=/tmp/gsa_bench_cache/C2
( time GSA_TIMING=1 ./bin/GSAlign -i $D/ref -q $D/qry.fa -o /tmp/ours_c2 ) 2>&1 | grep "timing\|real"
( time GSA_TIMING=1 ./bin/GSAlign -i $D/ref -q $D/qry.fa -o /tmp/ours_c2 ) 2>&1 | grep "timing\|real"
