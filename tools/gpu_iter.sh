set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/bench_dp.py --sizes 64,256,1024 --cells 2e9 2>&1 | cut -c1-110 | head -3
python bench.py --workload C3s --steps 3 --warmup 3 --no-cpu-baseline --no-dp-stress 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_ms_per_step'])"
