#!/bin/bash
# round 2, call A: GPU tests, default bench (C4, N=1), ncu capture of k_seed at C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 900 python bench.py > gpurun_out/r2a_bench_C4_n1.json 2> gpurun_out/r2a_bench_C4_n1.err
tail -3 gpurun_out/r2a_bench_C4_n1.err; cat gpurun_out/r2a_bench_C4_n1.json | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seed -s 3 -c 1 -f -o gpurun_out/r2a_kseed_C4 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-dp-stress --no-files --lanes 1 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
ls -la gpurun_out/
