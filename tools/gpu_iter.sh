set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload C2s --steps 3 --warmup 3 > gpurun_out/bench_C2s.json 2> gpurun_out/bench_C2s.err; tail -2 gpurun_out/bench_C2s.err; cat gpurun_out/bench_C2s.json
python bench.py --workload C3s --steps 3 --warmup 3 > gpurun_out/bench_C3s.json 2> gpurun_out/bench_C3s.err; tail -2 gpurun_out/bench_C3s.err; cat gpurun_out/bench_C3s.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_C2s.csv python bench.py --workload C2s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^k_seed$' -s 2 -c 2 -o gpurun_out/prof_kseed python bench.py --workload C2s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
