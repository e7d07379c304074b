set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --workload C2s --steps 3 --warmup 3 > gpurun_out/bench_C2s.json 2> gpurun_out/bench_C2s.err; tail -3 gpurun_out/bench_C2s.err; cat gpurun_out/bench_C2s.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C2s.csv python bench.py --workload C2s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_seed -s 2 -c 2 -o gpurun_out/prof_kseed python bench.py --workload C2s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
python bench.py --workload C2 --steps 3 --warmup 3 > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; tail -5 gpurun_out/bench_C2.err; cat gpurun_out/bench_C2.json
nproc; lscpu | grep "Model name"
