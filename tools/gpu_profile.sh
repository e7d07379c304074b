# round profile run: launch list, full captures of K1 and K3, bench lines at C2 / C3 / C5 (B200, one GPU)
set -x
mkdir -p gpurun_out
python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1   # builds the C2 index outside the profiled runs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_C2.csv python bench.py --workload C2 --steps 1 --warmup 0 --lanes 1 --no-cpu-baseline --no-dp-stress > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k k_seed -s 2 -c 2 -f -o gpurun_out/prof_kseed python bench.py --workload C2 --steps 1 --warmup 0 --lanes 1 --no-cpu-baseline --no-dp-stress > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k k_dpx -c 6 -f -o gpurun_out/prof_kdpx python tools/bench_dp.py --sizes 64,1024 --cells 2e9 --reps 1 > gpurun_out/ncu_c.log 2>&1
python tools/bench_dp.py --out gpurun_out/dp_bench.json > gpurun_out/dp_bench.log 2>&1
./bin/bench_random_access > gpurun_out/random_access.json 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --workload C2 --lanes 1 --no-cpu-baseline --no-dp-stress > gpurun_out/bench_C2_1lane.json 2> /dev/null
python bench.py --workload C2 > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; tail -2 gpurun_out/bench_C2.err | cut -c1-200; cut -c1-300 gpurun_out/bench_C2.json
python bench.py --workload C2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_C2_ref.json 2> gpurun_out/bench_C2_ref.err; cut -c1-300 gpurun_out/bench_C2_ref.json
for w in C5 C3; do timeout 1500 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; grep -v Warn gpurun_out/bench_$w.err | tail -4 | cut -c1-300; cut -c1-400 gpurun_out/bench_$w.json; done
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2
