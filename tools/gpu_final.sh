# what the driver runs at round end, in one go: GPU tests, smoke, the default bench line and the reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; tail -1 gpurun_out/bench_C2.err | cut -c1-200; cut -c1-200 gpurun_out/bench_C2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_C2_ref.json 2> gpurun_out/bench_C2_ref.err; cut -c1-200 gpurun_out/bench_C2_ref.json
