#!/bin/bash
# round 2, call F: warp-parallel traceback + acquire/release hand-over in k_dpx -- tests, per-contig phase times, PCIe probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2f_pytest.txt
tail -4 gpurun_out/r2f_pytest.txt
timeout 120 python tools/pcie_probe.py 2>&1 | tee gpurun_out/r2f_pcie.txt
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2f_contig.txt
timeout 300 python tools/bench_dp.py 2>&1 | tail -12 | tee gpurun_out/r2f_dp.txt
