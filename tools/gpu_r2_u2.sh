#!/bin/bash
# round 2, call U2: the GPU suite with the block logic in a kernel (N4) and the repeat-rich parity input
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2u2_pytest.txt
tail -12 gpurun_out/r2u2_pytest.txt
