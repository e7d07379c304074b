#!/bin/bash
# round 2, final validation on one GPU: the whole GPU suite on the final tree, then smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2final_pytest.txt
tail -10 gpurun_out/r2final_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -1
