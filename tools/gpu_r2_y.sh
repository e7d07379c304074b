#!/bin/bash
# round 2, call Y: the phases of ONE contig (device-resident, nothing else on the GPU) while unrelated bulk copies keep the PCIe link busy
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for bg in none h2d d2h both; do echo "== background copies: $bg"; timeout 300 python tools/prof_contig.py --reps 5 --bg $bg 2>&1 | grep -v "^\[bench" | tail -2; done | tee gpurun_out/r2y_bg_copies.txt
