#!/bin/bash
# round 2, call B: tests with the pack DP kernels, A/B of K3, launch list + k_seed capture at C4 (one contig)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2b_pytest.txt
tail -4 gpurun_out/r2b_pytest.txt
GSA_NO_PACK=1 timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "dp" 2>&1 | tail -2
timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2b_contig_pack.txt
GSA_NO_PACK=1 timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tee gpurun_out/r2b_contig_nopack.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_C4_contig.csv python tools/prof_contig.py --reps 2 > gpurun_out/r2b_ncu1.log 2>&1
tail -2 gpurun_out/r2b_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_seed$' -s 1 -c 1 -f -o gpurun_out/r2b_kseed_C4 python tools/prof_contig.py --reps 2 > gpurun_out/r2b_ncu2.log 2>&1
tail -2 gpurun_out/r2b_ncu2.log
ls -la gpurun_out/ | tail -8
