#!/bin/bash
# round 2, call W: block sums of k_frag_simple aggregated per CTA; C5 K1 without the min-blocks hint; chained-scan tile size A/B (rebuilt on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_lanes.py -m gpu -x -q 2>&1 | tail -2
echo "== CH_ITEMS 4 (as committed)"; timeout 300 python tools/prof_contig.py --reps 4 2>&1 | grep -v "^\[bench" | tail -2 | tee gpurun_out/r2w_contig.txt
echo "== C5 contig"; timeout 300 python tools/prof_contig.py --workload C5 --reps 3 2>&1 | grep -v "^\[bench" | tail -1 | tee -a gpurun_out/r2w_contig.txt
sed -i 's/#define CH_ITEMS 4/#define CH_ITEMS 8/' gsalign_b200/csrc/scan.cuh
make -s -C gsalign_b200/csrc -j8 2>&1 | grep -v "^$" | head -3
echo "== CH_ITEMS 8"; timeout 300 python tools/prof_contig.py --reps 4 2>&1 | grep -v "^\[bench" | tail -2 | tee -a gpurun_out/r2w_contig.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "seams" 2>&1 | tail -2
timeout 600 python bench.py --no-files --no-cpu-baseline --no-dp-stress --steps 4 > gpurun_out/r2w_bench_items8.json 2>/dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2w_bench_items8.json'))
print('items 8:', {k:j[k] for k in ('value','ms_per_step','phases_alone_ms_per_step')})
PY
