#!/bin/bash
# round 2, last call (1 GPU, ~1 minute): the drop-in CLI with the final host code on the C4s pair (240 Mbp x 240 Mbp in 24 contigs
# at C4's rates): stage clock of bin/GSAlign, the reference CLI on the same files and host cores, md5 of both outputs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nproc
python - <<'PY'
import time, bench
t = time.time(); bench.prepare_workload("C4s"); print("prepared C4s in %.1f s" % (time.time() - t))
PY
D=/tmp/gsa_bench_cache/C4s
for i in 1 2; do
  s=$(date +%s.%N)
  GSA_TIMING=1 bin/GSAlign -t $(nproc) -i $D/ref -q $D/qry.fa -o $D/ours 2>&1 | grep -E "timing|identifies|FatalError"
  echo "ours: whole process $(awk -v a=$s -v b=$(date +%s.%N) 'BEGIN{printf "%.2f", b-a}') s"
done
s=$(date +%s.%N)
oracle/_ref/GSAlign -t $(nproc) -i $D/ref -q $D/qry.fa -o $D/theirs 2>&1 | grep -E "took|identifies"
echo "reference -t $(nproc): whole process $(awk -v a=$s -v b=$(date +%s.%N) 'BEGIN{printf "%.2f", b-a}') s"
md5sum $D/ours.maf $D/theirs.maf $D/ours.vcf $D/theirs.vcf
ls -l $D/ours.maf $D/ours.vcf
} 2>&1 | tee gpurun_out/r2_cli_timing_C4s.txt
