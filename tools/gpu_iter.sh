set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload C2 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
D=/tmp/gsa_bench_cache/C2
( time GSA_TIMING=1 ./bin/GSAlign -i $D/ref -q $D/qry.fa -o /tmp/ours_c2 ) 2>&1 | grep "timing\|real"
( time GSA_TIMING=1 ./bin/GSAlign -i $D/ref -q $D/qry.fa -o /tmp/ours_c2 ) 2>&1 | grep "timing\|real"
( time ./oracle/_ref/GSAlign -t 16 -i $D/ref -q $D/qry.fa -o /tmp/ref_c2 ) 2>&1 | grep real
md5sum /tmp/ours_c2.maf /tmp/ref_c2.maf /tmp/ours_c2.vcf /tmp/ref_c2.vcf
