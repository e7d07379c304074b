# byte-for-byte parity of bin/GSAlign against the unmodified reference at BASELINE's full sizes (C3, C5)
set -x
mkdir -p gpurun_out
OUT=gpurun_out/parity_at_scale.txt; : > $OUT
TIMEFORMAT="%R s wall"
for w in C3 C5; do
  python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1   # generates the pair and the index
  D=/tmp/gsa_bench_cache/$w
  FL=""; [ $w = C5 ] && FL="-sen -slen 10 -idy 70"
  echo "== $w: bin/GSAlign -i ref -q qry.fa $FL" >> $OUT
  { time ./bin/GSAlign -i $D/ref -q $D/qry.fa -o /tmp/ours_$w $FL 2> /tmp/ours_$w.err ; } 2>> $OUT
  echo "== $w: oracle/_ref/GSAlign -t $(nproc) (unmodified reference)" >> $OUT
  { time ./oracle/_ref/GSAlign -t $(nproc) -i $D/ref -q $D/qry.fa -o /tmp/ref_$w $FL 2> /tmp/ref_$w.err ; } 2>> $OUT
  grep "Alignment#\|identifies" /tmp/ours_$w.err >> $OUT
  ( cd /tmp && md5sum ours_$w.maf ref_$w.maf ours_$w.vcf ref_$w.vcf; ls -l ours_$w.maf ours_$w.vcf | awk '{print $5, $9}' ) >> $OUT
  cmp /tmp/ours_$w.maf /tmp/ref_$w.maf && cmp /tmp/ours_$w.vcf /tmp/ref_$w.vcf && echo "$w: .maf and .vcf byte-identical" >> $OUT
  rm -f /tmp/ours_$w.* /tmp/ref_$w.*
done
cat $OUT
