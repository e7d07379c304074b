set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
declare -A MAF=( [C3]=427defc9f01e20f4be1347dd8dcdffcd [C5]=01220ffebcec078771eb96585a110ae2 )
declare -A VCF=( [C3]=4096eaad74fc9206c19f55cea2d6d070 [C5]=27822c149abf3c6555d2b47992c55c71 )
for w in C3 C5; do
  python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
  D=/tmp/gsa_bench_cache/$w
  FL=""; [ $w = C5 ] && FL="-sen -slen 10 -idy 70"
  GSA_TIMING=1 ./bin/GSAlign -t 16 -i $D/ref -q $D/qry.fa -o /tmp/ours_$w $FL 2>&1 | grep timing
  m=$(md5sum < /tmp/ours_$w.maf | cut -d' ' -f1); v=$(md5sum < /tmp/ours_$w.vcf | cut -d' ' -f1)
  [ "$m" = "${MAF[$w]}" ] && [ "$v" = "${VCF[$w]}" ] && echo "$w: md5 of .maf/.vcf equal the reference's (profiles/r1_parity_at_scale.txt)" || echo "$w: MISMATCH $m $v"
  rm -f /tmp/ours_$w.*
done
