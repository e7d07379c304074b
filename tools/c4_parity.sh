#!/bin/bash
# C4 (BASELINE.json config 4: synthetic 3 Gbp vs 3 Gbp, 1 % SNV + 0.1 % indel, 24 contigs of 125 Mbp):
# generates the pair on the box, builds the BWA-format index with bin/gsa_index (blockwise GPU sorter), runs bin/GSAlign and
# the unmodified reference (oracle/_ref/GSAlign -t nproc) on the same files and compares .maf / .vcf by md5.
# Usage: tools/c4_parity.sh [gpus] [workload dir]      output: gpurun_out/c4_parity.txt
set -u
GPUS=${1:-1}
D=${2:-/tmp/gsa_bench_cache/C4}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/c4_parity.txt
mkdir -p "$D" "$ROOT/gpurun_out"
exec > >(tee "$OUT") 2>&1
t() { local s=$(date +%s%N); "$@"; local rc=$?; echo "$(( ($(date +%s%N) - s) / 1000000 )) ms wall"; return $rc; }
echo "== box: $(nproc) cores, $(free -g | awk '/Mem:/{print $2}') GB RAM, $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1) x $(nvidia-smi -L | wc -l)"
if [ ! -f "$D/ok" ]; then
  echo "== generate C4 (gsalign_b200/synth.py write_pair_streams, seed 4)"
  t python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from gsalign_b200 import synth
n, k, snv, indel, seed = synth.CONFIGS["C4"]
print("query bp:", synth.write_pair_streams(n, k, snv, indel, seed, "$D/ref.fa", "$D/qry.fa", workers=8))
PY
  echo "== bin/gsa_index (GPU, blockwise suffix sorter)"
  t "$ROOT/bin/gsa_index" "$D/ref.fa" "$D/ref" && touch "$D/ok"
fi
ls -l "$D"
echo "== bin/GSAlign -gpus $GPUS"
GSA_TIMING=1 t "$ROOT/bin/GSAlign" -t "$(nproc)" -gpus "$GPUS" -i "$D/ref" -q "$D/qry.fa" -o "$D/ours" 2> "$D/ours.err"; echo "exit code $?"
grep -E "timing|Alignment#|identifies|FatalError|wall" "$D/ours.err"
( cd "$D" && md5sum ours.maf ours.vcf && ls -l ours.maf ours.vcf ) | tee "$D/ours.md5"
rm -f "$D/ours.maf"
if [ "${C4_SKIP_REF:-0}" != "1" ]; then
  echo "== oracle/_ref/GSAlign -t $(nproc) (unmodified reference, same index files)"
  t "$ROOT/oracle/_ref/GSAlign" -t "$(nproc)" -i "$D/ref" -q "$D/qry.fa" -o "$D/theirs" 2> "$D/theirs.err" > /dev/null
  grep -E "Alignment#|identifies|It took|wall" "$D/theirs.err"
  ( cd "$D" && md5sum theirs.maf theirs.vcf && ls -l theirs.maf theirs.vcf )
  a=$(cd "$D" && md5sum < theirs.maf); b=$(awk '/ours.maf/{print $1}' "$D/ours.md5" | head -1)
  c=$(cd "$D" && md5sum < theirs.vcf); e=$(awk '/ours.vcf/{print $1}' "$D/ours.md5" | head -1)
  if [ -s "$D/theirs.vcf" ] && [ -n "$b" ] && [ "${a%% *}" = "$b" ] && [ "${c%% *}" = "$e" ]; then echo "C4: .maf and .vcf byte-identical (md5)"; else echo "C4: MISMATCH"; fi
  rm -f "$D/theirs.maf"
fi
