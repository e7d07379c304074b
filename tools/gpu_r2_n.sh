#!/bin/bash
# round 2, call N: bulk copies in pieces / small copies by kernel or DMA -- the e2e leg under each combination
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-files --no-cpu-baseline --no-dp-stress --steps 4 --warmup 3 > gpurun_out/r2n_$name.json 2> /dev/null
  python - <<PY
import json
j=json.load(open('gpurun_out/r2n_$name.json'))
print('$name:', 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'Gbp/s', round(j['e2e']['ms_per_step'],1), 'ms', {k: round(v,2) for k,v in j['e2e']['per_contig_ms'].items()})
PY
}
run chunk4_kernel GSA_COPY_CHUNK_MB=4
run chunk0_kernel GSA_COPY_CHUNK_MB=0
run chunk1_kernel GSA_COPY_CHUNK_MB=1
run chunk4_dma GSA_COPY_CHUNK_MB=4 GSA_SMALL_COPY=dma
run chunk1_dma GSA_COPY_CHUNK_MB=1 GSA_SMALL_COPY=dma
