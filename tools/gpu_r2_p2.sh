#!/bin/bash
# round 2, final validation on two GPUs: gather / multi-GPU tests and the C4 CLI md5 with -gpus 2 on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_cli.py -m gpu -q -k "gather or multi_gpu or outbox" 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/r2final_pytest_n2.txt
C4_SKIP_REF=1 timeout 900 tools/c4_parity.sh 2 > /dev/null 2>&1
cp gpurun_out/c4_parity.txt gpurun_out/r2final_c4_cli_n2.txt
grep -E "timing|ours\.|identifies|FatalError" gpurun_out/r2final_c4_cli_n2.txt
echo "expected: b258ea61e14ee5d35df6d05590e995bc ours.maf / 2be8b88de7192c8a56c60b62fe0322a9 ours.vcf"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-dp-stress > gpurun_out/bench_C4_n2.json 2> gpurun_out/bench_C4_n2.err
tail -1 gpurun_out/bench_C4_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'files', d['e2e_files'])"
