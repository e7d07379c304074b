#!/bin/bash
# round 2, call Z: compute-sanitizer memcheck over the kernels added this round (k_variants, k_block_logic, FPack, k_jump_all, bulk-copy staging of k_seed,
# warp-parallel traceback), through the parity tests that exercise them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 "$@" > gpurun_out/r2z_$name.txt 2>&1; echo "$name: exit $? ; $(grep -E 'ERROR SUMMARY|passed|failed|smoke ok' gpurun_out/r2z_$name.txt | tr '\n' ' ')"; }
run smoke python __graft_entry__.py smoke
run seams python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "rearranged and rows32 and prm0"
run dpx python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "dpx_classes"
run variants python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "device_variant_records"
run gather python -m pytest tests/test_gpu_gather.py -m gpu -x -q -k "roundtrip and compact"
run lanes python -m pytest tests/test_gpu_lanes.py -m gpu -x -q -k "prefetched or degenerate"
