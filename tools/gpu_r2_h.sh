#!/bin/bash
# round 2, call H: L2 fetch granularity A/B for K1, 8-lane bench, launch list + full captures of K1 / K3 on a C4 contig, racecheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 64 32 128; do
  echo "== GSA_L2_FETCH=$g"; GSA_L2_FETCH=$g timeout 300 python tools/prof_contig.py --reps 3 2>&1 | grep -v "^\[bench" | tail -2
done | tee gpurun_out/r2h_l2fetch.txt
timeout 900 python bench.py --lanes 8 --no-files --no-cpu-baseline > gpurun_out/r2h_bench_C4_n1_l8.json 2> gpurun_out/r2h_bench_C4_n1_l8.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2h_bench_C4_n1_l8.json'))
print('8 lanes:', {k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['e2e']['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_C4_contig.csv python tools/prof_contig.py --reps 2 > gpurun_out/r2h_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seed -s 1 -c 1 -f -o gpurun_out/prof_kseed_C4 python tools/prof_contig.py --reps 2 > gpurun_out/r2h_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dpx -s 9 -c 9 -f -o gpurun_out/prof_kdpx_C4 python tools/prof_contig.py --reps 2 > gpurun_out/r2h_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_variants|k_gap_similarity" -c 3 -f -o gpurun_out/prof_misc_C4 python tools/prof_contig.py --reps 1 > gpurun_out/r2h_ncu4.log 2>&1
tail -2 gpurun_out/r2h_ncu2.log gpurun_out/r2h_ncu3.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_pipeline.py -k "dpx_classes" -x -q > gpurun_out/r2h_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2h_racecheck.txt | sort | uniq -c | tail -8
ls -la gpurun_out | tail -12
