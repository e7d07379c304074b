#!/bin/bash
# round 2, final artefacts (1 GPU): the default bench line, the reference arm, bench lines of the other configs, launch list,
# K1 capture in -sen mode, CLI timing at C3, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/bench_*.json
s=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_C4_n1.json 2> gpurun_out/bench_C4_n1.err
echo "default bench wall: $(( $(date +%s) - s )) s"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_C4_n1.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', j['e2e']['value'], 'files', j['e2e_files'] and j['e2e_files']['value'], 'cpu', j['cpu_baseline'] and j['cpu_baseline']['value'])
print('roofline', {k: j['roofline'][k] for k in ('achieved','frac','traffic','ms_per_launch')}, 'k2', j['roofline_k2']['frac'], 'k3', j['roofline_k3']['frac'], j['roofline_k3']['gcups'], 'stress', j['roofline_k3_stress']['frac'])
print(j['clocks'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_C4_reference.json 2> /dev/null; cut -c1-400 gpurun_out/bench_C4_reference.json
for w in C2 C3 C5; do
  timeout 900 python bench.py --workload $w --no-files --no-cpu-baseline --no-dp-stress --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_$w.json'))
print('$w', {k:j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], 'k3', j['roofline_k3'] and (j['roofline_k3']['frac'], j['roofline_k3']['gcups']), 'alone', j['phases_alone_ms_per_step'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_C4_contig.csv python tools/prof_contig.py --reps 2 > gpurun_out/r2v_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k k_seed -s 1 -c 1 -f -o gpurun_out/prof_kseed_C5 python tools/prof_contig.py --workload C5 --reps 2 > gpurun_out/r2v_ncu2.log 2>&1
tail -2 gpurun_out/r2v_ncu2.log
D=/tmp/gsa_bench_cache/C3
GSA_TIMING=1 bin/GSAlign -t $(nproc) -i $D/ref -q $D/qry.fa -o $D/ours 2>&1 | grep -E "timing|identifies" | tee gpurun_out/r2v_c3_cli.txt
rm -f $D/ours.maf $D/ours.vcf
python __graft_entry__.py smoke 2>&1 | tail -1
