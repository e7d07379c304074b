"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep, bench_*.json} into tracked summaries under profiles/.
usage: python tools/summarize_profile.py <round-tag>"""
import collections, csv, glob, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs("profiles", exist_ok=True)
out = [f"# profiles/{tag}: ncu launch lists, full captures and bench lines (B200, sm_100a)\n"]
for path in sorted(glob.glob("gpurun_out/launches_*.csv")):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"]))[:50]
        if re.search(r"k_ib_|k_fill_sa|k_build_|k_dpx_peak", name):   # index construction / upload / microbenchmark: setup, not the path
            continue
        agg[name][0] += 1; agg[name][1] += v; tot += v
    out.append(f"\n## launch list {os.path.basename(path)} (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)\n")
    out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
        out.append(f"| {k} | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    out.append(f"| **all** | {sum(c for c, _ in agg.values())} | {tot:.1f} | 100% |")
def _num(x):
    return float(x.replace(",", ""))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__inst_executed_pipe_alu.sum", "smsp__inst_executed_op_shared_ld.sum"]
for path in sorted(glob.glob("gpurun_out/prof_*.ncu-rep")):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(r.splitlines()))
    if len(rows) < 3:
        continue
    hdr = rows[0]
    out.append(f"\n## full capture {os.path.basename(path)} (`ncu --set full --clock-control none --import-source on`)\n")
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |\n|---|---|" + "---|" * (len(rows) - 2))
    ki = hdr.index("Kernel Name")
    out.append("| kernel | | " + " | ".join(re.sub(r"\(.*", "", row[ki])[:40] for row in rows[2:]) + " |")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {rows[1][i]} | " + " | ".join(row[i] for row in rows[2:]) + " |")
# dram bytes per k_seed launch of the full captures -> profiles/traffic.json (bench.py reports it as roofline.traffic)
tpath = "profiles/traffic.json"
tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
for wl, kp, bp, how in (("C2", "gpurun_out/prof_kseed.ncu-rep", 25_000_000, "bench.py --workload C2 --lanes 1"),
                        ("C4", "gpurun_out/prof_kseed_C4.ncu-rep", 125_000_247, "tools/prof_contig.py --workload C4 --contig 0: one 125 Mbp contig against the 6.2 G-symbol index")):
    if not os.path.exists(kp):
        continue
    r = subprocess.run(["ncu", "-i", kp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(r.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, ms, sect = [], [], []
    for row in rows[2:]:
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m); b += _num(row[i]) * mult[units[i]]
        tot.append(b)
        i = hdr.index("gpu__time_duration.sum"); ms.append(_num(row[i]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[units[i]])
        if "dram__sectors_read.sum" in hdr:
            sect.append(_num(row[hdr.index("dram__sectors_read.sum")]))
    per = sum(tot) / len(tot)
    tj[wl] = {"k_seed_dram_bytes_per_launch": per, "k_seed_dram_bytes_per_query_bp": per / bp, "launches": len(tot),
              "k_seed_launch_ms_under_ncu": sum(ms) / len(ms), "dram_sectors_read": sum(sect) / len(sect) if sect else None,
              "source": f"profiles/{tag}_summary.md, {os.path.basename(kp)} (ncu --set full --clock-control none, {how})"}
json.dump(tj, open(tpath, "w"), indent=1)
for extra in ("dp_bench.json", "random_access.json"):
    if os.path.exists("gpurun_out/" + extra):
        out.append(f"\n## {extra}\n\n```json\n{open('gpurun_out/' + extra).read().strip()}\n```")
for path in sorted(glob.glob("gpurun_out/bench_*.json")):
    txt = open(path).read().strip()
    if txt:
        out.append(f"\n## {os.path.basename(path)}\n\n```json\n{txt}\n```")
# multi-GPU bench lines (strong scaling over the contig shards) and the A/B experiments of the round, as run
for path, title in (("gpurun_out/r2m_bench_C4_n4.json", "N = 4, plain records, earlier in the round (tools/gpu_r2_m.sh)"),
                    ("gpurun_out/r2q_bench_C4_n8_compact.json", "N = 8, compact records, before N4 / 6 lanes (tools/gpu_r2_q.sh)"), ("gpurun_out/r2q_bench_C4_n8_plain.json", "N = 8, plain records, same box as the previous line")):
    if os.path.exists(path) and open(path).read().strip():
        out.append(f"\n## multi-GPU: {title}\n\n```json\n{open(path).read().strip()}\n```")
for path, title in (("gpurun_out/r2h_l2fetch.txt", "K1 vs cudaLimitMaxL2FetchGranularity (tools/gpu_r2_h.sh): no effect"),
                    ("gpurun_out/r2t_seed_occ.txt", "K1 at 6 vs 8 CTAs per SM (tools/gpu_r2_t.sh): 8 is slower"),
                    ("gpurun_out/r2s_pcie.txt", "PCIe link of the box, idle and under an HBM-streaming kernel (tools/pcie_probe.py)"),
                    ("gpurun_out/r2_e2e_experiments.txt", "e2e leg: hardware work queues, bulk copies in pieces, small copies by kernel or DMA, one direction only (calls J, N, L)"),
                    ("gpurun_out/r2y_bg_copies.txt", "ONE device-resident C4 contig, nothing else on the GPU, while unrelated bulk copies keep the PCIe link busy (tools/prof_contig.py --bg)"),
                    ("gpurun_out/r2w_contig.txt", "chained-scan tile size: 1024 elements (committed) vs 2048 (rebuilt on the box): cluster 1.05 vs 1.23 ms; C5 contig in between"),
                    ("gpurun_out/r2_lanes_sweep.txt", "contigs in flight per GPU (bench.py --lanes), same box, final tree"),
                    ("gpurun_out/r2v_c3_cli.txt", "bin/GSAlign at C3 (1 Gbp x 1 Gbp), stages"), ("gpurun_out/r2u_contig.txt", "one C4 contig, phase times: block logic in the kernel (two lines) / on the host (last line)")):
    if os.path.exists(path):
        out.append(f"\n## {title}\n\n```\n{open(path).read().strip()}\n```")
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_summary.md")
