"""dev tool: per-phase cycle / trip counters of k_seed (library built with -DSEED_PROFILE)"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from gsalign_b200 import bwaidx, capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
d, prefix, qry_fa = bench.prepare_workload(wl)
bi = bwaidx.load(prefix)
contigs = synth.read_fasta(qry_fa)
al = capi.Aligner(0)
al.set_params(**bench.WORKLOADS[wl]["prm"])
al.upload_index(bi)
buf = (ctypes.c_ulonglong * 16)()
names = ["tripsA", "tripsB", "activeA", "activeB", "cycA", "cycResolve", "cycB", "repairs", "warps", "searchesA", "searchesB"]
for rep in range(2):
    s = contigs[0][1].tobytes()
    al.contig_begin(s); al.lib.gsa_seed_profile(buf, 1)
    n = al.seed()
    al.lib.gsa_seed_profile(buf, 0)
    v = dict(zip(names, list(buf)))
    w = max(1, v["warps"])
    print(f"{wl} contig0 {len(s)} bp seeds {n} k_seed_ms {al.timing().k_seed_ms if False else ''}")
    print({k: round(v[k] / w, 1) for k in names}, "active/tripA", round(v["activeA"] / max(1, v["tripsA"]), 2), "active/tripB", round(v["activeB"] / max(1, v["tripsB"]), 2),
          "cyc/tripA", round(v["cycA"] / max(1, v["tripsA"])), "cyc/tripB", round(v["cycB"] / max(1, v["tripsB"])))
