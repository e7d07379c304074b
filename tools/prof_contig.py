#!/usr/bin/env python
"""Profiling driver: the hot path over ONE contig of a bench workload, device-resident, repeated; run it under ncu.
  python tools/prof_contig.py --workload C4 --contig 0 --reps 2
prints the library's per-phase CUDA-event times of the last repetition and the launch count of one contig."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gsalign_b200 import bwaidx, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C4")
ap.add_argument("--contig", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--bg", default="none", choices=["none", "h2d", "d2h", "both"], help="bulk PCIe copies of unrelated pinned buffers running on other streams "
                "while the contig is processed: what the phases of ONE contig lose to DMA traffic on the link")
args = ap.parse_args()
import torch  # noqa: E402
w = bench.WORKLOADS[args.workload]
d, prefix, qry_fa = bench.prepare_workload(args.workload)
contigs = synth.read_fasta_fast(qry_fa, max_bp=(args.contig + 1) * (w["n"] // w["k"]))
seq = contigs[args.contig][1]
al = capi.Aligner(0)
al.set_params(**w["prm"])
al.upload_index(bwaidx.load(prefix))
al.set_host_results(False)
dev = torch.from_numpy(np.ascontiguousarray(seq)).cuda()
import threading, time  # noqa: E402
stop = False
def bg_copies():
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    while not stop:
        if args.bg in ("h2d", "both"):
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if args.bg in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
        s1.synchronize(); s2.synchronize()
th = None
if args.bg != "none":
    th = threading.Thread(target=bg_copies); th.start(); time.sleep(0.5)
for r in range(args.reps):
    al.contig_begin_device(dev.data_ptr(), dev.shape[0]); al.seed(); al.cluster()
    al._chk(al.lib.gsa_fill(al.ctx, ctypes.byref(capi.Alignment())))
    t = al.timing()
    print(f"rep {r}: seed {t.seed_ms:.3f} (k_seed {t.k_seed_ms:.3f}) cluster {t.cluster_ms:.3f} fill {t.fill_ms:.3f} (k_dp {t.k_dp_ms:.3f}) total {t.total_ms:.3f} ms; "
          f"{t.n_seeds} seeds, {t.n_dp} DP problems, {t.dp_cells} cells, {t.launches} launches", flush=True)
stop = True
if th: th.join()
al.close()
