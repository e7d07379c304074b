#!/usr/bin/env python
"""K3 stress: DP-only batches through gsa_dp_batch() -> GCUPS and fraction of the measured packed-int16 DPX issue rate.

  python tools/bench_dp.py [--out profiles/r1_dp.json]

Batches: synthetic N x (LxL) pairs at 10 % divergence (C5-like fragments) for L = 16..2048 and a
C3-like mix.  Algorithmic work: 3 add-max per cell = 1.5 packed-s16x2 instructions per cell; the kernel's own count
(8 VIMNMX/VIADD.16x2 per 2 cells) is reported next to it.  A sample of every batch is checked against the oracle.
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gsalign_b200 import capi  # noqa: E402
from gsalign_b200.synth import make_dp_batch as make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cells", type=float, default=4e9, help="target cells per batch")
    ap.add_argument("--sizes", default="16,32,64,128,256,512,1024,2048")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    al = capi.Aligner(0)
    peaks = {n: al.dpx_peak(i) for i, n in enumerate(["VIADDMNMX.S16x2", "VIMNMX.S16x2(+LOP3)", "VIADD.16x2", "VIMNMX3.S16x2(+LOP3)"])}
    peak = peaks["VIADDMNMX.S16x2"]
    import orc
    O = orc.Oracle()
    rng = np.random.default_rng(5)
    rows = []
    for L in [int(x) for x in args.sizes.split(",")]:
        n_pairs = int(max(64, min(2_000_000, args.cells / (L * L))))
        rb, ro, qb, qo = make_batch(rng, n_pairs, L)
        cells = int(np.sum((ro[1:] - ro[:-1]) * (qo[1:] - qo[:-1])))
        al.dp_batch_arrays(rb, ro, qb, qo)  # warm-up (allocations)
        best = 1e30
        for _ in range(args.reps):
            o1, o2, ol, ms = al.dp_batch_arrays(rb, ro, qb, qo)
            best = min(best, ms)
        for i in list(range(0, n_pairs, max(1, n_pairs // 8)))[:8]:  # parity sample
            a = rb[ro[i]:ro[i + 1]].tobytes(); b = qb[qo[i]:qo[i + 1]].tobytes()
            off = int(ro[i] + qo[i]); Lx = int(ol[i])
            assert (o1[off:off + Lx].tobytes(), o2[off:off + Lx].tobytes()) == O.dp_align(a, b), (L, i)
        gcups = cells / (best * 1e-3) / 1e9
        rows.append({"L": L, "pairs": n_pairs, "cells": cells, "kernel_ms": best, "gcups": gcups,
                     "dpx_ginstr_algorithmic": 1.5 * gcups, "frac_of_dpx_peak_algorithmic": 1.5 * gcups / peak,
                     "dpx_ginstr_executed": 4.0 * gcups, "frac_of_dpx_peak_executed": 4.0 * gcups / peak})
        print(json.dumps(rows[-1]), flush=True)
    res = {"dpx_peak_ginstr_per_s": peaks, "peak_used": "VIADDMNMX.S16x2", "batches": rows,
           "note": "algorithmic = 1.5 packed s16x2 add-max per cell (SURVEY 8d); executed = 8 VIMNMX/VIADD.16x2 per lane-step of 2 cells"}
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    al.close()


if __name__ == "__main__":
    main()
