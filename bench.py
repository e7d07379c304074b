#!/usr/bin/env python
"""bench.py -- query Gbp/s of the seed -> cluster/chain -> gapped-fill path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]

A "step" is one pass of the hot path over every contig of the synthetic query genome of the workload
(default C2: 100 Mbp reference vs 100 Mbp query, 4 contigs, 1 % SNV, reference defaults).
  value     whole-job throughput with the query already resident in HBM (gsa_contig_begin_device ->
            gsa_seed -> gsa_cluster -> gsa_fill), CUDA events on the launching stream, max over ranks
  e2e       the same metric through the reference-facing call gsa_align_contig() on pinned HOST buffers:
            H2D of the contig and D2H of the alignment records are inside the timed region (wall clock
            between device synchronisations)
  roofline  the dominant kernel (K1 fm_seed): algorithmic bytes of the REFERENCE's algorithm (B_seed,
            SURVEY.md 8d, counted by the oracle on a bounded sample of the same input) / the kernel's
            CUDA-event duration / the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference binary (oracle/_ref/GSAlign -t nproc) on a bounded sample
N > 1: weak scaling -- every rank aligns its own copy of the workload against an index replicated in its
HBM (query contigs are independent); once per step the finished alignment records of all ranks
are gathered to rank 0 over NCCL (gsalign_b200/gather.py), the only collective of the path.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gsalign_b200 import bwaidx, synth  # noqa: E402

CACHE = os.environ.get("GSA_BENCH_CACHE", "/tmp/gsa_bench_cache")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")

# workload -> (N, K, p_snv, p_indel, seed, extra CLI flags / params)
WORKLOADS = {
    "C2": dict(n=100_000_000, k=4, snv=0.01, indel=0.0, seed=2, prm={}, flags=[]),
    "C2s": dict(n=20_000_000, k=4, snv=0.01, indel=0.0, seed=2, prm={}, flags=[]),       # dev-size C2
    "C3s": dict(n=20_000_000, k=4, snv=0.02, indel=0.002, seed=3, prm={}, flags=[]),     # dev-size C3
    "C3": dict(n=1_000_000_000, k=8, snv=0.02, indel=0.002, seed=3, prm={}, flags=[]),
    "C5": dict(n=500_000_000, k=4, snv=0.10, indel=0.0, seed=5, prm=dict(min_seed_len=10, sensitive=1, min_block_score=50, min_idy=70),
               flags=["-sen", "-slen", "10", "-idy", "70"]),
    "C5s": dict(n=10_000_000, k=4, snv=0.10, indel=0.0, seed=5, prm=dict(min_seed_len=10, sensitive=1, min_block_score=50, min_idy=70),
                flags=["-sen", "-slen", "10", "-idy", "70"]),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def prepare_workload(name: str):
    """Generates the pair, writes FASTA, builds (or reuses) the BWA-format index.  Untimed setup."""
    w = WORKLOADS[name]
    d = os.path.join(CACHE, name)
    os.makedirs(d, exist_ok=True)
    ref_fa, qry_fa, prefix = os.path.join(d, "ref.fa"), os.path.join(d, "qry.fa"), os.path.join(d, "ref")
    if not (os.path.exists(prefix + ".sa") and os.path.exists(qry_fa) and os.path.exists(os.path.join(d, "ok"))):
        t0 = time.time()
        ref, qry = synth.make_pair(w["n"], w["k"], w["snv"], w["indel"], w["seed"])
        synth.write_fasta(ref_fa, ref); synth.write_fasta(qry_fa, qry)
        log(f"generated {name} in {time.time() - t0:.1f}s")
        t0 = time.time()
        build_index(ref_fa, prefix)
        log(f"index built in {time.time() - t0:.1f}s")
        open(os.path.join(d, "ok"), "w").write("ok")
    return d, prefix, qry_fa


def build_index(ref_fa: str, prefix: str):
    """Index producer (offline step, not part of the metric): our GPU builder when present, else the
    reference's single-threaded indexer."""
    gpu_builder = os.path.join(ROOT, "bin", "gsa_index")
    if os.path.exists(gpu_builder):
        r = subprocess.run([gpu_builder, ref_fa, prefix], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if r.returncode == 0:
            return
        log("gsa_index failed, falling back to the reference indexer:", r.stderr.decode()[-300:])
    subprocess.run([os.path.join(REF_BIN, "bwt_index"), ref_fa, prefix], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class ClockSampler(threading.Thread):
    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag = gpu, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.25)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference_cli(prefix: str, qry_fa: str, out_prefix: str, threads: int, flags):
    """Runs the unmodified reference binary; returns (seconds of GenomeComparison-equivalent, total seconds).
    The phase start is the 'Step2.' stderr line (index + query load are excluded, emit included)."""
    exe = os.path.join(REF_BIN, "GSAlign")
    t0 = time.perf_counter()
    p = subprocess.Popen([exe, "-t", str(threads), "-i", prefix, "-q", qry_fa, "-o", out_prefix] + list(flags),
                         stderr=subprocess.PIPE, stdout=subprocess.DEVNULL)
    t_step2 = None
    buf = b""
    while True:
        chunk = p.stderr.read1(65536) if hasattr(p.stderr, "read1") else p.stderr.read(65536)
        if not chunk:
            break
        if t_step2 is None:
            buf += chunk
            if b"Step2." in buf:
                t_step2 = time.perf_counter(); buf = b""
    p.wait()
    t1 = time.perf_counter()
    return (t1 - (t_step2 if t_step2 is not None else t0)), (t1 - t0)


def sample_query(qry_fa: str, d: str, max_bp: int):
    """bounded sample of the workload for the CPU legs: whole contigs until max_bp"""
    contigs = synth.read_fasta(qry_fa)
    take, tot = [], 0
    for name, seq in contigs:
        if tot and tot + seq.shape[0] > max_bp:
            break
        take.append((name, seq)); tot += seq.shape[0]
    path = os.path.join(d, f"sample_{tot}.fa")
    if not os.path.exists(path):
        synth.write_fasta(path, take)
    return path, tot, len(take)


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    d, prefix, qry_fa = prepare_workload(args.workload)
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample_fa, bp, nct = sample_query(qry_fa, d, args.cpu_sample_bp)
    times = []
    for i in range(args.warmup + args.steps):
        t, _ = run_reference_cli(prefix, sample_fa, os.path.join(d, "ref_out"), cores, w["flags"])
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = bp / sec / 1e9
    line = {"impl": "reference", "metric": "query Gbp/sec", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8/int64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['n'] // 1_000_000} Mbp x {w['n'] // 1_000_000} Mbp, {w['k']} contigs, SNV {w['snv']}, indel {w['indel']}",
                       "sample": f"first {nct} query contig(s) = {bp} bp against the full index"},
            "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                             "sample": f"oracle/_ref/GSAlign -t {cores}, {bp} query bp, GenomeComparison phase incl. MAF/VCF emit"},
            "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-bp", type=int, default=50_000_000)
    ap.add_argument("--oracle-sample-bp", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dp-stress", action="store_true", help="skip the K3 stress batch (roofline_k3), e.g. under ncu")
    ap.add_argument("--lanes", type=int, default=0, help="contigs in flight per GPU (contexts sharing one index, one host thread each); "
                                                         "0 = min(4, host cores / ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from gsalign_b200 import capi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOADS[args.workload]

    # ---- untimed setup: data, index (rank 0 builds, everybody loads), upload -------------------------------
    if rank == 0:
        d, prefix, qry_fa = prepare_workload(args.workload)
    if world > 1:
        dist.barrier()
    d, prefix, qry_fa = prepare_workload(args.workload)
    t0 = time.time()
    bi = bwaidx.load(prefix)
    contigs = synth.read_fasta(qry_fa)
    total_bp = sum(s.shape[0] for _, s in contigs)
    al = capi.Aligner(local)
    al.set_params(**w["prm"])
    al.upload_index(bi)
    # lanes: contexts on this GPU sharing the uploaded index, one host thread and one stream each; query contigs are
    # independent, so the lanes keep the GPU busy across each other's host-side steps
    want_lanes = args.lanes if args.lanes > 0 else max(1, min(4, (os.cpu_count() or 4) // max(1, world)))
    n_lanes = max(1, min(want_lanes, len(contigs)))
    lanes = [al] + [capi.Aligner(local, owner=al) for _ in range(n_lanes - 1)]
    streams = [torch.cuda.Stream() for _ in lanes]
    for ln, st in zip(lanes, streams):
        ln.set_stream(st.cuda_stream)
    log(f"rank {rank}: index loaded + uploaded in {time.time() - t0:.1f}s; query {total_bp} bp in {len(contigs)} contigs; {n_lanes} lanes")
    dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda(non_blocking=False) for _, s in contigs]   # HBM-resident inputs
    pinned = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for _, s in contigs]               # pinned host inputs
    pinned_np = [p.numpy() for p in pinned]
    order = sorted(range(len(contigs)), key=lambda i: -contigs[i][1].shape[0])                          # longest first

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=n_lanes)

    def run_lanes(fn):
        """every lane pulls contig indices (longest first) until none is left; returns the per-contig results"""
        it = iter(order)
        lock = threading.Lock()
        out = [None] * len(contigs)

        def worker(k):
            while True:
                with lock:
                    i = next(it, None)
                if i is None:
                    return
                out[i] = fn(lanes[k], i)
        futs = [pool.submit(worker, k) for k in range(n_lanes)]
        for f in futs:
            f.result()
        return out

    # N > 1: the one collective of the path -- every rank packs its finished records (device memory) into an outbox and
    # rank 0 gathers them over NCCL once per step (gsalign_b200/gather.py); inside the timed region of `value`
    from gsalign_b200 import gather
    outbox = {"box": None, "boxes": None, "turn": 0, "free": [None, None], "lock": threading.Lock(), "bytes": 0}

    class _DevView:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def _dev_u8(ptr, nbytes):
        return torch.as_tensor(_DevView(ptr, nbytes), device=f"cuda:{local}") if nbytes else torch.empty(0, dtype=torch.uint8, device=f"cuda:{local}")

    def contig_device(ln, i):
        t = dev[i]
        ln.contig_begin_device(t.data_ptr(), t.shape[0]); ln.seed(); ln.cluster()
        ln._chk(ln.lib.gsa_fill(ln.ctx, ctypes.byref(capi.Alignment())))
        if world > 1:
            r = ln.result_device()
            nb = gather.record_bytes(r.n_blocks, r.n_frags, r.aln_bytes)
            with outbox["lock"]:
                outbox["bytes"] += nb
                off = outbox["box"].reserve(nb) if outbox["box"] is not None else None
            if off is not None:
                nbb = r.n_blocks * gather.BLOCK_BYTES
                blocks = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(r.blocks, ctypes.POINTER(ctypes.c_uint8)), shape=(nbb,)).copy()) if nbb else torch.empty(0, dtype=torch.uint8)
                with torch.cuda.stream(streams[lanes.index(ln)]):
                    outbox["box"].put(off, i, blocks, _dev_u8(r.frags, r.n_frags * gather.FRAG_BYTES), _dev_u8(r.aln1, r.aln_bytes), _dev_u8(r.aln2, r.aln_bytes))
        return ln.timing()

    def contig_device_plain(ln, i):
        t = dev[i]
        ln.contig_begin_device(t.data_ptr(), t.shape[0]); ln.seed(); ln.cluster()
        ln._chk(ln.lib.gsa_fill(ln.ctx, ctypes.byref(capi.Alignment())))
        return ln.timing()

    def gather_step():
        """after every lane finished its contigs: one NCCL gather of this step's records to rank 0.  Two outboxes alternate,
        so the gather of step k runs on the master stream while the lanes already fill the other box with step k+1."""
        if world == 1 or os.environ.get("GSA_BENCH_NO_GATHER"):   # the switch is a diagnosis aid
            if outbox["box"] is not None:
                outbox["box"].reset()
            return 0
        box = outbox["box"]
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); master.wait_event(ev)
        with torch.cuda.stream(master):
            inbox = gather.gather_to_root(box.buf, box.used)
        got = sum(int(b.numel()) for b in inbox) if inbox is not None else 0
        done = torch.cuda.Event(); done.record(master)
        t = outbox["turn"]
        outbox["free"][t] = done                   # this box may be refilled once its gather has run
        t ^= 1
        outbox["turn"] = t; outbox["box"] = outbox["boxes"][t]
        outbox["box"].reset()
        if outbox["free"][t] is not None:
            for st in streams:
                st.wait_event(outbox["free"][t])
        return got

    def contig_host(ln, i):
        r = ln.align_contig_raw(pinned_np[i])
        return r.n_frags * capi.FRAG_DTYPE.itemsize + 2 * r.aln_bytes + r.n_blocks * capi.BLOCK_DTYPE.itemsize

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs and results (block headers come to the host; fragments and rows stay in HBM, where
    # the record gather of N > 1 reads them); CUDA events: every lane's stream starts after e0 and e1 follows all of them ---
    for ln in lanes:
        ln.set_host_results(False)
    master = torch.cuda.Stream(priority=-1)        # the gather's NCCL kernels must not queue behind the lanes' compute kernels
    run_lanes(contig_device)                       # sizes the outbox (N > 1) and warms the allocators
    if world > 1:
        outbox["boxes"] = [gather.Outbox(int(outbox["bytes"] * 1.1) + (1 << 20), torch.device("cuda", local)) for _ in range(2)]
        outbox["box"] = outbox["boxes"][0]
    sampler = ClockSampler(local)
    if rank == 0:   # one sampler per box (every nvidia-smi call takes driver locks the launching threads also need); it runs
        sampler.start()   # from the warm-up on: the timed regions are a few tens of ms, one nvidia-smi call takes about as long
    for _ in range(args.warmup):
        run_lanes(contig_device)
        gather_step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_seed_ms, k_dp_ms, launches, seed_ms, cluster_ms, fill_ms, dp_cells, n_seeds = [], [], 0, 0.0, 0.0, 0.0, 0, 0
    gathered = 0
    e0.record(master)
    for st in streams:
        st.wait_event(e0)
    for _ in range(args.steps):
        for t in run_lanes(contig_device):
            k_seed_ms.append(t.k_seed_ms); k_dp_ms.append(t.k_dp_ms); launches += t.launches
            seed_ms += t.seed_ms; cluster_ms += t.cluster_ms; fill_ms += t.fill_ms; dp_cells += t.dp_cells; n_seeds += t.n_seeds
        gathered = gather_step()
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        master.wait_event(ev)
    e1.record(master)
    sync_all()
    dev_ms = e0.elapsed_time(e1)

    # ---- roofline pass: the dominant kernel timed alone (one lane, nothing else on the GPU), CUDA events on its stream ------
    k_seed_alone = []
    for _ in range(max(1, min(args.steps, 3))):
        for i in order:
            k_seed_alone.append((contig_device_plain(lanes[0], i).k_seed_ms, contigs[i][1].shape[0]))
    for ln in lanes:
        ln.set_host_results(True)

    # ---- e2e: host buffers through gsa_align_contig, wall clock between synchronisations -------------------------
    for _ in range(args.warmup):
        run_lanes(contig_host)
    sync_all()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        d2h = sum(run_lanes(contig_host))
    sync_all()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
    pool.shutdown()

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tl)
        launches = int(tl[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_per_step = dev_ms / args.steps
    value = world * total_bp / (ms_per_step * 1e-3) / 1e9
    e2e_val = world * total_bp / (e2e_s / args.steps) / 1e9

    # ---- roofline of the dominant kernel (K1): B_seed counted by the oracle on a bounded sample ---------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    import orc
    O = orc.Oracle()
    oix = O.index(bi)
    ctr = orc.OrcCounters()
    samp = contigs[0][1][: args.oracle_sample_bp].tobytes()
    prm = orc.params(**{k: v for k, v in w["prm"].items() if k in ("min_seed_len", "sensitive")})
    sq, _, _ = O.seed_contig(oix, prm, samp, ctr)
    bseed_per_bp = ctr.algorithmic_bytes(len(samp), len(sq)) / len(samp)
    mean_k_seed_ms = float(np.mean([t for t, _ in k_seed_alone]))
    bp_per_launch = float(np.mean([b for _, b in k_seed_alone]))
    achieved = bseed_per_bp * bp_per_launch / (mean_k_seed_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read+write per k_seed launch from the committed ncu --set full capture of this workload
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get("k_seed_dram_bytes_per_launch")
    roofline = {"kernel": "k_seed (K1 fm_seed)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "note": f"achieved = B_seed of the reference's algorithm ({bseed_per_bp:.1f} B/query bp, oracle counters on the first {len(samp)} bp) "
                        f"x {bp_per_launch:.0f} bp per launch / {mean_k_seed_ms:.3f} ms mean launch (timed alone on its stream after the timed region; "
                        f"inside it, with {n_lanes} contigs in flight, launches overlap and average {float(np.mean(k_seed_ms)):.3f} ms); an EFFECTIVE fraction: this kernel replaces "
                        "per-base rank walks by a prefix table + full SA + text compare, actual DRAM bytes are in profiles/"}

    # ---- K3 roofline: DP-only stress batch against the packed-int16 DPX issue rate measured on this box -----
    roofline_k3 = None
    try:
        if args.no_dp_stress:
            raise RuntimeError("skipped (--no-dp-stress)")
        dpx_peak = al.dpx_peak(0)
        rb, ro, qb, qo = synth.make_dp_batch(np.random.default_rng(5), 1024, 1024)
        cells = int(np.sum((ro[1:] - ro[:-1]) * (qo[1:] - qo[:-1])))
        al.dp_batch_arrays(rb, ro, qb, qo)
        dp_ms = min(al.dp_batch_arrays(rb, ro, qb, qo)[3] for _ in range(3))
        gcups = cells / (dp_ms * 1e-3) / 1e9
        roofline_k3 = {"kernel": "k_dpx (K3 gapped fill)", "bound": "dpx", "achieved": 1.5 * gcups, "peak": dpx_peak, "unit": "G s16x2-instr/s",
                       "frac": 1.5 * gcups / dpx_peak, "gcups": gcups, "executed_frac": 4.0 * gcups / dpx_peak,
                       "note": f"stress batch of 1024 pairs ~1024x1024 at 10 % divergence ({cells} cells, {dp_ms:.3f} ms, gsa_dp_batch kernel span); "
                               "achieved = 1.5 packed add-max per cell (SURVEY 8d) x GCUPS; peak = VIADDMNMX.S16x2 issue rate measured by "
                               "gsa_dpx_peak() on this GPU; executed_frac counts the 8 VIMNMX/VIADD.16x2 the kernel issues per 2 cells"}
    except Exception as e:  # noqa: BLE001
        log("K3 stress batch failed:", e)

    # ---- CPU baseline: the unmodified reference on a bounded sample ---------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and os.path.exists(os.path.join(REF_BIN, "GSAlign")):
        cores = os.cpu_count() or 1
        sample_fa, bp, nct = sample_query(qry_fa, d, args.cpu_sample_bp)
        sec, tot = run_reference_cli(prefix, sample_fa, os.path.join(d, "cpu_out"), cores, w["flags"])
        cpu = {"value": bp / sec / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
               "sample": f"oracle/_ref/GSAlign -t {cores} on the first {nct} query contig(s) ({bp} bp) vs the full index; "
                         f"GenomeComparison phase {sec:.2f}s (whole process {tot:.2f}s)"}

    line = {"metric": "query Gbp/sec", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16/int64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['n'] // 1_000_000} Mbp ref x {total_bp} bp query in {len(contigs)} contigs, SNV {w['snv']}, indel {w['indel']}, "
                                   f"params {w['prm'] or 'reference defaults'}",
                       "parallelism": f"{world} x (full index replica + own query copy), {n_lanes} contigs in flight per GPU; "
                                      + ("no collective (1 GPU)" if world == 1 else f"one NCCL record gather to rank 0 per step ({gathered} bytes) inside the timed region of value"),
                       "host_cores": os.cpu_count(),
                       "l2": "inputs larger than L2 (index " + f"{(bi.seq_len * 4 + bi.seq_len // 2 + bi.seq_len // 4) / 1e6:.0f} MB resident, query {total_bp / 1e6:.0f} MB); no flush needed"},
            "e2e": {"value": e2e_val, "unit": "Gbp/s", "h2d_bytes_per_step": int(total_bp), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "phases_ms_per_step_note": "per-contig CUDA-event spans summed over contigs; lanes overlap, so they add up to more than ms_per_step",
            "phases_ms_per_step": {"seed": seed_ms / args.steps, "cluster": cluster_ms / args.steps, "fill": fill_ms / args.steps,
                                   "k_seed": float(np.sum(k_seed_ms)) / args.steps, "k_dp": float(np.sum(k_dp_ms)) / args.steps},
            "counts_per_step": {"seeds": n_seeds // args.steps, "dp_cells": dp_cells // args.steps},
            "roofline": roofline, "roofline_k3": roofline_k3, "cpu_baseline": cpu, "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
