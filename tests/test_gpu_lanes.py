"""Lanes (gsa_create_shared), device-resident results (gsa_set_host_results / gsa_result_device) and degenerate contigs,
through the C ABI, against a single default context and the oracle."""
import ctypes
import os
import threading

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _make(workdir, n=600_000, k=3, seed=41):
    from conftest import build_index
    from gsalign_b200 import bwaidx, synth
    d = os.path.join(workdir, f"lanes_{n}_{seed}")
    os.makedirs(d, exist_ok=True)
    ref, qry = synth.make_pair(n, k, 0.02, 0.002, seed)
    synth.write_fasta(os.path.join(d, "ref.fa"), ref)
    build_index(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
    return bwaidx.load(os.path.join(d, "ref")), qry


def _result(al, seq):
    """block headers, fragment records and the rows the fragments point at (the pools have undefined holes: DP rows are
    right-aligned inside their slots)"""
    blocks, frags, a1, a2 = al.align_contig(seq)
    rows = [(a1[int(f["aln_off"]):int(f["aln_off"]) + int(f["aln_len"])].tobytes(), a2[int(f["aln_off"]):int(f["aln_off"]) + int(f["aln_len"])].tobytes())
            for f in frags if not f["bSeed"]]
    return blocks.tobytes(), frags.tobytes(), rows


def test_lanes_match_single_context(workdir):
    """three lanes sharing one index, driven concurrently by three host threads, give the records of one context"""
    from gsalign_b200 import capi
    bi, qry = _make(workdir)
    owner = capi.Aligner(0)
    owner.upload_index(bi)
    want = [_result(owner, s.tobytes()) for _, s in qry]
    lanes = [owner, capi.Aligner(0, owner=owner), capi.Aligner(0, owner=owner)]
    got = [None] * len(qry)
    errs = []

    def work(k):
        try:
            for rep in range(3):       # several rounds so that the lanes really overlap
                got[k] = _result(lanes[k], qry[k][1].tobytes())
        except Exception as e:         # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=work, args=(k,)) for k in range(len(qry))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    # frag records hold offsets into the row pools, which are laid out identically: byte equality is meaningful
    assert got == want
    for ln in lanes[1:]:
        ln.close()
    owner.close()


def test_device_resident_results_equal_host_results(workdir):
    import torch
    from gsalign_b200 import capi
    bi, qry = _make(workdir)
    al = capi.Aligner(0)
    al.upload_index(bi)
    seq = qry[0][1].tobytes()
    blocks, frags, a1, a2 = al.align_contig(seq)
    al.set_host_results(False)
    out = capi.Alignment()
    al.contig_begin(seq); al.seed(); al.cluster()
    al._chk(al.lib.gsa_fill(al.ctx, ctypes.byref(out)))
    assert out.n_blocks == len(blocks) and out.n_frags == len(frags) and not out.frags and not out.aln1
    r = al.result_device()

    class View:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    d_frags = torch.as_tensor(View(r.frags, r.n_frags * capi.FRAG_DTYPE.itemsize), device="cuda:0").cpu().numpy()
    assert d_frags.tobytes() == frags.tobytes()
    for ptr, host in ((r.aln1, a1), (r.aln2, a2)):
        dev = torch.as_tensor(View(ptr, r.aln_bytes), device="cuda:0").cpu().numpy()
        # only the bytes the fragments point at are defined (DP rows are right-aligned inside their slots)
        for f in frags:
            if not f["bSeed"]:
                o, n = int(f["aln_off"]), int(f["aln_len"])
                assert dev[o:o + n].tobytes() == host[o:o + n].tobytes()
    al.set_host_results(True)
    al.close()


@pytest.mark.parametrize("seq", [b"", b"A", b"ACGTACGTAC", b"N" * 5000, b"ACGT" * 10 + b"N" * 40 + b"ACGT" * 3,
                                 bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(7).integers(0, 4, 30_000)])])
def test_degenerate_contigs(workdir, seq, oracle):
    """empty / shorter than a seed / all-N / unrelated contigs: no error, and seeds + blocks equal the oracle's (none)"""
    from gsalign_b200 import capi
    bi, _ = _make(workdir)
    al = capi.Aligner(0)
    al.upload_index(bi)
    ix = oracle.index(bi)
    al.contig_begin(seq)
    n = al.seed()
    q, r, l = al.fetch_seeds(n)
    oq, orr, ol = oracle.seed_contig(ix, orc.params(), seq)
    assert np.array_equal(q, oq) and np.array_equal(r, orr) and np.array_equal(l, ol)
    al.cluster()
    blocks, frags, a1, a2 = al.fill()
    want = oracle.cluster(ix, orc.params(), seq, oq, orr, ol, 2) if len(oq) else []
    assert len(blocks) == len(want)
    al.close()


def test_prefetched_upload_same_results(workdir):
    """gsa_contig_prefetch: a contig whose upload was started ahead (two may be pending) gives the records of the plain call;
    an announced contig that is not the next one simply waits for its turn; a third pending upload is refused."""
    import torch
    from gsalign_b200 import capi
    bi, qry = _make(workdir)
    al = capi.Aligner(0)
    al.upload_index(bi)
    seqs = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory().numpy() for _, s in qry]
    want = [_result(al, s) for s in seqs]
    for rep in range(2):
        # the bench's pattern: announce the next contig, then work on the current one
        al.prefetch(seqs[0])
        got = []
        for k in range(len(seqs)):
            if k + 1 < len(seqs):
                al.prefetch(seqs[k + 1])
            got.append(_result(al, seqs[k]))
        assert got == want
    # out of order: 2 is announced, 0 and 1 run first without an announcement of their own
    al.prefetch(seqs[2])
    assert _result(al, seqs[0]) == want[0]
    al.prefetch(seqs[1])
    with pytest.raises(capi.GsaError):
        al.prefetch(seqs[0])                       # two uploads are pending already
    assert _result(al, seqs[1]) == want[1] and _result(al, seqs[2]) == want[2]
    al.close()
