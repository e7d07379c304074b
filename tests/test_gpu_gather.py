"""The record gather (gather.cu) through the C ABI: outbox packing, the NCCL exchange and the host-side record walk."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records_equal(a, b):
    """blocks and fragments whole; the row pools over the fragments' own ranges (rows of DP fragments are right-aligned in
    their slots: the bytes in front of them are never written)"""
    (ba, fa, a1, a2), (bb, fb, b1, b2) = a, b
    if not (np.array_equal(ba, bb) and np.array_equal(fa, fb)):
        return False
    for f in fa[fa["bSeed"] == 0]:
        o, n = int(f["aln_off"]), int(f["aln_len"])
        if not (np.array_equal(a1[o:o + n], b1[o:o + n]) and np.array_equal(a2[o:o + n], b2[o:o + n])):
            return False
    return True


@pytest.mark.parametrize("form", ["compact", "plain"])
def test_outbox_roundtrip_one_rank(ecoli, form, monkeypatch):
    """one-rank communicator: two lanes append their contigs (device to device), the gather leaves the image where it is, and the
    records read back on the host equal what gsa_fill hands out directly -- in the compact form records travel in by default
    (8 bytes per fragment, expanded by gsa_record_frags) and in the plain one (GSA_GATHER_RAW=1)"""
    from gsalign_b200 import capi
    monkeypatch.setenv("GSA_GATHER_RAW", "1" if form == "plain" else "0")
    al = capi.Aligner(0)
    al.upload_index(ecoli["index"])
    lane = capi.Aligner(0, owner=al)
    q = np.frombuffer(ecoli["query"], dtype=np.uint8)
    pieces = [q[:1_500_000], q[1_500_000:3_200_000], q[3_200_000:]]
    direct = [al.align_contig(p) for p in pieces]
    al.comm_init_rank(capi.Aligner.comm_unique_id(), 0, 1)
    for step in range(2):                                # the second pass reuses the outbox after a reset
        for ln in (al, lane):
            ln.set_host_results(False)
        for i, p in enumerate(pieces):
            ln = (al, lane)[i & 1]
            ln.contig_begin(p); ln.seed(); ln.cluster(); ln.fill()
            al.outbox_append(ln, 10 + i)
        assert al.outbox_bytes() > 0
        raw_bytes = sum(len(d[1]) * capi.FRAG_DTYPE.itemsize for d in direct)
        assert (al.outbox_bytes() < raw_bytes) == (form == "compact")   # 40 bytes per fragment shrink to 8 (+ anchors)
        al.gather_records(0); al.gather_wait()
        recs = al.inbox_records(0)
        al.outbox_reset()
        assert sorted(r[0] for r in recs) == [10, 11, 12]
        for r in recs:
            assert _records_equal(r[1:], direct[r[0] - 10]), r[0]
    lane.close(); al.close()


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
from gsalign_b200 import bwaidx, capi, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
bi = bwaidx.load(sys.argv[2]); contigs = synth.read_fasta(sys.argv[3])
al = capi.Aligner(rank); al.upload_index(bi)
uid = [capi.Aligner.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
al.comm_init_rank(uid[0], rank, world)
al.set_host_results(False)
for i, (name, seq) in enumerate(contigs):
    if i % world == rank:
        al.contig_begin(seq); al.seed(); al.cluster(); al.fill(); al.outbox_append(al, i)
al.gather_records(0); al.gather_wait()
if rank == 0:
    got = {}
    for r in range(world):
        for rec in al.inbox_records(r):
            got[rec[0]] = rec[1:]
    al.set_host_results(True)
    assert sorted(got) == list(range(len(contigs))), sorted(got)
    for i, (name, seq) in enumerate(contigs):
        want = al.align_contig(seq)
        fa = want[1]
        assert np.array_equal(got[i][0], want[0]) and np.array_equal(got[i][1], fa), i
        for f in fa[fa["bSeed"] == 0]:
            o, n = int(f["aln_off"]), int(f["aln_len"])
            assert np.array_equal(got[i][2][o:o + n], want[2][o:o + n]) and np.array_equal(got[i][3][o:o + n], want[3][o:o + n]), i
    print("GATHER_OK", len(contigs))
dist.barrier(); al.close(); dist.destroy_process_group()
"""


def test_nccl_gather_two_ranks(workdir):
    """one process per GPU (torchrun): the records rank 0 unpacks from the gathered images are byte-equal to its own single-GPU
    results for every contig.  Needs 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import build_index
    from test_gpu_pipeline import make_rearranged
    d = make_rearranged(workdir)
    if not os.path.exists(os.path.join(d, "ref.sa")):
        build_index(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
    w = os.path.join(d, "gather_worker.py")
    open(w, "w").write(WORKER)
    n = min(4, torch.cuda.device_count())
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29577",
                        w, ROOT, os.path.join(d, "ref"), os.path.join(d, "qry.fa")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
