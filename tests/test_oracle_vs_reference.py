"""Pins the oracle (oracle/gsa_oracle.c): (1) against golden vectors generated from the UNMODIFIED reference
(tests/golden/ecoli_golden.json, made by tests/golden/make_golden.py) -- runs anywhere; (2) directly against the
reference compiled in oracle/_ref/libgsref.so where it exists (fresh process per index: the reference uses globals)."""
import hashlib
import json
import os
import pickle
import random
import subprocess
import sys

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ecoli_golden.json")))


def digest_blocks(blocks):
    h = hashlib.sha256()
    for b in sorted(blocks, key=lambda b: (b[0], b[3])):
        h.update(repr((b[0], b[3])).encode())
    return h.hexdigest()


@pytest.fixture(scope="module")
def ecoli_oracle(ecoli, oracle):
    ix = oracle.index(ecoli["index"])
    ctr = orc.OrcCounters()
    q, r, l = oracle.seed_contig(ix, orc.params(), ecoli["query"], ctr)
    return ix, (q, r, l), ctr


def test_golden_seeds(ecoli_oracle):
    _, (q, r, l), ctr = ecoli_oracle
    assert len(q) == GOLD["n_seeds"] == 102812
    assert hashlib.sha256(q.astype("<i4").tobytes() + r.astype("<i8").tobytes() + l.astype("<i4").tobytes()).hexdigest() == GOLD["seeds_sha256"]
    # traffic counters of the reference's algorithm on C1 (0.989 / 0.188 / 0.682 / 0.0222 per query bp)
    assert (ctr.n_search, ctr.n_seedhit, ctr.n_sa_reads) == (98808, 93825, 102812)
    assert ctr.n_ext_steps == 4590150 and ctr.n_split == 871833 and ctr.n_lf_steps == 3166327


def test_golden_search_vectors(ecoli, oracle, ecoli_oracle):
    ix = ecoli_oracle[0]
    for start, stop, ln, fq, loc in GOLD["search_vectors"]:
        l2, f2, loc2 = oracle.bwt_search(ix, ecoli["query"], start, stop)
        assert (l2, f2, sorted(loc2)) == (ln, fq, loc)


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_golden_cluster_stages(ecoli, oracle, ecoli_oracle, stage):
    ix, (q, r, l), _ = ecoli_oracle
    blocks = oracle.cluster(ix, orc.params(), ecoli["query"], q, r, l, stage)
    assert len(blocks) == GOLD["stage_blocks"][str(stage)]
    assert digest_blocks(blocks) == GOLD["stage_sha256"][str(stage)]


def test_golden_fill(ecoli, oracle, ecoli_oracle):
    """stage 2 -> (dedup leaves the 1 big block) -> normal pairs -> fragment alignment == the reference's rows and sums"""
    ix, (q, r, l), _ = ecoli_oracle
    blocks = oracle.cluster(ix, orc.params(), ecoli["query"], q, r, l, 2)
    big = max(blocks, key=lambda b: b[0])
    frags = oracle.normal_pairs(big[3])
    score = aln_len = 0
    rows = []
    for f in frags:
        if f[0]:
            score += f[3]; aln_len += f[3]
            continue
        a1, a2, s, _ = oracle.frag_align(ix, ecoli["query"], f[1], f[2], f[3], f[4])
        rows += [a1, a2]; score += s; aln_len += len(a1)
    assert [score, aln_len, 0, len(frags)] == GOLD["final"][0]
    assert len(rows) == GOLD["n_aln"] and hashlib.sha256(b"\n".join(rows)).hexdigest() == GOLD["aln_sha256"]


def test_golden_dp_vectors(oracle):
    for a, b, x, y in GOLD["dp_vectors"]:
        assert oracle.dp_align(a.encode(), b.encode()) == (x.encode(), y.encode())


def test_dp_edge_cases(oracle):
    assert oracle.dp_align(b"A", b"A") == (b"A", b"A")
    assert oracle.dp_align(b"A", b"C") == (b"A", b"C")
    assert oracle.dp_align(b"ACGT", b"A") == (b"ACGT", b"A---")                # leftover columns become one gap
    x, y = oracle.dp_align(b"AAAA", b"AAAATT")
    assert x.replace(b"-", b"") == b"AAAA" and y.replace(b"-", b"") == b"AAAATT" and len(x) == len(y)
    x, y = oracle.dp_align(b"acgtNNacgt", b"ACGTACGT")                          # case kept, N scores 0
    assert x.replace(b"-", b"") == b"acgtNNacgt" and y.replace(b"-", b"") == b"ACGTACGT"


# ---- direct comparison with the compiled reference (only where oracle/_ref exists) -------------------------------
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(orc.REF_DIR, "libgsref.so")), reason="oracle/_ref not built")


@needs_ref
def test_dp_random_vs_reference(ecoli, oracle):
    code = r"""
import sys, random, pickle
sys.path.insert(0, %r); sys.path.insert(0, %r)
import orc
R = orc.Reference(%r)
random.seed(9)
out = []
for t in range(1500):
    m = random.randint(1, 90)
    a = ''.join(random.choice('ACGTacgtN') if random.random() < 0.1 else random.choice('ACGT') for _ in range(m))
    if t %% 3:
        b = ''.join((random.choice('ACGT') if random.random() < 0.1 else ('' if random.random() < 0.04 else c)) for c in a) or 'A'
    else:
        b = ''.join(random.choice('ACGT') for _ in range(random.randint(1, 90)))
    out.append((a, b) + R.ksw2(a.encode(), b.encode()))
pickle.dump(out, open(sys.argv[1], 'wb'))
""" % (HERE, os.path.dirname(HERE), ecoli["prefix"])
    out = os.path.join(ecoli["dir"], "dp_ref.pkl")
    subprocess.run([sys.executable, "-c", code, out], check=True)
    for a, b, x, y in pickle.load(open(out, "rb")):
        assert oracle.dp_align(a.encode(), b.encode()) == (x, y), (a, b)


@needs_ref
@pytest.mark.parametrize("prm", [dict(), dict(min_seed_len=10, sensitive=1, min_block_score=50)])
def test_rearranged_all_seams_vs_reference(workdir, oracle, prm):
    from conftest import build_index
    from gsalign_b200 import bwaidx, synth
    from test_gpu_pipeline import make_rearranged, run_reference
    d = make_rearranged(workdir)
    if not os.path.exists(os.path.join(d, "ref.sa")):
        build_index(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
    ref = run_reference(os.path.join(d, "ref"), os.path.join(d, "qry.fa"), os.path.join(d, "cpu_%d.pkl" % len(prm)), **prm)
    ix = oracle.index(bwaidx.load(os.path.join(d, "ref")))
    P = orc.params(**prm)
    key = lambda b: (b[0], b[3])
    for (name, seq), rc in zip(synth.read_fasta(os.path.join(d, "qry.fa")), ref):
        s = seq.tobytes()
        q, r, l = oracle.seed_contig(ix, P, s)
        assert np.array_equal(q, rc["seeds"][0]) and np.array_equal(r, rc["seeds"][1]) and np.array_equal(l, rc["seeds"][2])
        for stage in (0, 1, 2):
            mine = oracle.cluster(ix, P, s, q, r, l, stage)
            assert sorted(mine, key=key) == sorted(rc["stages"][stage], key=key), (name, stage)
        k = 0
        for b in rc["stages"][4]:                      # fragment lists + rows after GenerateFragAlignment
            assert oracle.normal_pairs([f for f in b[3] if f[0]]) == b[3]
            sc = al = 0
            for f in b[3]:
                if f[0]:
                    sc += f[3]; al += f[3]
                    continue
                a1, a2, inc, _ = oracle.frag_align(ix, s, f[1], f[2], f[3], f[4])
                assert (a1, a2) == (rc["aln"][k], rc["aln"][k + 1])
                k += 2; sc += inc; al += len(a1)
            assert (sc, al) == (b[0], b[1])
