"""K1 parity: gsa_seed through the C ABI vs the oracle's restatement of IdentifyLocalMEM/BWT_Search."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[False, True], ids=["rows32", "rows64"])
def aligner(ecoli, request):
    """both row widths of the device index: 32-bit rows (texts < 2^32 symbols) and the 64-bit layout human-size texts use"""
    from gsalign_b200 import capi
    a = capi.Aligner(0, wide=request.param)
    a.upload_index(ecoli["index"])
    yield a
    a.close()


def _seeds(aligner, seq, **prm):
    aligner.set_params(**prm)
    aligner.contig_begin(seq)
    n = aligner.seed()
    return aligner.fetch_seeds(n)


def test_ecoli_seeds_bit_exact(aligner, ecoli, oracle):
    ix = oracle.index(ecoli["index"])
    q, r, l = _seeds(aligner, ecoli["query"])
    oq, orr, ol = oracle.seed_contig(ix, orc.params(), ecoli["query"])
    assert len(q) == len(oq) == 102812
    assert np.array_equal(q, oq) and np.array_equal(r, orr) and np.array_equal(l, ol)


@pytest.mark.parametrize("slen,sen", [(10, 1), (12, 0), (20, 0), (30, 0)])
def test_seed_params(aligner, ecoli, oracle, slen, sen):
    ix = oracle.index(ecoli["index"])
    seq = ecoli["query"][1_000_000:1_200_000]
    q, r, l = _seeds(aligner, seq, min_seed_len=slen, sensitive=sen)
    oq, orr, ol = oracle.seed_contig(ix, orc.params(min_seed_len=slen, sensitive=sen), seq)
    assert np.array_equal(q, oq) and np.array_equal(r, orr) and np.array_equal(l, ol)


def test_seed_edge_cases(aligner, ecoli, oracle):
    """N runs, lower case, reverse strand, junction-straddling pieces, tiny and empty contigs."""
    from gsalign_b200 import synth
    ix = oracle.index(ecoli["index"])
    base = np.frombuffer(ecoli["query"], dtype=np.uint8)
    rc = synth.revcomp_ascii(base[2_000_000:2_030_000])
    pieces = [base[:25_000].copy(), np.frombuffer(b"N" * 37, dtype=np.uint8), rc, np.frombuffer(b"nnnnRYKM", dtype=np.uint8),
              base[4_630_000:].copy(), base[:5_000].copy()]
    pieces[0][100:140] = ord("N")
    pieces[0][9_990:10_010] = np.frombuffer(bytes(pieces[0][9_990:10_010]).lower(), dtype=np.uint8)
    seq = np.concatenate(pieces).tobytes()
    for s in (seq, seq[:9], seq[:15], seq[:10_000], seq[:10_001], b"", b"N" * 50):
        q, r, l = _seeds(aligner, s)
        oq, orr, ol = oracle.seed_contig(ix, orc.params(), s)
        assert np.array_equal(q, oq) and np.array_equal(r, orr) and np.array_equal(l, ol), len(s)
