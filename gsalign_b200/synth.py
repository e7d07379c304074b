"""Seeded synthetic genome pairs (SURVEY.md section 8d generator spec).

Reference: K contigs ``chr1..K`` of equal length, bases iid uniform over ACGT (upper case),
PRNG = numpy ``PCG64(seed)``.  Query contig k = copy of reference contig k with
  (i)  SNV: every base independently with probability ``p_snv`` replaced by one of the three other
       bases, uniformly;
  (ii) indel: every position independently with probability ``p_indel`` starts an event, 50/50
       insertion of L iid bases / deletion of L bases, L uniform 1..10.
Query names ``qchr1..K``; 80-column FASTA.  The generator is deterministic in (seed, sizes, rates)
and must never change once numbers have been published against it.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)

# name -> (N, K, p_snv, p_indel, seed)           BASELINE.json configs / SURVEY.md 8d
CONFIGS = {
    "C2": (100_000_000, 4, 0.01, 0.0, 2),
    "C3": (1_000_000_000, 8, 0.02, 0.002, 3),
    "C4": (3_000_000_000, 24, 0.01, 0.001, 4),
    "C5": (500_000_000, 4, 0.10, 0.0, 5),
}


def mutate(codes: np.ndarray, rng: np.random.Generator, p_snv: float, p_indel: float) -> np.ndarray:
    """codes: uint8 array of 0..3.  Returns the mutated copy (uint8 codes)."""
    n = codes.shape[0]
    out = codes.copy()
    if p_snv > 0:
        hit = rng.random(n) < p_snv
        k = int(hit.sum())
        out[hit] = (out[hit] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    if p_indel <= 0:
        return out
    ev = np.flatnonzero(rng.random(n) < p_indel)
    is_ins = rng.random(ev.shape[0]) < 0.5
    length = rng.integers(1, 11, size=ev.shape[0])
    keep = np.ones(n, dtype=bool)
    dpos, dlen = ev[~is_ins], length[~is_ins]
    for L in range(1, 11):  # deletions: drop [pos, pos+L)
        p = dpos[dlen >= L] + (L - 1)
        keep[p[p < n]] = False
    ins_len = np.zeros(n, dtype=np.int64)
    ins_len[ev[is_ins]] = length[is_ins]
    count = ins_len + keep
    ends = np.cumsum(count)
    total = int(ends[-1]) if n else 0
    res = rng.integers(0, 4, size=total, dtype=np.uint8)  # inserted bases stay random
    kept_idx = np.flatnonzero(keep)
    res[ends[kept_idx] - 1] = out[kept_idx]               # the kept base is the last of its slot
    return res


def make_pair(n_total: int, k: int, p_snv: float, p_indel: float, seed: int):
    """Returns (ref_contigs, qry_contigs): lists of (name, uint8 ASCII array)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    per = n_total // k
    ref, qry = [], []
    for i in range(k):
        codes = rng.integers(0, 4, size=per, dtype=np.uint8)
        ref.append((f"chr{i + 1}", _ACGT[codes]))
        qry.append((f"qchr{i + 1}", _ACGT[mutate(codes, rng, p_snv, p_indel)]))
    return ref, qry


def revcomp_ascii(a: np.ndarray) -> np.ndarray:
    lut = np.zeros(256, dtype=np.uint8)
    for x, y in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
        lut[x] = y
    return lut[a[::-1]]


def write_fasta(path: str, contigs, width: int = 80) -> None:
    with open(path, "wb") as f:
        for name, seq in contigs:
            f.write(b">" + name.encode() + b"\n")
            n = seq.shape[0]
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def read_fasta(path: str):
    """Minimal FASTA reader for tests (names cut at the first whitespace)."""
    import lzma
    op = lzma.open if path.endswith(".xz") else open
    out, name, parts = [], None, []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line[:1] == b">":
                if name is not None:
                    out.append((name, np.frombuffer(b"".join(parts), dtype=np.uint8)))
                name, parts = line[1:].split()[0].decode(), []
            else:
                parts.append(line)
    if name is not None:
        out.append((name, np.frombuffer(b"".join(parts), dtype=np.uint8)))
    return out


def make_dp_batch(rng, n_pairs, L, div=0.10, jitter=0.1):
    """DP-only stress batch: n_pairs fragment pairs, query = ref with `div` substitutions, lengths
    L*(1 +- jitter); returns (ref bytes, ref offsets, query bytes, query offsets) for gsa_dp_batch"""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens_r = np.maximum(1, (L * (1 + jitter * (rng.random(n_pairs) * 2 - 1))).astype(np.int64))
    lens_q = np.maximum(1, lens_r + rng.integers(-min(3, L // 4), min(3, L // 4) + 1, size=n_pairs))
    ro = np.zeros(n_pairs + 1, dtype=np.int64); ro[1:] = np.cumsum(lens_r)
    qo = np.zeros(n_pairs + 1, dtype=np.int64); qo[1:] = np.cumsum(lens_q)
    rb = acgt[rng.integers(0, 4, size=int(ro[-1]) + 1, dtype=np.uint8)]
    qb = acgt[rng.integers(0, 4, size=int(qo[-1]) + 1, dtype=np.uint8)]
    # copy the common prefix of every pair from the reference, then substitute
    for i in range(n_pairs) if n_pairs <= 4096 else []:
        k = int(min(lens_r[i], lens_q[i])); qb[qo[i]:qo[i] + k] = rb[ro[i]:ro[i] + k]
    if n_pairs > 4096:  # vectorised: position-wise copy where both exist
        idx_pair = np.repeat(np.arange(n_pairs), np.minimum(lens_r, lens_q))
        within = np.arange(idx_pair.shape[0]) - np.repeat(np.concatenate([[0], np.cumsum(np.minimum(lens_r, lens_q))[:-1]]), np.minimum(lens_r, lens_q))
        qb[qo[idx_pair] + within] = rb[ro[idx_pair] + within]
    sub = rng.random(qb.shape[0]) < div
    qb[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    return rb, ro, qb, qo
