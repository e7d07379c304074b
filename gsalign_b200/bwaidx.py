"""Reader for BWA-format index files (the on-disk contract of the reference).

Formats (written by src/BWT_Index/bwtindex.c:53-75, bwt.c:174-196,
bntseq.c:59-89,192-201; read by src/bwt_index.cpp:15-121):
  .bwt  5 x u64 (primary, L2[1..4]) then u32 words: every 128 symbols are preceded by 4 x u64 counts
  .sa   7 x u64 (primary, L2[1..4], sa_intv, seq_len) then (n_sa - 1) x u64; sa[0] = -1 is implicit
  .pac  forward strand, 2 bit/base MSB first, + trailing length byte(s)
  .ann  text: "l_pac n_seqs seed" then per contig "gi name comment" / "offset len n_ambs"
This module is harness-side plumbing for tests and bench.py; the product's loader is C++
(gsalign_b200/csrc/host/index_io.cpp).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class BwaIndex:
    prefix: str
    primary: int
    L2: np.ndarray            # u64[5]
    seq_len: int              # 2N
    bwt: np.ndarray           # u32 words (interleaved Occ)
    sa_intv: int
    sa: np.ndarray            # u64[n_sa], sa[0] = 2^64-1
    pac: np.ndarray           # u8
    l_pac: int                # N
    names: list = field(default_factory=list)
    contig_off: np.ndarray = None   # i64[n]
    contig_len: np.ndarray = None   # i32[n]


def load(prefix: str) -> BwaIndex:
    raw = np.fromfile(prefix + ".bwt", dtype=np.uint8)
    hdr = raw[:40].view(np.uint64)
    bwt = raw[40:].view(np.uint32)
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = hdr[1:5]
    seq_len = int(L2[4])
    sraw = np.fromfile(prefix + ".sa", dtype=np.uint64)
    sa_intv = int(sraw[5])
    n_sa = (seq_len + sa_intv) // sa_intv
    sa = np.empty(n_sa, dtype=np.uint64)
    sa[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    sa[1:] = sraw[7:7 + n_sa - 1]
    names, offs, lens = [], [], []
    with open(prefix + ".ann") as f:
        l_pac, n_seqs, _seed = f.readline().split()[:3]
        l_pac, n_seqs = int(l_pac), int(n_seqs)
        for _ in range(n_seqs):
            names.append(f.readline().split()[1])
            o, ln, _ = f.readline().split()[:3]
            offs.append(int(o)); lens.append(int(ln))
    pac = np.fromfile(prefix + ".pac", dtype=np.uint8)[: l_pac // 4 + 1].copy()
    return BwaIndex(prefix, int(hdr[0]), L2, seq_len, np.ascontiguousarray(bwt), sa_intv, sa, pac, l_pac,
                    names, np.array(offs, dtype=np.int64), np.array(lens, dtype=np.int32))


def text(idx: BwaIndex) -> np.ndarray:
    """T = F . revcomp(F) as uint8 codes 0..3 (length 2N)."""
    n = idx.l_pac
    shifts = np.array([6, 4, 2, 0], dtype=np.uint8)
    f = ((idx.pac[:, None] >> shifts[None, :]) & 3).reshape(-1)[:n].astype(np.uint8)
    return np.concatenate([f, (3 - f[::-1]).astype(np.uint8)])
