"""ctypes bindings for the test oracles: oracle/liboracle.so (our C restatement) and
oracle/_ref/libgsref.so (the unmodified reference behind ref_shim.cpp).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_AVAILABLE = os.path.isdir("/root/reference/src")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OrcIndex(C.Structure):
    _fields_ = [("bwt", C.POINTER(C.c_uint32)), ("bwt_size", C.c_uint64), ("primary", C.c_uint64),
                ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("sa", C.POINTER(C.c_uint64)),
                ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("pac", C.POINTER(C.c_uint8)),
                ("l_pac", C.c_int64), ("n_contigs", C.c_int32), ("contig_off", C.POINTER(C.c_int64)),
                ("contig_len", C.POINTER(C.c_int32))]


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("min_seed_len", "sensitive", "max_indel", "min_block_score",
                                         "min_aln_len", "min_idy")]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_search", "n_ext_steps", "n_split", "n_sa_reads", "n_lf_steps",
                                          "n_seedhit", "n_short", "n_freqskip")]

    def algorithmic_bytes(self, query_bp: int, n_seeds: int) -> int:
        """B_seed of SURVEY.md section 8d."""
        return (64 * (self.n_ext_steps + self.n_split) + 64 * self.n_lf_steps + 8 * self.n_sa_reads
                + query_bp + 16 * n_seeds)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    return os.path.join(ORACLE_DIR, "liboracle.so")


def build_ref():
    if REF_AVAILABLE:
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref", "-j8"], check=True)
    return REF_DIR


def params(**kw) -> OrcParams:
    p = OrcParams(15, 0, 25, 200, 200, 70)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def parse_blocks(stream: np.ndarray):
    """int64 stream -> list of (score, aln_len, bDup, [(bSeed,q,r,qLen,rLen), ...])"""
    out, w = [], 1
    for _ in range(int(stream[0])):
        score, aln_len, dup, nf = (int(x) for x in stream[w:w + 4]); w += 4
        fr = stream[w:w + 5 * nf].reshape(nf, 5); w += 5 * nf
        out.append((score, aln_len, dup, [tuple(int(v) for v in row) for row in fr]))
    return out


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_seed_contig.restype = C.c_int64
        L.orc_cluster.restype = C.c_int64
        L.orc_normal_pairs.restype = C.c_int64
        L.orc_dp_align.restype = C.c_int32
        L.orc_frag_align.restype = C.c_int32
        L.orc_gap_similarity.restype = C.c_int32
        L.orc_text_char.restype = C.c_char
        self._keep = []

    def index(self, bi) -> OrcIndex:
        ix = OrcIndex()
        ix.bwt = _p(bi.bwt, C.c_uint32); ix.bwt_size = bi.bwt.shape[0]; ix.primary = bi.primary
        for i in range(5):
            ix.L2[i] = int(bi.L2[i])
        ix.seq_len = bi.seq_len; ix.sa = _p(bi.sa, C.c_uint64); ix.n_sa = bi.sa.shape[0]; ix.sa_intv = bi.sa_intv
        ix.pac = _p(bi.pac, C.c_uint8); ix.l_pac = bi.l_pac; ix.n_contigs = len(bi.names)
        ix.contig_off = _p(bi.contig_off, C.c_int64); ix.contig_len = _p(bi.contig_len, C.c_int32)
        self._keep.append(bi)
        return ix

    def bwt_search(self, ix, seq: bytes, start, stop, min_seed_len=15, ctr=None):
        ln, fq = C.c_int32(), C.c_int32()
        loc = (C.c_uint64 * 100)()
        self.lib.orc_bwt_search(C.byref(ix), seq, C.c_int32(start), C.c_int32(stop), C.c_int32(min_seed_len),
                                C.byref(ln), C.byref(fq), loc, C.byref(ctr) if ctr is not None else None)
        return ln.value, fq.value, [int(loc[i]) for i in range(fq.value)]

    def seed_contig(self, ix, prm, seq: bytes, ctr=None):
        q, r, l = C.POINTER(C.c_int32)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        n = self.lib.orc_seed_contig(C.byref(ix), C.byref(prm), seq, C.c_int64(len(seq)), C.byref(q), C.byref(r),
                                     C.byref(l), C.byref(ctr) if ctr is not None else None)
        qa = np.ctypeslib.as_array(q, (max(n, 1),))[:n].copy()
        ra = np.ctypeslib.as_array(r, (max(n, 1),))[:n].copy()
        la = np.ctypeslib.as_array(l, (max(n, 1),))[:n].copy()
        for ptr in (q, r, l):
            self.lib.orc_free(ptr)
        return qa, ra, la

    def cluster(self, ix, prm, seq: bytes, q, r, l, stage: int):
        out = C.POINTER(C.c_int64)()
        q = np.ascontiguousarray(q, dtype=np.int32); r = np.ascontiguousarray(r, dtype=np.int64)
        l = np.ascontiguousarray(l, dtype=np.int32)
        w = self.lib.orc_cluster(C.byref(ix), C.byref(prm), seq, C.c_int64(len(seq)), C.c_int64(q.shape[0]),
                                 _p(q, C.c_int32), _p(r, C.c_int64), _p(l, C.c_int32), C.c_int32(stage), C.byref(out))
        s = np.ctypeslib.as_array(out, (w,)).copy()
        self.lib.orc_free(out)
        return parse_blocks(s)

    def normal_pairs(self, frags):
        a = np.ascontiguousarray(np.array(frags, dtype=np.int64).reshape(-1, 5))
        out = C.POINTER(C.c_int64)()
        n = self.lib.orc_normal_pairs(C.c_int64(a.shape[0]), _p(a, C.c_int64), C.byref(out))
        s = np.ctypeslib.as_array(out, (n * 5,)).copy().reshape(n, 5)
        self.lib.orc_free(out)
        return [tuple(int(v) for v in row) for row in s]

    def dp_align(self, ref_frag: bytes, qry_frag: bytes):
        m, n = len(ref_frag), len(qry_frag)
        o1, o2 = C.create_string_buffer(m + n + 1), C.create_string_buffer(m + n + 1)
        L = self.lib.orc_dp_align(ref_frag, C.c_int32(m), qry_frag, C.c_int32(n), o1, o2)
        return o1.raw[:L], o2.raw[:L]

    def frag_align(self, ix, seq: bytes, qpos, rpos, qlen, rlen):
        o1, o2 = C.create_string_buffer(qlen + rlen + 1), C.create_string_buffer(qlen + rlen + 1)
        sc, dp = C.c_int32(), C.c_int32()
        L = self.lib.orc_frag_align(C.byref(ix), seq, C.c_int32(qpos), C.c_int64(rpos), C.c_int32(qlen),
                                    C.c_int32(rlen), o1, o2, C.byref(sc), C.byref(dp))
        return o1.raw[:L], o2.raw[:L], sc.value, dp.value

    def gap_similarity(self, ix, seq: bytes, q1, q2, r1, r2) -> int:
        return self.lib.orc_gap_similarity(C.byref(ix), seq, C.c_int32(q1), C.c_int32(q2), C.c_int64(r1), C.c_int64(r2))


class Reference:
    """The real thing (oracle/_ref/libgsref.so).  One index per process (the reference uses globals)."""

    def __init__(self, prefix: str):
        path = os.path.join(build_ref(), "libgsref.so")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_seed_contig.restype = C.c_int64
        L.ref_stage_size.restype = C.c_int64
        L.ref_aln_size.restype = C.c_int64
        L.ref_genome_size.restype = C.c_int64
        L.ref_ksw2.restype = C.c_int
        L.ref_gap_similarity.restype = C.c_int
        devnull = os.open(os.devnull, os.O_WRONLY)
        self._saved_err = os.dup(2)
        os.dup2(devnull, 2)  # the reference prints a progress line per 10 kb chunk
        try:
            if L.ref_load_index(prefix.encode(), 4) != 0:
                raise RuntimeError("ref_load_index failed")
        finally:
            os.dup2(self._saved_err, 2)
        self.set_params()

    def set_params(self, min_seed_len=15, sensitive=0, max_indel=25, min_block_score=200, min_aln_len=200,
                   min_idy=70, one=0):
        self.lib.ref_set_params(min_seed_len, sensitive, max_indel, min_block_score, min_aln_len, min_idy, one)

    def _quiet(self, fn, *a):
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)
        try:
            return fn(*a)
        finally:
            os.dup2(saved, 2); os.close(saved); os.close(devnull)

    def bwt_search(self, seq: bytes, start, stop):
        ln, fq = C.c_int(), C.c_int()
        loc = (C.c_uint64 * 100)()
        self.lib.ref_bwt_search(seq, len(seq), start, stop, C.byref(ln), C.byref(fq), loc)
        return ln.value, fq.value, [int(loc[i]) for i in range(fq.value)]

    def seed_contig(self, seq: bytes):
        n = self._quiet(self.lib.ref_seed_contig, seq, C.c_int64(len(seq)))
        q = np.empty(n, dtype=np.int32); r = np.empty(n, dtype=np.int64); l = np.empty(n, dtype=np.int32)
        self.lib.ref_get_seeds(_p(q, C.c_int32), _p(r, C.c_int64), _p(l, C.c_int32))
        return q, r, l

    def cluster(self):
        """Runs all phases on the seeds of the last seed_contig(); returns {stage: blocks}, aln strings."""
        self._quiet(self.lib.ref_cluster)
        stages = {}
        for s in range(6):
            w = self.lib.ref_stage_size(s)
            a = np.empty(w, dtype=np.int64)
            self.lib.ref_stage_copy(s, _p(a, C.c_int64))
            stages[s] = parse_blocks(a)
        n = self.lib.ref_aln_size()
        buf = C.create_string_buffer(n + 1)
        self.lib.ref_aln_copy(buf)
        return stages, buf.raw[:n].split(b"\n")[:-1]

    def ksw2(self, ref_frag: bytes, qry_frag: bytes):
        m, n = len(ref_frag), len(qry_frag)
        o1, o2 = C.create_string_buffer(m + n + 1), C.create_string_buffer(m + n + 1)
        L = self.lib.ref_ksw2(ref_frag, m, qry_frag, n, o1, o2)
        return o1.raw[:L], o2.raw[:L]

    def gap_similarity(self, q1, q2, r1, r2) -> int:
        return self.lib.ref_gap_similarity(q1, q2, C.c_int64(r1), C.c_int64(r2))

    def text(self, pos, n) -> bytes:
        buf = C.create_string_buffer(n)
        self.lib.ref_text(C.c_int64(pos), C.c_int64(n), buf)
        return buf.raw
