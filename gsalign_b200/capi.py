"""ctypes binding of include/gsalign_b200.h (harness side: tests, bench.py, smoke).

The product is the shared library ``gsalign_b200/libgsalign_b200.so`` (CUDA kernels + C ABI) and the
``bin/GSAlign`` CLI built on it; this module only lets Python call the same entry points a C/C++ host
would.  There is NO fallback: if the library is missing or no GPU is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsalign_b200.so")


class GsaError(RuntimeError):
    pass


class IndexView(C.Structure):
    _fields_ = [("bwt", C.POINTER(C.c_uint32)), ("bwt_size", C.c_uint64), ("primary", C.c_uint64),
                ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("sa", C.POINTER(C.c_uint64)),
                ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("pac", C.POINTER(C.c_uint8)),
                ("l_pac", C.c_int64), ("n_contigs", C.c_int32), ("contig_off", C.POINTER(C.c_int64)),
                ("contig_len", C.POINTER(C.c_int32))]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("min_seed_len", "sensitive", "max_indel", "min_block_score",
                                         "min_aln_len", "min_idy", "one_on_one")]


class Frag(C.Structure):
    _fields_ = [("rPos", C.c_int64), ("qPos", C.c_int32), ("qLen", C.c_int32), ("rLen", C.c_int32),
                ("bSeed", C.c_int32), ("aln_off", C.c_int64), ("aln_len", C.c_int32), ("reserved", C.c_int32)]


FRAG_DTYPE = np.dtype([("rPos", "<i8"), ("qPos", "<i4"), ("qLen", "<i4"), ("rLen", "<i4"), ("bSeed", "<i4"),
                       ("aln_off", "<i8"), ("aln_len", "<i4"), ("reserved", "<i4")])
BLOCK_DTYPE = np.dtype([("score", "<i4"), ("aln_len", "<i4"), ("bDup", "<i4"), ("n_frags", "<i4"), ("frag_beg", "<i8")])
VARIANT_DTYPE = np.dtype([("rPos", "<i8"), ("qPos", "<i4"), ("gPos", "<i4"), ("len", "<i4"), ("kind", "<i4")])   # gsa_variant


class Alignment(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("blocks", C.c_void_p), ("n_frags", C.c_int64), ("frags", C.c_void_p),
                ("aln_bytes", C.c_int64), ("aln1", C.c_void_p), ("aln2", C.c_void_p)]


class VariantList(C.Structure):
    _fields_ = [("n_variants", C.c_int64), ("variants", C.c_void_p), ("block_first", C.c_void_p), ("block_count", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "seed_ms", "cluster_ms", "fill_ms", "d2h_ms", "host_ms")] + \
               [(n, C.c_int64) for n in ("n_seeds", "n_dp", "dp_cells", "n_frags", "launches")] + \
               [(n, C.c_float) for n in ("k_seed_ms", "k_dp_ms", "total_ms")]


EXPORTS = ["gsa_create", "gsa_destroy", "gsa_last_error", "gsa_index_upload", "gsa_set_params", "gsa_default_params",
           "gsa_contig_begin", "gsa_contig_begin_device", "gsa_seed", "gsa_cluster", "gsa_fill", "gsa_align_contig",
           "gsa_get_timing", "gsa_fetch_seeds", "gsa_dump_blocks", "gsa_dp_batch", "gsa_set_stream", "gsa_set_dump", "gsa_dpx_peak", "gsa_create_shared", "gsa_result_device", "gsa_set_host_results", "gsa_dp_batch_identity",
           "gsa_set_wide_index", "gsa_index_clone", "gsa_index_bytes", "gsa_index_selfcheck",
           "gsa_comm_unique_id", "gsa_comm_init_rank", "gsa_comm_init_all", "gsa_comm_destroy", "gsa_outbox_reset", "gsa_outbox_reserve",
           "gsa_outbox_append", "gsa_outbox_bytes", "gsa_gather_records", "gsa_gather_records_all", "gsa_gather_wait", "gsa_inbox_device",
           "gsa_inbox_host", "gsa_record_next", "gsa_record_frags", "gsa_variants", "gsa_contig_prefetch", "gsa_split_hazard"]


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise GsaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gsa_last_error.restype = C.c_char_p
    lib.gsa_dump_blocks.restype = C.c_int64
    lib.gsa_index_bytes.restype = C.c_int64
    lib.gsa_outbox_bytes.restype = C.c_int64
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _torch_first():
    """The library binds NCCL at run time (dlopen of libnccl.so.2).  PyTorch bundles its own, newer libnccl under the same
    soname: whichever copy a process loads first is the one both get, and torch does not import against an older one.  A
    Python host that may import torch later therefore imports it before the library opens NCCL."""
    try:
        import torch  # noqa: F401
    except ImportError:
        pass


def _as_array(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


def walk_image(lib, ptr, nbytes: int):
    """the records of an outbox image in host memory as (contig, blocks, frags, aln1, aln2) numpy copies; fragment lists that
    travelled in the compact form are expanded by gsa_record_frags"""
    out, off, contig, al = [], C.c_int64(0), C.c_int64(), Alignment()
    while True:
        start = off.value
        rc = lib.gsa_record_next(ptr, C.c_int64(nbytes), C.byref(off), C.byref(contig), C.byref(al))
        if rc < 0:
            raise GsaError("malformed outbox image")
        if rc == 0:
            break
        blocks = _as_array(al.blocks, al.n_blocks, BLOCK_DTYPE).copy()
        frags = np.zeros(al.n_frags, dtype=FRAG_DTYPE)
        if al.n_frags and lib.gsa_record_frags(ptr, C.c_int64(nbytes), C.c_int64(start), frags.ctypes.data_as(C.c_void_p), C.c_int32(4)) != 0:
            raise GsaError("malformed outbox image (fragment list)")
        a1 = _as_array(al.aln1, al.aln_bytes, np.dtype(np.uint8)).copy()
        a2 = _as_array(al.aln2, al.aln_bytes, np.dtype(np.uint8)).copy()
        out.append((contig.value, blocks, frags, a1, a2))
    return out


class Aligner:
    """One context per GPU; mirrors the per-contig loop of GenomeComparison (reference src/GSAlign.cpp:473)."""

    def __init__(self, device: int = 0, owner: "Aligner | None" = None, wide: bool = False):
        """owner: create a lane on the owner's GPU that shares its uploaded index (gsa_create_shared);
        wide: force the 64-bit row layout of the device index whatever the text size (gsa_set_wide_index)"""
        self.lib = load_library()
        self.ctx = C.c_void_p()
        self._owner = owner
        if owner is not None:
            rc = self.lib.gsa_create_shared(owner.ctx, C.byref(self.ctx))
            if rc != 0:
                raise GsaError(f"gsa_create_shared failed with {rc}: {self.lib.gsa_last_error(owner.ctx).decode()}")
        else:
            rc = self.lib.gsa_create(C.c_int(device), C.byref(self.ctx))
            if rc != 0:
                raise GsaError(f"gsa_create(device={device}) failed with {rc}: no usable B200; there is no CPU fallback")
        self._keep = None
        if wide:
            self._chk(self.lib.gsa_set_wide_index(self.ctx, C.c_int(1)))

    def clone_index_from(self, src: "Aligner"):
        """replica of src's device index on this context's GPU, copied GPU to GPU (gsa_index_clone)"""
        self._chk(self.lib.gsa_index_clone(self.ctx, src.ctx))

    def index_selfcheck(self, n_samples: int = 1 << 20) -> int:
        bad = C.c_int64()
        self._chk(self.lib.gsa_index_selfcheck(self.ctx, C.c_int64(n_samples), C.byref(bad)))
        return bad.value

    def index_bytes(self) -> int:
        return int(self.lib.gsa_index_bytes(self.ctx))

    def close(self):
        if self.ctx:
            self.lib.gsa_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise GsaError(f"gsalign_b200 error {rc}: {self.lib.gsa_last_error(self.ctx).decode()}")

    def upload_index(self, bi):
        v = IndexView()
        v.bwt = _p(bi.bwt, C.c_uint32); v.bwt_size = bi.bwt.shape[0]; v.primary = bi.primary
        for i in range(5):
            v.L2[i] = int(bi.L2[i])
        v.seq_len = bi.seq_len; v.sa = _p(bi.sa, C.c_uint64); v.n_sa = bi.sa.shape[0]; v.sa_intv = bi.sa_intv
        v.pac = _p(bi.pac, C.c_uint8); v.l_pac = bi.l_pac; v.n_contigs = len(bi.names)
        v.contig_off = _p(bi.contig_off, C.c_int64); v.contig_len = _p(bi.contig_len, C.c_int32)
        self._chk(self.lib.gsa_index_upload(self.ctx, C.byref(v)))

    def set_params(self, **kw):
        p = Params()
        self.lib.gsa_default_params(C.byref(p))
        for k, val in kw.items():
            setattr(p, k, val)
        self._chk(self.lib.gsa_set_params(self.ctx, C.byref(p)))

    def set_stream(self, cuda_stream: int):
        self._chk(self.lib.gsa_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def set_dump(self, enable: bool):
        self._chk(self.lib.gsa_set_dump(self.ctx, C.c_int(1 if enable else 0)))

    def contig_begin(self, seq):
        """seq: bytes or uint8 ndarray (host)."""
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        self._keep = a
        self._chk(self.lib.gsa_contig_begin(self.ctx, a.ctypes.data_as(C.c_char_p), C.c_uint32(a.shape[0])))

    def contig_begin_device(self, dev_ptr: int, n: int):
        self._chk(self.lib.gsa_contig_begin_device(self.ctx, C.c_void_p(dev_ptr), C.c_uint32(n)))

    def seed(self) -> int:
        n = C.c_int64()
        self._chk(self.lib.gsa_seed(self.ctx, C.byref(n)))
        return n.value

    def fetch_seeds(self, n):
        q = np.empty(n, dtype=np.int32); r = np.empty(n, dtype=np.int64); l = np.empty(n, dtype=np.int32)
        self._chk(self.lib.gsa_fetch_seeds(self.ctx, _p(q, C.c_int32), _p(r, C.c_int64), _p(l, C.c_int32)))
        return q, r, l

    def cluster(self) -> int:
        n = C.c_int32()
        self._chk(self.lib.gsa_cluster(self.ctx, C.byref(n)))
        return n.value

    def dump_blocks(self, stage: int) -> np.ndarray:
        w = self.lib.gsa_dump_blocks(self.ctx, C.c_int32(stage), None)
        if w < 0:
            self._chk(int(w))
        out = np.empty(w, dtype=np.int64)
        w2 = self.lib.gsa_dump_blocks(self.ctx, C.c_int32(stage), _p(out, C.c_int64))
        if w2 < 0:
            self._chk(int(w2))
        return out

    def _alignment(self, al: Alignment):
        blocks = _as_array(al.blocks, al.n_blocks, BLOCK_DTYPE).copy()
        frags = _as_array(al.frags, al.n_frags, FRAG_DTYPE).copy()
        a1 = _as_array(al.aln1, al.aln_bytes, np.dtype("u1")).copy()
        a2 = _as_array(al.aln2, al.aln_bytes, np.dtype("u1")).copy()
        return blocks, frags, a1, a2

    def fill(self):
        al = Alignment()
        self._chk(self.lib.gsa_fill(self.ctx, C.byref(al)))
        return self._alignment(al)

    def align_contig(self, seq):
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        self._keep = a
        al = Alignment()
        self._chk(self.lib.gsa_align_contig(self.ctx, a.ctypes.data_as(C.c_char_p), C.c_uint32(a.shape[0]), C.byref(al)))
        return self._alignment(al)

    def variants(self, n_blocks: int):
        """N3: the variant records of the last fill()/align_contig() (device scan of the rows): (records, block_first,
        block_count), the last two aligned with the blocks of that result"""
        vl = VariantList()
        self._chk(self.lib.gsa_variants(self.ctx, C.byref(vl)))
        rec = _as_array(vl.variants, vl.n_variants, VARIANT_DTYPE).copy() if vl.n_variants else np.zeros(0, dtype=VARIANT_DTYPE)
        if n_blocks == 0 or not vl.block_first:
            return rec, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        return rec, _as_array(vl.block_first, n_blocks, np.dtype("<i8")).copy(), _as_array(vl.block_count, n_blocks, np.dtype("<i8")).copy()

    def split_hazard(self) -> int:
        """hazard H14: split phases of the last cluster() whose pushes crossed a power of two (reference behaviour undefined there)"""
        return int(self.lib.gsa_split_hazard(self.ctx))

    def prefetch(self, a: np.ndarray):
        """starts the upload of the contig the next align_contig*() call on this context will get (gsa_contig_prefetch)"""
        self._chk(self.lib.gsa_contig_prefetch(self.ctx, a.ctypes.data_as(C.c_char_p), C.c_uint32(a.shape[0])))

    def align_contig_raw(self, a: np.ndarray) -> Alignment:
        """No copies of the result (bench path): returns the struct pointing into the library's pinned buffers."""
        al = Alignment()
        self._chk(self.lib.gsa_align_contig(self.ctx, a.ctypes.data_as(C.c_char_p), C.c_uint32(a.shape[0]), C.byref(al)))
        return al

    def set_host_results(self, enable: bool):
        self._chk(self.lib.gsa_set_host_results(self.ctx, C.c_int(1 if enable else 0)))

    def result_device(self) -> Alignment:
        """the last gsa_fill() result with DEVICE pointers for frags / aln1 / aln2 (blocks: host)"""
        al = Alignment()
        self._chk(self.lib.gsa_result_device(self.ctx, C.byref(al)))
        return al

    # ---- multi-GPU record gather (gather.cu) ----------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        _torch_first()
        buf = C.create_string_buffer(128)
        rc = load_library().gsa_comm_unique_id(buf, C.c_int32(128))
        if rc != 0:
            raise GsaError(f"gsa_comm_unique_id failed with {rc} (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init_rank(self, uid: bytes, rank: int, n_ranks: int):
        _torch_first()
        self._chk(self.lib.gsa_comm_init_rank(self.ctx, C.c_char_p(uid), C.c_int32(rank), C.c_int32(n_ranks)))

    def outbox_reset(self):
        self._chk(self.lib.gsa_outbox_reset(self.ctx))

    def outbox_append(self, lane: "Aligner", contig: int):
        rc = self.lib.gsa_outbox_append(self.ctx, lane.ctx, C.c_int64(contig))
        if rc != 0:
            raise GsaError(f"gsa_outbox_append failed with {rc}: {self.lib.gsa_last_error(lane.ctx).decode()} / {self.lib.gsa_last_error(self.ctx).decode()}")

    def outbox_bytes(self) -> int:
        return int(self.lib.gsa_outbox_bytes(self.ctx))

    def gather_records(self, root: int = 0):
        self._chk(self.lib.gsa_gather_records(self.ctx, C.c_int32(root)))

    def gather_wait(self):
        self._chk(self.lib.gsa_gather_wait(self.ctx))

    def inbox_records(self, rank: int):
        """on the root after gather_records + gather_wait: the records of `rank` as a list of
        (contig, blocks, frags, aln1, aln2) numpy copies"""
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._chk(self.lib.gsa_inbox_host(self.ctx, C.c_int32(rank), C.byref(ptr), C.byref(nbytes)))
        return walk_image(self.lib, ptr, nbytes.value)

    def timing(self) -> Timing:
        t = Timing()
        self._chk(self.lib.gsa_get_timing(self.ctx, C.byref(t)))
        return t

    def dpx_peak(self, which: int = 0) -> float:
        """issue rate of a packed-int16 DPX instruction, 1e9 thread-level instructions per second"""
        v = C.c_double()
        self._chk(self.lib.gsa_dpx_peak(self.ctx, C.c_int(which), C.byref(v)))
        return v.value

    def dp_batch_arrays(self, rb, ro, qb, qo):
        """gsa_dp_batch on concatenated uint8 arrays + int64 offset arrays; returns (rows1, rows2, lens, kernel_ms)"""
        n = ro.shape[0] - 1
        tot = int(ro[-1] + qo[-1]) + 1
        o1 = np.empty(tot, dtype=np.uint8); o2 = np.empty(tot, dtype=np.uint8); ol = np.zeros(max(n, 1), dtype=np.int32)
        ms = C.c_float()
        self._chk(self.lib.gsa_dp_batch(self.ctx, C.c_int32(n), rb.ctypes.data_as(C.c_char_p), _p(ro, C.c_int64),
                                        qb.ctypes.data_as(C.c_char_p), _p(qo, C.c_int64), o1.ctypes.data_as(C.c_char_p),
                                        o2.ctypes.data_as(C.c_char_p), _p(ol, C.c_int32), C.byref(ms)))
        return o1, o2, ol, ms.value

    def dp_batch_identity(self, refs, qrys):
        """refs/qrys: lists of bytes.  Returns list of (row1, row2, identical columns)."""
        n = len(refs)
        ro = np.zeros(n + 1, dtype=np.int64); qo = np.zeros(n + 1, dtype=np.int64)
        ro[1:] = np.cumsum([len(x) for x in refs]); qo[1:] = np.cumsum([len(x) for x in qrys])
        rb = np.frombuffer(b"".join(refs) + b"\0", dtype=np.uint8); qb = np.frombuffer(b"".join(qrys) + b"\0", dtype=np.uint8)
        tot = int(ro[-1] + qo[-1]) + 1
        o1 = np.zeros(tot, dtype=np.uint8); o2 = np.zeros(tot, dtype=np.uint8)
        ol = np.zeros(max(n, 1), dtype=np.int32); oi = np.zeros(max(n, 1), dtype=np.int32)
        self._chk(self.lib.gsa_dp_batch_identity(self.ctx, C.c_int32(n), rb.ctypes.data_as(C.c_char_p), _p(ro, C.c_int64),
                                                 qb.ctypes.data_as(C.c_char_p), _p(qo, C.c_int64), o1.ctypes.data_as(C.c_char_p),
                                                 o2.ctypes.data_as(C.c_char_p), _p(ol, C.c_int32), _p(oi, C.c_int32)))
        res = []
        for i in range(n):
            off = int(ro[i] + qo[i]); L = int(ol[i])
            res.append((o1[off:off + L].tobytes(), o2[off:off + L].tobytes(), int(oi[i])))
        return res

    def dp_batch(self, refs, qrys):
        """refs/qrys: lists of bytes.  Returns (list of (row1,row2)), kernel_ms."""
        n = len(refs)
        ro = np.zeros(n + 1, dtype=np.int64); qo = np.zeros(n + 1, dtype=np.int64)
        ro[1:] = np.cumsum([len(x) for x in refs]); qo[1:] = np.cumsum([len(x) for x in qrys])
        rb = np.frombuffer(b"".join(refs) + b"\0", dtype=np.uint8); qb = np.frombuffer(b"".join(qrys) + b"\0", dtype=np.uint8)
        tot = int(ro[-1] + qo[-1]) + 1
        o1 = np.zeros(tot, dtype=np.uint8); o2 = np.zeros(tot, dtype=np.uint8); ol = np.zeros(max(n, 1), dtype=np.int32)
        ms = C.c_float()
        self._chk(self.lib.gsa_dp_batch(self.ctx, C.c_int32(n), rb.ctypes.data_as(C.c_char_p), _p(ro, C.c_int64),
                                        qb.ctypes.data_as(C.c_char_p), _p(qo, C.c_int64), o1.ctypes.data_as(C.c_char_p),
                                        o2.ctypes.data_as(C.c_char_p), _p(ol, C.c_int32), C.byref(ms)))
        res = []
        for i in range(n):
            off = int(ro[i] + qo[i]); L = int(ol[i])
            res.append((o1[off:off + L].tobytes(), o2[off:off + L].tobytes()))
        return res, ms.value
