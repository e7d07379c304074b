"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/gsalign_b200.h declares; it refuses to run without a device (no CPU fallback); host-side helpers."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gsalign_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsa_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gsalign_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gsalign_b200.h but not exported"
    assert sorted(capi.EXPORTS) == [n for n in names if n in capi.EXPORTS]


def test_no_cpu_fallback():
    """without a GPU gsa_create must fail (loudly), never fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gsalign_b200 import capi
    with pytest.raises(capi.GsaError):
        capi.Aligner(0)


def test_product_never_touches_the_oracle():
    """nothing under gsalign_b200/ or include/ may reference oracle/ (test infrastructure only)"""
    bad = []
    for base in ("gsalign_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp:
                continue
            for f in files:
                if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"gsa_oracle|liboracle|libgsref|orc_[a-z]+\(", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_synth_generator_is_deterministic():
    from gsalign_b200 import synth
    r1, q1 = synth.make_pair(40_000, 2, 0.01, 0.001, 4)
    r2, q2 = synth.make_pair(40_000, 2, 0.01, 0.001, 4)
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(r1 + q1, r2 + q2))
    assert [n for n, _ in r1] == ["chr1", "chr2"] and [n for n, _ in q1] == ["qchr1", "qchr2"]
    assert not np.array_equal(r1[0][1][:2000], q1[0][1][:2000])
    rs, qs = synth.make_pair(40_000, 2, 0.01, 0.0, 4)           # SNV only: same length, ~1 % substitutions
    assert 0.003 < np.mean(rs[0][1] != qs[0][1]) < 0.03
    assert abs(len(q1[0][1]) - 20_000) < 200


def test_bwaidx_loader_roundtrip(ecoli):
    from gsalign_b200 import bwaidx
    bi = ecoli["index"]
    assert bi.seq_len == 2 * bi.l_pac == 2 * 4639675 and bi.sa_intv == 32
    assert bi.bwt.shape[0] * 4 + 40 == os.path.getsize(ecoli["prefix"] + ".bwt")
    assert int(bi.L2[4]) == bi.seq_len and bi.names == ["NC_000913"]
    t = bwaidx.text(bi)
    assert t.shape[0] == bi.seq_len and np.array_equal(t[:50], 3 - t[::-1][:50])
    # T is its own reverse complement: base counts of A/T and C/G agree with L2
    cnt = np.bincount(t, minlength=4)
    assert [int(bi.L2[i + 1] - bi.L2[i]) for i in range(4)] == cnt.tolist()


def test_lpt_sharding():
    from gsalign_b200.shard import lpt_assign
    lens = [125] * 24
    parts = lpt_assign(lens, 8)
    assert sorted(sum(parts, [])) == list(range(24)) and all(len(p) == 3 for p in parts)
    parts = lpt_assign([100, 1, 1, 1, 50, 49], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum([100, 1, 1, 1, 50, 49][i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 2


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsalign_b200.shard import lpt_assign
    lens = [300, 200, 120, 90, 80, 10]
    mine = lpt_assign(lens, world)[rank]
    # what bench.py does at N > 1: per-rank time -> MAX over ranks, units -> SUM over ranks
    t = torch.tensor([float(sum(lens[i] for i in mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(u)
    got = [None] * world
    dist.all_gather_object(got, mine)          # the record gather of SURVEY.md 8e, on host objects
    q.put((rank, float(t[0]), int(u[0]), got))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == res[0][1] and r[2] == 6 for r in res)
    assert sorted(res[0][3][0] + res[0][3][1]) == list(range(6))
    assert res[0][1] == 400.0


def _gather_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsalign_b200 import gather
    from gsalign_b200.shard import lpt_assign
    lens = [300, 200, 120, 90, 80, 10]
    mine = lpt_assign(lens, world)[rank]
    rng = np.random.default_rng(100 + rank)
    box = gather.Outbox(1 << 20, "cpu")
    sent = []
    for c in mine:   # one fake record per contig this rank owns; contig 5 has no alignment at all
        nb, nf, ab = (0, 0, 0) if c == 5 else (1 + c, 3 * c + 2, 17 * c + 5)
        parts = [torch.from_numpy(rng.integers(0, 256, size=k, dtype=np.uint8)) for k in (nb * gather.BLOCK_BYTES, nf * gather.FRAG_BYTES, ab, ab)]
        off = box.reserve(gather.record_bytes(nb, nf, ab))
        box.put(off, c, *parts)
        sent.append((c, [p.numpy().tobytes() for p in parts]))
    inbox = gather.gather_to_root(box.buf, box.used)
    got = None
    if rank == 0:
        got = [[(c, [x.tobytes() for x in rest]) for c, *rest in gather.unpack(b)] for b in inbox]
    q.put((rank, sent, got))
    dist.destroy_process_group()


def test_two_rank_gloo_record_gather():
    """the N > 1 record gather (gsalign_b200/gather.py): what every rank packed arrives on rank 0, byte for byte"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    got = res[0][2]
    assert got is not None and len(got) == 2
    for r in range(2):
        assert got[r] == res[r][1]
    assert sorted(c for r in range(2) for c, _ in got[r]) == list(range(6))


_LOADER_HARNESS = r'''
#include "host.h"
int main(int argc, char **argv)
{
	std::vector<QueryChr> q;
	bool ok = load_query_file(argv[1], q);
	printf("ok=%d n=%zu\n", ok ? 1 : 0, q.size());
	for (auto &c : q) printf("[%s] %zu %s\n", c.name.c_str(), c.seq.size(), c.seq.c_str());
	return 0;
}
'''


@pytest.mark.parametrize("std,piece", [("c++17", None), ("c++23", None), ("c++23", "7"), ("c++17", "1")])
def test_query_fasta_reader_edge_cases(tmp_path, std, piece):
    """the CLI's parallel, memory-mapped FASTA reader keeps LoadQueryFile's rules (reference src/main.cpp:35-114): header
    trimming, CR stripping, empty lines and records, no trailing newline, sequence before a header, non-letters.
    The reader cuts records into pieces at line starts (GSA_FASTA_PIECE = bytes per piece: tiny values put a border after
    every line) and, built as C++23 like bin/GSAlign, leaves the strings' letters to the copying threads."""
    import subprocess
    host = os.path.join(ROOT, "gsalign_b200", "csrc", "host")
    src = tmp_path / "ld.cpp"
    src.write_text(_LOADER_HARNESS)
    exe = str(tmp_path / "ld")
    subprocess.run(["g++", "-O1", f"-std={std}", "-I", host, "-o", exe, str(src), os.path.join(host, "io.cpp"), "-lpthread"], check=True)
    env = dict(os.environ, **({"GSA_FASTA_PIECE": piece} if piece else {}))

    def run(text: bytes):
        f = tmp_path / "q.fa"
        f.write_bytes(text)
        r = subprocess.run([exe, str(f)], capture_output=True, text=True, env=env)
        return r.stdout.splitlines(), r.stderr

    out, _ = run(b"\n\n>chr1 some comment\nACGTNNacgt\r\n\nGGGG\n>chr2|x:1\nTTTT\n>empty\n>last\nAC")
    assert out == ["ok=1 n=4", "[chr1] 14 ACGTNNacgtGGGG", "[chr2-x] 4 TTTT", "[empty] 0 ", "[last] 2 AC"]
    out, _ = run(b"ACGT\n>chr1\nAC\n")                       # sequence before any header
    assert out[0] == "ok=0 n=0"
    out, err = run(b">c\nAC1GT\n")                           # a non-letter: the line is echoed, the load fails
    assert out[0] == "AC1GT" and "ok=0" in out[1] and "non-alphabet" in err
    out, err = run(b">c\nACGT\n>d\nAC>GT\n")                 # '>' inside a sequence line is a non-letter too
    assert out[0] == "AC>GT" and "ok=0" in out[1]
    out, err = run(b">c\nACGT\nAC-T\nAAAA\nA.A\n>d\nA*\n")       # several bad lines: the first one in file order is the one echoed
    assert out[0] == "AC-T" and "ok=0" in out[1]
    out, _ = run(b">only header")
    assert out == ["ok=1 n=1", "[only] 0 "]
    out, _ = run(b">x\r\nAC\r\n\r\nGT\r\n")                   # CR LF line ends: the header keeps its '\r' like the reference's getline does
    assert out == ["ok=1 n=1", "[x", "] 4 ACGT"]           # (splitlines() cuts at the name's '\r')
    big = b">a\n" + b"ACGT" * 50_000 + b"\n" + b"".join(b">s%d\n%s\n" % (i, b"GATTACA" * (i + 1)) for i in range(40))
    out, _ = run(big)
    assert out[0] == "ok=1 n=41" and out[1].startswith("[a] 200000 ACGTACGT") and out[41] == "[s39] 280 " + "GATTACA" * 40
    # a file large enough (> 4 MB per slice) for the record-start scan to be cut into slices on several threads: records of
    # every size, so that slice borders fall inside names, inside sequence lines, on the '>' itself and on the newline before it
    import hashlib
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    recs, parts = [], []
    for i in range(700):
        L = int(rng.integers(0, 120_000)) if i % 50 else int(rng.integers(2_000_000, 5_000_000))
        seq = acgt[rng.integers(0, len(acgt), size=L)].tobytes()
        w = int(rng.integers(50, 90))
        recs.append((f"r{i}", seq))
        parts.append(b">r%d with > inside the comment\n" % i + b"\n".join(seq[k:k + w] for k in range(0, L, w)) + (b"\n" if L else b""))
    f = tmp_path / "big.fa"
    f.write_bytes(b"".join(parts))
    assert f.stat().st_size > 40_000_000
    src2 = tmp_path / "ld2.cpp"
    src2.write_text(_LOADER_HARNESS.replace('printf("[%s] %zu %s\\n", c.name.c_str(), c.seq.size(), c.seq.c_str());',
                                            '{ unsigned long long h = 1469598103934665603ull; for (char ch : c.seq) h = (h ^ (unsigned char)ch) * 1099511628211ull; printf("[%s] %zu %llx\\n", c.name.c_str(), c.seq.size(), h); }'))
    exe2 = str(tmp_path / "ld2")
    subprocess.run(["g++", "-O2", f"-std={std}", "-I", host, "-o", exe2, str(src2), os.path.join(host, "io.cpp"), "-lpthread"], check=True)
    env2 = dict(os.environ, **({"GSA_FASTA_PIECE": "100003"} if piece else {}))
    got = subprocess.run([exe2, str(f)], capture_output=True, text=True, env=env2).stdout.splitlines()

    def fnv(b):
        h = 1469598103934665603
        for x in np.frombuffer(b, dtype=np.uint8).tolist():
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    assert got[0] == f"ok=1 n={len(recs)}"
    for line, (name, seq) in zip(got[1:], recs):
        nm, ln, hx = line.split()
        assert nm == f"[{name}]" and int(ln) == len(seq), (line, name, len(seq))
    # hashing 60 MB byte by byte in Python is slow: check the hash of every 10th record and of the short ones
    for line, (name, seq) in list(zip(got[1:], recs))[::10]:
        if len(seq) < 200_000:
            assert int(line.split()[2], 16) == fnv(seq), name


def test_record_walk_host_only():
    """gsa_record_next is host code: walks a hand-built outbox image (two records) and rejects a truncated one"""
    import ctypes as C
    from gsalign_b200 import capi
    lib = capi.load_library()

    def pad(b):
        return b + b"\0" * (-len(b) % 16)
    blocks = np.zeros(2, dtype=capi.BLOCK_DTYPE); blocks["score"] = [7, 9]; blocks["n_frags"] = [1, 2]
    frags = np.zeros(3, dtype=capi.FRAG_DTYPE); frags["qPos"] = [5, 6, 7]
    a1, a2 = b"AC-GT", b"ACTGT"
    rec0 = np.array([3, 2, 3, 5], dtype=np.int64).tobytes() + pad(blocks.tobytes()) + pad(frags.tobytes()) + pad(a1) + pad(a2)
    rec1 = np.array([8, 0, 0, 0], dtype=np.int64).tobytes()
    img = rec0 + rec1
    buf = C.create_string_buffer(img, len(img))
    off, contig, al = C.c_int64(0), C.c_int64(), capi.Alignment()
    assert lib.gsa_record_next(buf, C.c_int64(len(img)), C.byref(off), C.byref(contig), C.byref(al)) == 1
    assert contig.value == 3 and al.n_blocks == 2 and al.n_frags == 3 and al.aln_bytes == 5
    assert C.string_at(al.aln1, 5) == a1 and C.string_at(al.aln2, 5) == a2
    got = np.frombuffer(C.string_at(al.frags, 3 * capi.FRAG_DTYPE.itemsize), dtype=capi.FRAG_DTYPE)
    assert list(got["qPos"]) == [5, 6, 7]
    assert lib.gsa_record_next(buf, C.c_int64(len(img)), C.byref(off), C.byref(contig), C.byref(al)) == 1
    assert contig.value == 8 and al.n_blocks == 0
    assert lib.gsa_record_next(buf, C.c_int64(len(img)), C.byref(off), C.byref(contig), C.byref(al)) == 0
    off = C.c_int64(0)
    assert lib.gsa_record_next(buf, C.c_int64(len(rec0) - 16), C.byref(off), C.byref(contig), C.byref(al)) < 0


def test_compact_record_expansion_host_only():
    """gsa_record_frags is host code: a hand-built record in the compact form (8 bytes per fragment + anchors) expands to the
    fragment records the rule in gather.cu defines -- positions chained from the anchors, row offsets from the running row
    slots, gapped alignments right-aligned in a slot of qLen + rLen -- and malformed anchors are rejected"""
    import ctypes as C
    from gsalign_b200 import capi
    lib = capi.load_library()

    def pad(b):
        return b + b"\0" * (-len(b) % 16)
    # (bSeed, qLen, rLen, aln_len, gapped) ; a second block restarts the positions at fragment 5
    spec = [(1, 20, 20, 20, 0), (0, 3, 5, 6, 1), (1, 10, 10, 10, 0), (0, 0, 4, 4, 0), (1, 7, 7, 7, 0),
            (1, 15, 15, 15, 0), (0, 2, 2, 2, 0), (1, 9, 9, 9, 0), (0, 6, 0, 6, 0), (1, 30, 30, 30, 0)]
    anchors = [(0, 1000, 50, 0), (5, 70_000_000_000, 400, 12)]     # (first fragment, rPos, qPos, row base); rows before fragment 5: 8 + 4
    want = np.zeros(len(spec), dtype=capi.FRAG_DTYPE)
    k = -1
    for i, (seed, ql, rl, al_, gp) in enumerate(spec):
        if k + 1 < len(anchors) and anchors[k + 1][0] == i:
            k += 1
            r, q, row = anchors[k][1], anchors[k][2], anchors[k][3]
        slot = 0 if seed else (ql + rl if gp else (rl if ql == 0 else ql))
        typ = 0 if seed else (1 if ql == 0 else 2 if rl == 0 else 4 if gp else 3)
        want[i] = (r, q, ql, rl, seed, 0 if seed else row + (slot - al_ if gp else 0), al_, typ)
        r += rl; q += ql; row += slot
    cfrag = np.array([ql | (rl << 21) | (al_ << 42) | (seed << 62) | (gp << 63) for seed, ql, rl, al_, gp in spec], dtype=np.uint64)
    anc = b"".join(np.array([f, r], dtype=np.int64).tobytes() + np.array([q, 0], dtype=np.int32).tobytes() + np.array([row], dtype=np.int64).tobytes() for f, r, q, row in anchors)
    blocks = np.zeros(2, dtype=capi.BLOCK_DTYPE); blocks["n_frags"] = [5, 5]; blocks["frag_beg"] = [0, 5]
    rows = 8 + 4 + 2 + 6
    a1, a2 = bytes(range(65, 65 + rows)), bytes(range(97, 97 + rows))

    def image(anchor_bytes, n_anchor):
        return (np.array([-1 - 4, 2, len(spec), rows], dtype=np.int64).tobytes() + np.array([n_anchor, 0, 0, 0], dtype=np.int64).tobytes() +
                pad(blocks.tobytes()) + pad(cfrag.tobytes()) + anchor_bytes + pad(a1) + pad(a2))
    img = image(anc, 2) + np.array([9, 0, 0, 0], dtype=np.int64).tobytes()       # a plain, empty record follows
    buf = C.create_string_buffer(img, len(img))
    recs = capi.walk_image(lib, buf, len(img))
    assert [r[0] for r in recs] == [4, 9]
    assert recs[0][2].tobytes() == want.tobytes()
    assert recs[0][3].tobytes() == a1 and recs[0][4].tobytes() == a2 and list(recs[0][1]["frag_beg"]) == [0, 5]
    # the first anchor must sit on fragment 0, anchors must be in order
    for bad in (anc[32:] + anc[:32], anc.replace(np.array([5], dtype=np.int64).tobytes(), np.array([50], dtype=np.int64).tobytes(), 1)):
        b = image(bad, 2)
        bb = C.create_string_buffer(b, len(b))
        out = np.zeros(len(spec), dtype=capi.FRAG_DTYPE)
        assert lib.gsa_record_frags(bb, C.c_int64(len(b)), C.c_int64(0), out.ctypes.data_as(C.c_void_p), C.c_int32(1)) < 0


def test_block_logic_array_form_and_sort_restatement(tmp_path):
    """The block logic also runs in a kernel (block_logic.cuh on plain arrays, with libstdc++'s std::sort restated in
    stdsort.cuh because the order it leaves ties in is observable).  Compiled for the host, both are pinned here: the sort
    against std::sort (tie-heavy, sorted, reversed and median-of-three-killer inputs), the split / dedup logic against the
    std::vector + std::sort form of block_logic.cpp on 3 000 random block lists (duplicates, overlaps, both strands, -one)."""
    csrc = os.path.join(ROOT, "gsalign_b200", "csrc")
    exe = str(tmp_path / "blocklogic_harness")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-I", csrc, "-I", "/usr/local/cuda/include", "-o", exe,
                    os.path.join(ROOT, "tests", "blocklogic_harness.cpp"), os.path.join(csrc, "block_logic.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("same"), out.stdout + out.stderr


def test_header_is_plain_c(tmp_path):
    """include/gsalign_b200.h is the drop-in boundary: it must compile as C99 on its own (no C++ or CUDA types in the signatures)"""
    src = tmp_path / "hdr.c"
    src.write_text('#include "gsalign_b200.h"\nint main(void) { gsa_variant v; gsa_variant_list l; gsa_alignment a; (void)v; (void)l; (void)a;\n'
                   '  return (int)sizeof(gsa_frag) - 40 + (int)sizeof(gsa_block) - 24 + (int)sizeof(gsa_variant) - 24; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "hdr"), str(src)], check=True)
    assert subprocess.run([str(tmp_path / "hdr")]).returncode == 0   # the record sizes the image formats rely on


def test_compact_record_expansion_threads_agree():
    """a long compact record (60 000 fragments, an anchor every 256 and at every block start): gsa_record_frags spreads the
    stretches between anchors over host threads; 1 and 4 threads must write the same fragment list, and the list must chain"""
    import ctypes as C
    from gsalign_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    nf = 60_000
    seed = (np.arange(nf) % 2 == 0).astype(np.uint64)
    ql = np.where(seed == 1, rng.integers(15, 200, nf), rng.integers(0, 30, nf)).astype(np.uint64)
    rl = np.where(seed == 1, ql, rng.integers(0, 30, nf)).astype(np.uint64)
    rl[(seed == 0) & (ql == 0) & (rl == 0)] = 3
    gp = ((seed == 0) & (ql > 0) & (rl > 0) & (rng.random(nf) < 0.5)).astype(np.uint64)
    slot = np.where(seed == 1, 0, np.where(gp == 1, ql + rl, np.where(ql == 0, rl, ql))).astype(np.int64)
    al_ = np.where(seed == 1, ql, np.where(gp == 1, np.maximum(ql, rl), slot)).astype(np.uint64)
    cfrag = (ql | (rl << np.uint64(21)) | (al_ << np.uint64(42)) | (seed << np.uint64(62)) | (gp << np.uint64(63))).astype(np.uint64)
    starts = sorted(set(range(0, nf, 256)) | {int(x) for x in rng.integers(1, nf, 40)})     # block starts restart the positions
    qpos, rpos, row = np.zeros(nf, np.int64), np.zeros(nf, np.int64), np.concatenate([[0], np.cumsum(slot)[:-1]])
    q = r = 0
    restart = set(starts) - set(range(0, nf, 256))
    for i in range(nf):
        if i in restart:
            q, r = int(rng.integers(0, 1 << 27)), int(rng.integers(0, 1 << 33))
        qpos[i], rpos[i] = q, r
        q += int(ql[i]); r += int(rl[i])
    anc = b"".join(np.array([i, rpos[i]], dtype=np.int64).tobytes() + np.array([qpos[i], 0], dtype=np.int32).tobytes() + np.array([row[i]], dtype=np.int64).tobytes() for i in starts)
    ab = int(slot.sum())
    pad = lambda b: b + b"\0" * (-len(b) % 16)
    img = (np.array([-1 - 0, 0, nf, ab], dtype=np.int64).tobytes() + np.array([len(starts), 0, 0, 0], dtype=np.int64).tobytes() +
           pad(cfrag.tobytes()) + anc + pad(b"x" * ab) + pad(b"y" * ab))
    buf = C.create_string_buffer(img, len(img))
    outs = []
    for nt in (1, 4):
        out = np.zeros(nf, dtype=capi.FRAG_DTYPE)
        assert lib.gsa_record_frags(buf, C.c_int64(len(img)), C.c_int64(0), out.ctypes.data_as(C.c_void_p), C.c_int32(nt)) == 0
        outs.append(out)
    assert outs[0].tobytes() == outs[1].tobytes()
    got = outs[0]
    assert np.array_equal(got["qPos"], qpos) and np.array_equal(got["rPos"], rpos) and np.array_equal(got["qLen"], ql.astype(np.int32))
    gaps = got["bSeed"] == 0
    assert np.array_equal(got["aln_off"][gaps], (row + np.where(gp == 1, slot - al_.astype(np.int64), 0))[gaps]) and not got["aln_off"][~gaps].any()
