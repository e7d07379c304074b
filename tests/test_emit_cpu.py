"""Host logic on CPU: the CLI's MAF / ALN / VCF emitters (gsalign_b200/csrc/host/emit.cpp, multi-threaded) fed with the
alignment records of the UNMODIFIED reference (its own containers, through oracle/_ref/libgsref.so) must write the files the
reference CLI writes, byte for byte.  No GPU involved: this pins the emitters independently of the kernels."""
import os
import struct
import subprocess

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(ROOT, "gsalign_b200", "csrc", "host")

FRAG = np.dtype([("rPos", "<i8"), ("qPos", "<i4"), ("qLen", "<i4"), ("rLen", "<i4"), ("bSeed", "<i4"), ("aln_off", "<i8"), ("aln_len", "<i4"), ("reserved", "<i4")])
BLOCK = np.dtype([("score", "<i4"), ("aln_len", "<i4"), ("bDup", "<i4"), ("n_frags", "<i4"), ("frag_beg", "<i8")])


def _records(ref_contig):
    """one contig of tests/ref_worker.py's result -> the byte stream tests/emit_harness.cpp reads"""
    ref_aln, k = {}, 0
    for b in ref_contig["stages"][4]:            # aln strings come in stage-4 order (before the identity filter)
        for f in b[3]:
            if f[0] == 0:
                ref_aln[(f[1], f[2], f[3], f[4])] = (ref_contig["aln"][k], ref_contig["aln"][k + 1]); k += 2
    final = ref_contig["stages"][5]
    blocks = np.zeros(len(final), dtype=BLOCK)
    frags, a1, a2 = [], bytearray(), bytearray()
    for bi, (score, aln_len, dup, fr) in enumerate(final):
        blocks[bi] = (score, aln_len, dup, len(fr), len(frags))
        for (seed, q, r, ql, rl) in fr:
            if seed:
                frags.append((r, q, ql, rl, 1, 0, ql, 0))
            else:
                x, y = ref_aln[(q, r, ql, rl)]
                assert len(x) == len(y)
                frags.append((r, q, ql, rl, 0, len(a1), len(x), 0))
                a1 += x; a2 += y
    fa = np.array(frags, dtype=FRAG) if frags else np.zeros(0, dtype=FRAG)
    return struct.pack("<i", len(final)) + blocks.tobytes() + struct.pack("<q", len(fa)) + fa.tobytes() + struct.pack("<q", len(a1)) + bytes(a1) + bytes(a2)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emit") / "emit_harness")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", HOST, "-o", exe, os.path.join(HERE, "emit_harness.cpp"), os.path.join(HOST, "emit.cpp"),
                    os.path.join(HOST, "io.cpp"), "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("fmt,threads,prm,flags,dev_vars", [(1, 1, {}, [], False), (1, 5, {}, [], False), (2, 3, {}, ["-fmt", "2"], False),
                                                            (1, 4, dict(min_seed_len=10, sensitive=1, min_block_score=50), ["-sen"], False),
                                                            (1, 1, {}, [], True), (1, 6, {}, [], True),
                                                            (1, 3, dict(min_seed_len=10, sensitive=1, min_block_score=50), ["-sen"], True)])
def test_emitters_match_reference_cli(workdir, harness, fmt, threads, prm, flags, dev_vars):
    """dev_vars: the harness derives the variant records gsa_variants() delivers (rule of include/gsalign_b200.h) and the
    emitters take the single-GPU CLI's route (alleles fetched per record, written in place by all threads) instead of
    scanning the rows."""
    from conftest import build_index
    from test_gpu_pipeline import make_rearranged, run_reference
    exe = os.path.join(orc.REF_DIR, "GSAlign")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/GSAlign not built")
    d = make_rearranged(workdir)
    prefix, qry = os.path.join(d, "ref"), os.path.join(d, "qry.fa")
    build_index(os.path.join(d, "ref.fa"), prefix)
    tag = f"{fmt}_{threads}_{'sen' if prm else 'def'}_{int(dev_vars)}"
    ref = run_reference(prefix, qry, os.path.join(d, f"emit_ref_{tag}.pkl"), **prm)
    rec = os.path.join(d, f"records_{tag}.bin")
    with open(rec, "wb") as f:
        for c in ref:
            f.write(_records(c))
    ours, theirs = os.path.join(d, f"emit_ours_{tag}"), os.path.join(d, f"emit_ref_{tag}")
    env = dict(os.environ, GSA_EMIT_CHUNK="40")   # small chunks: the blocks here have a few thousand fragments, the threads must all get some
    if dev_vars:
        env["GSA_HARNESS_VARS"] = "1"
    subprocess.run([harness, prefix, qry, rec, ours, str(fmt), str(threads)], check=True, stderr=subprocess.DEVNULL, env=env)
    subprocess.run([exe, "-t", "1", "-i", prefix, "-q", qry, "-o", theirs] + flags, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ext = ".maf" if fmt == 1 else ".aln"
    for e in (ext, ".vcf"):
        with open(ours + e, "rb") as a, open(theirs + e, "rb") as b:
            assert a.read() == b.read(), e
    assert os.path.getsize(ours + ".vcf") > 1000 and os.path.getsize(ours + ext) > 100_000


@pytest.mark.parametrize("n,distinct,threads", [(0, 1, 4), (17, 3, 4), (300_000, 7, 1), (300_000, 1000, 8), (1_000_000, 50_000, 16), (400_000, 1, 5)])
def test_threaded_sort_moves_like_std_sort(harness, n, distinct, threads):
    """output_variants sorts (chr, pos) keys whose ties keep whatever order libstdc++'s introsort leaves (hazard H5): the
    threaded sort must produce std::sort's exact permutation, ties included."""
    out = subprocess.run([harness, "sort", str(n), str(distinct), str(threads), "7"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "same", out.stdout


@pytest.mark.parametrize("pattern", [0, 1, 2, 3])
@pytest.mark.parametrize("par_min", [64, 5000, 200_000])
@pytest.mark.parametrize("n,distinct,threads", [(100_000, 50, 8), (700_000, 30_000, 5), (1_500_000, 100_000_000, 16), (5000, 5, 3)])
def test_threaded_partition_swaps_like_std_sort(harness, n, distinct, threads, par_min, pattern):
    """Long ranges are partitioned by several threads (partition_threads: stoppers counted per chunk, the cut found from the
    sums, the same pairs swapped).  GSA_SORT_PAR_MIN lowers the range length from which that happens so that every level of
    a small sort goes through it; the inputs are random keys, ascending runs with repeated neighbours (the shape
    VariantIdentification pushes), descending and sorted keys.  The permutation must be std::sort's."""
    env = dict(os.environ, GSA_SORT_PAR_MIN=str(par_min))
    for seed in (1, 2):
        out = subprocess.run([harness, "sort", str(n), str(distinct), str(threads), str(seed), str(pattern)], capture_output=True, text=True, env=env)
        assert out.returncode == 0 and out.stdout.strip() == "same", (out.stdout, seed)


@pytest.mark.parametrize("contigs,bp,snv,indel,ext,fmt,min_bytes", [(3, 10_000_000, "0.01", "0.001", 0, 1, 10_000_000), (2, 3_000_000, "0.05", "0.03", 0, 1, 5_000_000),
                                                                     (2, 4_000_000, "0.10", "0", 0, 1, 5_000_000), (3, 2_000_000, "0.01", "0.002", 37, 1, 5_000_000),
                                                                     (3, 2_000_000, "0.01", "0.002", 205, 2, 5_000_000), (2, 6_000_000, "0.02", "0.004", 0, 2, 10_000_000)])
def test_fabricated_contigs_threads_and_variant_routes_same_bytes(harness, tmp_path, contigs, bp, snv, indel, ext, fmt, min_bytes):
    """tools/emit_rig.cpp fabricates the records of contig pairs at the BASELINE rates (one block of ~185 000 fragments per
    10 Mbp contig, indels included), at an indel-dense setting (neighbouring events merge into long gapped fragments) and at
    C5's divergence; with blocks that run past the end of their reference contig (iExtension trims them: the rows get a NUL
    that cuts the printed text, src/tools.cpp:192-202) and in the .aln format (80-column windows formatted by all threads):
    the files must not depend on the thread count, nor on whether the variants come from
    the row scan or from derived device-style records -- at chunk sizes that give every thread several stretches."""
    rig = str(tmp_path / "emit_rig")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", rig, os.path.join(ROOT, "tools", "emit_rig.cpp")], check=True)
    d = str(tmp_path)
    subprocess.run([rig, d, str(contigs), str(bp), snv, indel, "9", str(ext)], check=True, stderr=subprocess.DEVNULL)
    outs = []
    for tag, threads, env in (("t1", 1, {}), ("t8", 8, {"GSA_EMIT_CHUNK": "3000"}), ("t5v", 5, {"GSA_EMIT_CHUNK": "3000", "GSA_HARNESS_VARS": "1"}),
                              ("t16v", 16, {"GSA_HARNESS_VARS": "1"})):
        out = os.path.join(d, tag)
        subprocess.run([harness, os.path.join(d, "ref"), os.path.join(d, "qry.fa"), os.path.join(d, "records.bin"), out, str(fmt), str(threads)],
                       check=True, stderr=subprocess.DEVNULL, env=dict(os.environ, **env))
        outs.append(out)
    for e in (".maf" if fmt == 1 else ".aln", ".vcf"):
        want = open(outs[0] + e, "rb").read()
        assert len(want) > (min_bytes if e != ".vcf" else 100_000)
        for o in outs[1:]:
            assert open(o + e, "rb").read() == want, (o, e)
