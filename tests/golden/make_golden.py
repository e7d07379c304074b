"""Generates tests/golden/ecoli_golden.json from the UNMODIFIED reference (oracle/_ref/libgsref.so) run in this
container on config C1 (test/ecoli.fa vs test/ecoli.mut).  The file pins the oracle (and through it the CUDA path)
where /root/reference is absent.  Usage: python tests/golden/make_golden.py <index prefix> <query fasta>"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def digest_blocks(blocks):
    h = hashlib.sha256()
    for b in sorted(blocks, key=lambda b: (b[0], b[3])):
        h.update(repr((b[0], b[3])).encode())
    return h.hexdigest()


def main():
    import numpy as np
    import orc
    from gsalign_b200 import synth
    prefix, qpath = sys.argv[1:3]
    R = orc.Reference(prefix)
    seq = synth.read_fasta(qpath)[0][1].tobytes()
    q, r, l = R.seed_contig(seq)
    stages, aln = R.cluster()
    g = {"n_seeds": int(len(q)),
         "seeds_sha256": hashlib.sha256(q.astype("<i4").tobytes() + r.astype("<i8").tobytes() + l.astype("<i4").tobytes()).hexdigest(),
         "stage_blocks": {str(s): len(stages[s]) for s in range(6)},
         "stage_sha256": {str(s): digest_blocks(stages[s]) for s in range(4)},
         "final": [[b[0], b[1], b[2], len(b[3])] for b in stages[5]],
         "aln_sha256": hashlib.sha256(b"\n".join(aln)).hexdigest(), "n_aln": len(aln)}
    # known-answer DP vectors straight from ksw2_alignment
    import random
    random.seed(5)
    dp = []
    for t in range(40):
        m = random.randint(1, 40); a = "".join(random.choice("ACGT") for _ in range(m))
        b = "".join((random.choice("ACGTN") if random.random() < 0.15 else c) for c in a)
        if t % 2:
            c = random.randrange(len(b)); b = b[:c] + b[c + random.randint(1, 4):] or "A"
        x, y = R.ksw2(a.encode(), b.encode())
        dp.append([a, b, x.decode(), y.decode()])
    g["dp_vectors"] = dp
    # known-answer single searches
    ks = []
    for start in (0, 17, 1000, 123456, 2000000, 4639000):
        stop = min(len(seq), (start // 10000 + 1) * 10000)
        ln, fq, loc = R.bwt_search(seq, start, stop)
        ks.append([start, stop, ln, fq, sorted(loc)])
    g["search_vectors"] = ks
    json.dump(g, open(os.path.join(HERE, "ecoli_golden.json"), "w"), indent=1)
    print("wrote ecoli_golden.json", g["n_seeds"], g["stage_blocks"])


if __name__ == "__main__":
    main()
