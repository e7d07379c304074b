"""Runs the UNMODIFIED reference (oracle/_ref/libgsref.so) on one index + query in a fresh process
(the reference keeps its state in globals) and pickles every intermediate.  Test infrastructure only."""
import pickle
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))


def main():
    prefix, query_path, out_path = sys.argv[1:4]
    prm = dict(kv.split("=") for kv in sys.argv[4:])
    prm = {k: int(v) for k, v in prm.items()}
    import orc
    from gsalign_b200 import synth
    R = orc.Reference(prefix)
    R.set_params(**prm)
    res = []
    for name, seq in synth.read_fasta(query_path):
        q, r, l = R.seed_contig(seq.tobytes())
        stages, aln = R.cluster()
        res.append({"name": name, "seeds": (q, r, l), "stages": stages, "aln": aln})
    with open(out_path, "wb") as f:
        pickle.dump(res, f)


if __name__ == "__main__":
    main()
