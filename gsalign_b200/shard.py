"""Query-contig sharding over GPUs: contigs are independent units, so they are dealt
longest-first to the least loaded rank; no collective sits on the data path.  Mirrors the dealing in
bin/GSAlign (-gpus N, gsalign_b200/csrc/host/main.cpp)."""
from __future__ import annotations


def lpt_assign(lengths, n_ranks: int):
    """Returns a list of n_ranks lists of contig indices (each ascending), longest-processing-time first."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * n_ranks
    out = [[] for _ in range(n_ranks)]
    for i in order:
        g = min(range(n_ranks), key=lambda r: (load[r], r))
        out[g].append(i)
        load[g] += lengths[i]
    return [sorted(x) for x in out]
