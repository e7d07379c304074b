"""debug aid: where do two BWA index file sets differ (header fields, first differing word)"""
import sys
import numpy as np
a, b = sys.argv[1], sys.argv[2]
for ext in ("bwt", "sa"):
    x = np.fromfile(f"{a}.{ext}", dtype=np.uint32); y = np.fromfile(f"{b}.{ext}", dtype=np.uint32)
    hx = x[:14].view(np.uint64); hy = y[:14].view(np.uint64)
    print(ext, "sizes", x.shape[0], y.shape[0], "hdr", hx[:7].tolist(), hy[:7].tolist())
    m = min(x.shape[0], y.shape[0])
    d = np.flatnonzero(x[:m] != y[:m])
    print(ext, "differing words", d.shape[0], "first", d[:10].tolist(), "last", d[-3:].tolist())
    if d.shape[0]:
        i = int(d[0]); print("  ", hex(int(x[i])), hex(int(y[i])))
