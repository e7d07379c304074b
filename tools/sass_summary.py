#!/usr/bin/env python
"""profiles/<tag>_sass.md from `cuobjdump -sass gsalign_b200/libgsalign_b200.so`: per kernel, the count of the mnemonics that prove
the DPX path (VIMNMX.S16x2, VIADD.16x2), the TMA bulk copies (UBLKCP) and their mbarrier (SYNCS.*), plus short excerpts.
usage: python tools/sass_summary.py <tag>     (no GPU needed)"""
import collections, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(root, "gsalign_b200", "libgsalign_b200.so")], capture_output=True, text=True).stdout
keep = {}
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", f)
    keep[name] = (len(ops), collections.Counter(ops), f)
names = list(keep)
dm = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")))
out = [f"# profiles/{tag}: SASS evidence (cuobjdump -sass gsalign_b200/libgsalign_b200.so, sm_100a only; tools/sass_summary.py)\n",
       "Instruction counts per kernel of the built library: the packed-int16 DPX path (`VIMNMX.S16x2` with two predicate outputs,",
       "`VIADD.16x2`), TMA bulk copies (`UBLKCP.S.G`) with their mbarrier (`SYNCS.ARRIVE.TRANS64`, `SYNCS.PHASECHK.TRANS64.TRYWAIT`),",
       "and the acquire / release hand-over between the strips of `k_dpx<W>` (`.STRONG` loads / stores on the progress counters).\n",
       "| kernel | SASS instructions | VIMNMX.S16x2 | VIADD.16x2 | PRMT | SHFL | UBLKCP | SYNCS.* | LD/ST .STRONG | LDG | STG |", "|---|---|---|---|---|---|---|---|---|---|---|"]


def cnt(c, pat):
    return sum(v for k, v in c.items() if re.search(pat, k))


for name, (n, c, f) in sorted(keep.items(), key=lambda x: dm[x[0]]):
    d = dm[name]
    if not re.search(r"k_seed<|k_dpx<|k_dpx_pack<(16|1),|k_variants|k_gap_similarity|k_chain<FGroup>|k_copy_words", d):
        continue
    d = re.sub(r"\(.*", "", d)
    out.append(f"| `{d}` | {n} | {cnt(c, '^VIMNMX.*16x2')} | {cnt(c, '^VIADD.*16x2')} | {cnt(c, '^PRMT')} | {cnt(c, '^SHFL')} | {cnt(c, '^UBLKCP')} | "
               f"{cnt(c, '^SYNCS')} | {cnt(c, 'STRONG')} | {cnt(c, '^LDG')} | {cnt(c, '^STG')} |")
strip = lambda l: re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip()
for name, (n, c, f) in keep.items():
    d = dm[name]
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    if d.startswith("void k_seed<true>"):
        idx = [i for i, l in enumerate(lines) if "UBLKCP" in l or "SYNCS" in l]
        out.append("\n## `k_seed<true>`: the chunk's query staged by two bulk copies on the warp's mbarrier\n\n```")
        out += [strip(l) for l in lines[max(0, idx[0] - 3):idx[4] + 2]]
        out.append("```")
    if d.startswith("void k_dpx<16, 1, false, false>"):
        idx = [i for i, l in enumerate(lines) if "VIMNMX.S16x2" in l]
        out.append("\n## `k_dpx<16,1,false,false>`: one wavefront step (two cells per lane): shuffle, score permute, 4 x VIADD.16x2 + 4 x VIMNMX.S16x2 "
                   "with predicate outputs, predicated flag ORs\n\n```")
        out += [strip(l) for l in lines[idx[0] - 10:idx[4] + 4]]
        out.append("```")
        idx = [i for i, l in enumerate(lines) if "STRONG" in l]
        out.append("\n## `k_dpx<16,...>`: strip hand-over (acquire load in the spin on the previous strip's progress, release store of the own)\n\n```")
        for i in idx[:4]:
            out += [strip(l) for l in lines[max(0, i - 2):i + 3]] + ["        ..."]
        out.append("```")
open(os.path.join(root, "profiles", f"{tag}_sass.md"), "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_sass.md")
