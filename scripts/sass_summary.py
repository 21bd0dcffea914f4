#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor load),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus FFMA / HMMA for contrast.

    python scripts/sass_summary.py > profiles/sass_summary.txt

Runs in the build container (cuobjdump reads the in-tree libb200knn.so, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
LIB = os.path.join(ROOT, "self-supervised-wafermaps_b200", "lib", "libb200knn.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "REDUX", "FFMA", "HMMA", "DFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels; columns = instruction counts in the SASS of each kernel")
    print("# " + " ".join(f"{p:>8}" for p in ["total"] + PAT) + "  kernel")
    agg = collections.Counter()
    for fn, c in counts.items():
        short = re.sub(r"b200knn::\(anonymous namespace\)::", "", names.get(fn, fn))
        short = re.sub(r"\(CUtensorMap_st.*", "", short)
        short = re.sub(r"\(b200knn::.*", "", short)
        print("  " + " ".join(f"{c.get(p, 0):>8}" for p in ["total"] + PAT) + "  " + short[:150])
        agg.update(c)
    print("# sum " + " ".join(f"{agg.get(p, 0):>8}" for p in ["total"] + PAT))


if __name__ == "__main__":
    sys.exit(main())
