#!/usr/bin/env python
"""profiles/sass_summary.txt: per kernel of 3dvision_b200/libb3d.so, the counts of the SASS mnemonics that show what the
kernel is built from on sm_100a — UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk = TMA bulk copy),
SYNCS (mbarrier), FFMA2/FADD2/FMUL2 (packed fp32), FFMA/FADD/FMUL, SHFL, VOTE, ATOM/RED, DADD/DFMA/DMUL (fp64).
Usage: python scripts/sass_summary.py > profiles/sass_summary.txt      (cuobjdump -sass; needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3dvision_b200", "libb3d.so")
WANT = ["UTCHMMA", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "SHFL", "VOTE", "ATOM", "RED", "DADD", "DFMA", "DMUL", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for w in WANT:
                if op == w or (w in ("ATOM", "RED") and op.startswith(w)):
                    cur[w] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, architectures {arch}")
    print("# counts of SASS instructions per kernel (static code, not executed counts); columns with no hit anywhere are omitted")
    cols = [w for w in WANT if any(k[w] for k in kernels.values())]
    print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in cols))
    tot = collections.Counter()
    for (mangled, cnt), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*", "", name).replace("b3d::", "").replace("void ", "")[:78]
        print(f"{short:78s} {cnt['_total']:6d} " + " ".join(f"{cnt[c]:7d}" for c in cols))
        tot.update(cnt)
    print(f"{'TOTAL':78s} {tot['_total']:6d} " + " ".join(f"{tot[c]:7d}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
