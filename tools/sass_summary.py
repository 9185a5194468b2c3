#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libmobocmf_b200.so (cuobjdump -sass): the instructions that show which hardware
paths a kernel uses - DMMA (FP64 tensor instruction, mma.sync.m8n8k4.f64), UBLKCP (bulk-copy / TMA engine), SYNCS
(mbarrier), LDGSTS (cp.async), BAR (named barriers), USETMAXREG (setmaxnreg), DFMA/DMUL/DADD (FP64 pipe), LDL/STL
(spills).   python tools/sass_summary.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mobocmf_b200", "lib", "libmobocmf_b200.so")
OPS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "BAR", "USETMAXREG", "LDS", "STS", "LDG", "STG",
       "LDL", "STL", "SHFL", "UTMALDG", "UTCQMMA", "UTCHMMA", "LDTM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts, order, cur, k = {}, [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[k]).replace("mobo::", "")
            k += 1
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op.split(".")[0]] += 1
            counts[cur]["_total"] += 1
    print("# %s: SASS instruction counts per kernel (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-34s %7s " % ("kernel", "total") + " ".join("%6s" % o[:6] for o in OPS))
    for name in sorted(order, key=lambda n: -counts[n]["_total"]):
        c = counts[name]
        print("%-34s %7d " % (name[:34], c["_total"]) + " ".join("%6d" % c[o] for o in OPS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("%-34s %7d " % ("ALL", tot["_total"]) + " ".join("%6d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
