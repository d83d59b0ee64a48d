#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): registers / stack / shared memory per kernel
(`cuobjdump --dump-resource-usage`) and how often the instructions the design relies on appear in each
kernel's SASS (DPX max-plus forms, TMA bulk copies + mbarrier waits, uniform-datapath constant loads,
popcounts, FP64).  python tools/sass_census.py > profiles/<round>_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scoary_b200", "libscoary_b200.so")
WATCH = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "VIADD", "UBLKCP", "SYNCS", "LDCU", "LDC", "ULDC", "POPC", "REDUX", "DFMA",
         "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BRA.U", "BRX", "SHFL", "IMAD", "LOP3", "PRMT"]


def demangle(n):
    m = re.search(r"\d+sb(\d+)", n)
    return n[m.end():m.end() + int(m.group(1))] + ("<" + ",".join(re.findall(r"L[bi](\d)E", n)) + ">" if "IL" in n else "") if m else n


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        line = line.strip()
        if line.startswith("Function"):
            name = demangle(line.split()[1].rstrip(":"))
        elif line.startswith("REG:") and name:
            usage[name] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, total = collections.defaultdict(collections.Counter), collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
                    break
    print("# static census of scoary_b200/libscoary_b200.so (sm_100a); built by tools/sass_census.py")
    print("%-32s %4s %5s %6s %6s  %s" % ("kernel", "regs", "stack", "smem", "instrs", "watched instructions (count)"))
    for k in sorted(total, key=lambda k: -total[k]):
        if k.startswith("pipe_rate") or k.startswith("int32_peak"):
            continue
        u = usage.get(k, {})
        print("%-32s %4s %5s %6s %6d  %s" % (k[:32], u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), total[k],
                                          " ".join("%s=%d" % kv for kv in counts[k].most_common())))


if __name__ == "__main__":
    sys.exit(main())
