#!/usr/bin/env python
"""SASS opcode census of libxesn_b200.so, per kernel: which kernels use tcgen05 (UTCIMMA), TMA (UTMALDG), TMEM loads
(LDTM), tcgen05 commit barriers (UTCBAR), FP64 tensor MMA (DMMA), mbarriers (SYNCS).
    python tools/sass_census.py > profiles/sass_census_rNN.txt        (CPU only: cuobjdump on the built library)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "xesn_b200", "libxesn_b200.so")
OPS = ("UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "DMMA", "SYNCS", "MUFU", "DFMA", "DADD", "DMUL")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows, name, cnt, total = [], None, collections.Counter(), 0
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in sass.splitlines():
        if "Function :" in line:
            if name:
                rows.append((name, cnt, total))
            name, cnt, total = line.split("Function :")[1].strip(), collections.Counter(), 0
            continue
        m = ins.match(line)
        if m and name:
            total += 1
            op = m.group(1).split(".")[0]
            if op in OPS:
                cnt[op] += 1
    if name:
        rows.append((name, cnt, total))
    demangle = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode census of libxesn_b200.so (cuobjdump -sass, sm_100a), {len(rows)} kernels")
    print("# " + " | ".join(("kernel",) + OPS + ("instructions",)))
    tot = collections.Counter()
    for (nm, c, t), dm in sorted(zip(rows, demangle), key=lambda x: (-x[0][1]["UTCIMMA"], -x[0][1]["DMMA"], x[1])):
        short = re.sub(r"\(anonymous namespace\)::", "", dm)
        short = re.sub(r"\(.*", "", short)
        print(" | ".join([short] + [str(c[o]) for o in OPS] + [str(t)]))
        tot.update(c)
    print("# total | " + " | ".join(str(tot[o]) for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
