#!/usr/bin/env python
"""Opcode histogram of every kernel in libbfx.so (cuobjdump -sass), written to profiles/sass_opcodes.txt.
The Blackwell-specific mnemonics are listed per kernel: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), UTMALDG / UTMASTG (TMA bulk tensor load / store), UTMAPF (tensormap prefetch),
LDGSTS (cp.async), SYNCS (mbarrier), UCGABAR (cluster barrier).
usage: python scripts/sass_opcodes.py [lib] [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bayesflux.jl_b200", "libbfx.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_opcodes.txt")
KEY = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDGSTS", "SYNCS", "UCGABAR", "UTCATOMSWS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kern, hist = None, collections.OrderedDict()
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
tot = collections.Counter()
with open(out, "w") as f:
    f.write(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: opcode histogram per kernel (static instruction counts)\n")
    f.write("# kernel | instructions | " + " ".join(KEY) + " | ten most frequent opcodes\n")
    for k, h in hist.items():
        n = sum(h.values())
        tot.update(h)
        name = demangle(k)
        name = re.sub(r"\((int|bool|unsigned)\)", "", name)
        name = re.sub(r"\(.*$", "", name)[:130]
        f.write(f"{name} | {n} | " + " ".join(f"{key}={h.get(key, 0)}" for key in KEY if h.get(key, 0)) + " | " +
                " ".join(f"{o}:{c}" for o, c in h.most_common(10)) + "\n")
    f.write("# whole library: " + " ".join(f"{key}={tot.get(key, 0)}" for key in KEY) + f" kernels={len(hist)}\n")
print(open(out).read().split("\n")[-2])
