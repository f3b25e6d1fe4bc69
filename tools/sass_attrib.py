#!/usr/bin/env python
"""Attribute executed warp instructions / stall samples of one kernel to source lines.
usage: tools/sass_attrib.py <nvdisasm -gi output> <mangled kernel name> <ncu --page source --csv file> [per-step divisor]"""
import collections
import csv
import re
import sys

sass, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
N = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":")][0]
instrs, chain, fresh = [], [], True
i = start + 1
while i < len(lines) and not lines[i].lstrip().startswith(".section"):
    l = lines[i]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2)), "/root/repo" in m.group(1)))
    elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        mine = [c for c in chain if c[2]]
        instrs.append((mine[0][:2] if mine else ("?", 0), tuple(c[:2] for c in mine)))
        fresh = True
    i += 1
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
assert len(data) == len(instrs), (len(data), len(instrs))
by_line, by_line_s = collections.Counter(), collections.Counter()
by_outer = collections.Counter()
tot = sum(int(r[iI]) for r in data)
tots = sum(int(r[iS]) for r in data)
for (loc, ch), r in zip(instrs, data):
    by_line[loc] += int(r[iI])
    by_line_s[loc] += int(r[iS])
    # outermost frame below the kernel body = second to last entry
    outer = ch[-2] if len(ch) >= 2 else (ch[-1] if ch else ("?", 0))
    by_outer[outer] += int(r[iI])
print(f"total warp instr {tot}  per unit {tot / N:.0f}   samples {tots}")
print("--- by innermost line (own files) ---")
for loc, c in by_line.most_common(45):
    print(f"{loc[0]:>18s}:{loc[1]:<4d} instr/unit {c / N:9.0f} {100 * c / tot:5.1f}%   samples {100 * by_line_s[loc] / tots:5.1f}%")
print("--- by call site in the step function ---")
for loc, c in by_outer.most_common(25):
    print(f"{loc[0]:>18s}:{loc[1]:<4d} instr/unit {c / N:9.0f} {100 * c / tot:5.1f}%")
if len(sys.argv) > 6:
    fname, lo = sys.argv[5], int(sys.argv[6])
    agg = collections.Counter()
    for (loc, ch), r in zip(instrs, data):
        inside = [c for c in ch if c[0] == fname and c[1] >= lo]
        if inside:
            agg[inside[-1]] += int(r[iI])
    print(f"--- by outermost line inside {fname} >= {lo} ---")
    for loc, c in sorted(agg.items(), key=lambda kv: kv[0][1]):
        if c / tot > 0.002:
            print(f"{loc[0]:>18s}:{loc[1]:<4d} instr/unit {c / N:9.0f} {100 * c / tot:5.1f}%")
