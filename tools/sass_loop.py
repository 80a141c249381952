#!/usr/bin/env python
"""Static SASS view of one kernel: lists backward branches (loops) with their instruction counts and an
opcode histogram of a chosen address range. usage: sass_loop.py <lib.so> <substring of mangled name> [lo hi]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
f = next(x for x in funcs if pat in x.split("\n", 1)[0])
ins = []
for line in f.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f.split("\n", 1)[0], len(ins), "instructions")
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    h = collections.Counter()
    for a, s in ins:
        if lo <= a <= hi:
            op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
            h[op] += 1
    tot = sum(h.values()); print("range", hex(lo), hex(hi), tot, "instructions")
    for op, n in h.most_common(): print(f"  {op:10s} {n}")
else:
    for a, s in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", s)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16); print(f"loop {t:#x}..{a:#x}: {(a - t) // 16 + 1} instr   [{s}]")
