#!/usr/bin/env python
"""Static view of a kernel's SASS (no GPU needed): find the loops (backward branches), print each
loop's instruction count and opcode mix, so per-iteration instruction budgets can be compared
between builds.   cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_loops.py k.sass"""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)", t) or re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m and t.split()[0].lstrip("@!P0123456789U ").startswith("BRA") or (m and "BRA" in t):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
loops.sort(key=lambda p: p[1] - p[0])
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for s, e in loops:
    n = e - s + 1
    if n < minlen:
        continue
    mix = collections.Counter()
    for _, t in ins[s:e + 1]:
        op = t.split()
        op = op[1] if op[0].startswith("@") else op[0]
        mix[op.split(".")[0]] += 1
    inner = [(s2, e2) for s2, e2 in loops if s2 >= s and e2 <= e and (s2, e2) != (s, e)]
    print(f"loop {ins[s][0]:#06x}..{ins[e][0]:#06x}: {n} instr, inner loops {len(inner)}: " +
          ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
