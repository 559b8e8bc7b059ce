#!/usr/bin/env python
"""Static loop census of a kernel's SASS (cuobjdump -sass -fun <kernel> file): every backward branch closes a
loop; prints each loop's instruction count and opcode histogram.  Used to track the instruction budget of the
sample loops of the column kernels without a GPU."""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, txt) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?(0x[0-9a-f]+)", txt)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
min_size = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (s, e) in loops:
    n = e - s + 1
    if n < min_size:
        continue
    inner = [(s2, e2) for (s2, e2) in loops if s2 >= s and e2 <= e and (s2, e2) != (s, e)]
    hist = collections.Counter()
    for _, txt in ins[s:e + 1]:
        op = txt.split()[0]
        if op.startswith("@"):
            op = txt.split()[1]
        hist[op.split(".")[0]] += 1
    print(f"loop 0x{ins[s][0]:x}..0x{ins[e][0]:x}: {n} instr, {len(inner)} inner loops")
    print("   " + ", ".join(f"{k}:{v}" for k, v in hist.most_common(24)))
