#!/usr/bin/env python
"""Print one address range of a kernel's SASS (cuobjdump), one instruction per line, with a pipe-class histogram.
    python tools/sass_loop.py LIB KERNEL_SUBSTR START_HEX END_HEX [-q]
"""
import re, subprocess, sys, collections
lib, name, a0, a1 = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
quiet = len(sys.argv) > 5
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if name not in f.split("\n", 1)[0]:
        continue
    hist = collections.Counter(); n = 0
    for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f):
        ad = int(m.group(1), 16)
        if a0 <= ad <= a1:
            ins = m.group(2).strip()
            op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
            hist[op.split(".")[0]] += 1; n += 1
            if not quiet: print(f"{ad:05x}  {ins}")
    print(n, "instr", dict(hist.most_common()))
    break
