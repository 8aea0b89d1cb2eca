#!/usr/bin/env python
"""sass_trace.py <object-or-so> <kernel-substring> [lo hi]  -- per-kernel SASS summary: max register, local-memory ops,
FP64 / LDS counts and a run-length opcode trace (D = DFMA/DMUL/DADD) so that the LDS -> first-use distance is visible."""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat not in name:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    regs = [int(x) for x in re.findall(r"\bR(\d+)\b", f)]
    ops = [o.split(".")[0] for _, o, _ in ins]
    print(name[:150])
    print("  instructions", len(ins), "maxR", max(regs), "FP64", sum(o in ("DFMA", "DMUL", "DADD") for o in ops), "LDS", ops.count("LDS"),
          "LDL", ops.count("LDL"), "STL", ops.count("STL"), "UTMALDG", ops.count("UTMALDG"))
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        out, prev, c = [], None, 0
        for a, o, _ in ins:
            if not (lo <= a < hi):
                continue
            o = o.split(".")[0]
            cat = "D" if o in ("DFMA", "DMUL", "DADD") else o
            if cat == prev:
                c += 1
            else:
                if prev:
                    out.append("%d%s" % (c, prev))
                prev, c = cat, 1
        out.append("%d%s" % (c, prev))
        print("  " + " ".join(out))
