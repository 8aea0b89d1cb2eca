#!/usr/bin/env python
"""sass_excerpt.py [object] > profiles/r2_sass_tmarch_ws.txt -- cuobjdump -sass excerpt of the fused MD-step kernel: static instruction mix,
the TMA / mbarrier / setmaxnreg evidence, a producer-warp window around its tensor copies and a link-warp window of the staple arithmetic."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gaugefields.jl_b200", "build", "default", "tmarch.o")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
body = None
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if "k_tmarch_wsILb1ELb1ELb1" in f.split("\n")[0]:
        body = f
assert body, "kernel not found"
lines = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines]
ops = [re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l).group(1) for l in lines]
mix = collections.Counter(ops)
print("cuobjdump -sass excerpt of k_tmarch_ws<READ_Z=1,WRITE_Z=1,DO_EXP=1> (gaugefields.jl_b200/csrc/tmarch.cu, the fused MD-step kernel that bench.py times;")
print("final kernel of round 2: round barrier, four producer warps, mbarrier waits with a suspend-time hint)")
print("built by gaugefields.jl_b200/build.py: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3; regenerate: python scripts/sass_excerpt.py\n")
print("static instruction mix (%d instructions): %s" % (len(lines), ", ".join("%s %d" % kv for kv in mix.most_common(24))))
lds128 = sum("LDS.128" in l for l in lines)
print("TMA / mbarrier / register-reallocation evidence: UTMALDG %d, SYNCS %d, USETMAXREG %d, ELECT %d, LDS.128 %d, STG %d, NANOSLEEP %d\n"
      % (mix["UTMALDG"], mix["SYNCS"], mix["USETMAXREG"], mix["ELECT"], lds128, mix["STG"], mix["NANOSLEEP"]))


def window(title, center, before, after):
    print("---- " + title)
    for l in lines[max(0, center - before):center + after]:
        print(l.rstrip())
    print("        ...\n")


first_tma = next(i for i, l in enumerate(lines) if "UTMALDG" in l)
window("producer warps: tensor copies of one part, completion counted in bytes on an mbarrier", first_tma, 14, 10)
for i, l in enumerate(lines):
    if "USETMAXREG" in l:
        window("register reallocation between the producer warpgroup and the 8 link warps (setmaxnreg)", i, 1, 3)
tw = next(i for i, l in enumerate(lines) if "TRYWAIT" in l)
window("mbarrier wait (try_wait with a suspend-time hint)", tw, 3, 4)
# densest DFMA window that also holds LDS.128
best, besti = -1, 0
for i in range(0, len(lines) - 80, 8):
    w = lines[i:i + 80]
    d = sum("DFMA" in l for l in w)
    if sum("LDS.128" in l for l in w) >= 4 and d > best:
        best, besti = d, i
window("link warps: a window of the staple arithmetic (operands are LDS.128 from the TMA-filled ring, two-row SU(3) products in DFMA)", besti, 0, 80)
