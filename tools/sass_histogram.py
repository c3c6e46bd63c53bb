"""SASS opcode histogram of every object of the built library (cuobjdump -sass): which kernels carry tcgen05 / TMA / TMEM / packed-FMA
instructions.  UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA tensor loads / stores, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, FFMA2 = fma.rn.f32x2, LDGSTS = cp.async, REDG / ATOMG = global reductions / atomics.
usage: sass_histogram.py [build_dir] > profiles/rN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

bdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unicorn_b200", "build")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "HMMA", "MUFU",
         "LDGSTS", "LDG", "STG", "LDS", "STS", "REDG", "ATOMG", "ATOMS", "RED", "SHFL", "BAR", "ELECT", "ACQBULK", "FFMA", "IMAD", "LOP3"]
for obj in sorted(f for f in os.listdir(bdir) if f.endswith(".o")):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(bdir, obj)], capture_output=True, text=True).stdout
    kern, per = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and kern:
            per[kern][m.group(1)] += 1
    print(f"== {obj}")
    for k, c in per.items():
        tot = sum(c.values())
        watched = ", ".join(f"{w} {c[w]}" for w in WATCH if c[w])
        print(f"  {k[:110]:110s} {tot:6d} instr | {watched}")
