"""Per-kernel summary of the LAST frame of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv of tools/profile_frame.py):
launch count, total / mean device time and share, grouped by kernel name (template arguments kept for conv_gemm).
usage: launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = [(r[kn], float(r[mv].replace(",", "")) / (1e3 if r[mu] in ("ns", "nsecond") else 1)) for r in rows[hi + 1:] if len(r) > mv and r[mv]]
start = max(i for i, (n, _) in enumerate(L) if "stem_ln_kernel" in n)
L = L[start:]


def short(n):
    n = re.sub(r"^void ", "", n)
    m = re.match(r"(uc::)?(\w+)(<[^>]*>)?", n)
    base = m.group(2) if m else n
    return base + (m.group(3) if m and m.group(3) and "conv_gemm" in base else "")


agg = collections.OrderedDict()
for n, us in L:
    a = agg.setdefault(short(n), [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(us for _, us in L)
print(f"last frame: {len(L)} launches, {tot:.0f} us serialised")
fam = collections.defaultdict(float)
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={n:4d} total {us:8.1f} us  mean {us / n:7.2f} us  {100 * us / tot:5.1f}%")
    fam[k.split("<")[0]] += us
print("--- by family")
for k, us in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f"{k:32s} {us:8.1f} us {100 * us / tot:5.1f}%")
