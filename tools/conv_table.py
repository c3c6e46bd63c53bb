"""Join an ncu launch list (conv_gemm launches only, issue order) with the engine's conv trace of the LAST frame:
per-layer device time, TFLOP/s, tile count.  usage: conv_table.py launches.csv trace.json"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = [(r[kn].split("<")[1].split(">")[0], float(r[mv].replace(",", "")) / (1e3 if r[mu] == "ns" else 1))
     for r in rows[hi + 1:] if len(r) > mv and "conv_gemm" in r[kn]]
tr = json.load(open(sys.argv[2]))
L = L[-len(tr):]
for (tpl, _), e in zip(L, tr):  # the kernel template tells which N tile the C side resolved block_n=0 to; check the rest
    bn, _, cl = [int(v) for v in tpl.split(",")]
    assert not e["bn"] or (e["bn"] % 1000 == bn and (e["bn"] >= 1000) == (cl == 2)), (tpl, e)
    e["bn"] = bn + (1000 if cl == 2 else 0)
t = [us for _, us in L]
agg = collections.OrderedDict()
for e, us in zip(tr, t):
    key = (e["M"], e["N"], e["K"], e["k"], e["s"], e["bn"], e["act"], e["gn"])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(t); print(f"{len(t)} conv launches, {tot:.0f} us")
print(f"{'M':>6} {'N':>5} {'K':>5} k s {'bn':>4} act gn   n   us/launch  TFLOP/s  tiles  share")
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, k, s, bn, act, gn = key
    b = bn % 1000
    tiles = -(-M // 128) * -(-N // b)
    print(f"{M:6d} {N:5d} {K:5d} {k} {s} {bn:4d} {act:3d} {gn:2d} {n:3d} {us/n:10.1f} {2.0*M*N*K*n/us/1e6:8.1f} {tiles:6d} {100*us/tot:5.1f}%")
