"""profiles/r2_roofline_inputs.json: the per-launch DRAM traffic numbers bench.py quotes in its `roofline*` objects, read from the committed
ncu summaries (tools/ncu_summary.py output) and, for the whole frame, from an ncu launch list with dram__bytes metrics.
usage: make_roofline_inputs.py summary1.csv [summary2.csv ...] [--frame frame_dram.csv]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
args = sys.argv[1:]
frame_csv = None
if "--frame" in args:
    i = args.index("--frame")
    frame_csv = args[i + 1]
    args = args[:i] + args[i + 2:]
out = {"source": "ncu --set full --clock-control none, one launch each, L2 not flushed by ncu (--cache-control none is NOT set: cold caches)", "kernels": {}}
for path in args:
    per = {}
    for rep, metric, unit, value in csv.reader(open(path)):
        if rep == "report":
            continue
        per.setdefault(rep, {})[metric] = (unit, value)
    for rep, m in per.items():
        rd = float(m["dram__bytes_read.sum"][1]) * UNIT[m["dram__bytes_read.sum"][0]]
        wr = float(m["dram__bytes_write.sum"][1]) * UNIT[m["dram__bytes_write.sum"][0]]
        out["kernels"][rep.replace(".ncu-rep", "")] = {"kernel": m["Kernel Name"][1], "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                                                        "us": float(m["gpu__time_duration.sum"][1]), "summary": os.path.relpath(path, ROOT)}
if frame_csv:
    rows = list(csv.reader(open(frame_csv)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    idc = h.index("ID")
    L = [(int(r[idc]), r[kn], r[mn], float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)) for r in rows[hi + 1:] if len(r) > mv and r[mv]]
    start = max(i for i, n, _, _ in L if "stem_ln_kernel" in n)
    tot = sum(v for i, n, m, v in L if i >= start and m.startswith("dram__bytes"))
    out["frame"] = {"dram_bytes": tot, "launch_list": os.path.relpath(frame_csv, ROOT),
                    "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over every kernel of the last (steady-state) 800x1280 SOT frame"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_roofline_inputs.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
