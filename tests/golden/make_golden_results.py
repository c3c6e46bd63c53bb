"""Golden text of the reference's MOT / MOTS result writers (unicorn/evaluators/mot_evaluator.py:37-72), produced by executing
exactly those three function definitions of the UNMODIFIED reference file (the module itself does not import here: mmcv,
pycocotools, motmetrics are absent).  Run in the build container:  python tests/golden/make_golden_results.py"""
import ast
import json
import os
import tempfile

import numpy as np

SRC = "/root/reference/unicorn/evaluators/mot_evaluator.py"
tree = ast.parse(open(SRC).read())
want = {"write_results", "write_results_no_score", "write_results_mots"}
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
ns = {"logger": type("L", (), {"info": staticmethod(lambda *a, **k: None)})}
exec(compile(mod, SRC, "exec"), ns)

rng = np.random.default_rng(5)
res, res_ns, res_mots = [], [], []
for frame in range(1, 6):
    n = int(rng.integers(2, 6))
    tlwhs = (rng.random((n, 4)) * np.array([1200, 700, 300, 400])).tolist()
    ids = [int(v) for v in rng.integers(-1, 40, n)]
    scores = rng.random(n).tolist()
    res.append((frame, tlwhs, ids, scores))
    res_ns.append((frame, tlwhs, ids))
    res_mots.append((frame, ids, 2, 720, 1280, [f"rle{frame}_{k}" for k in range(n)]))
out = {"results": res, "results_no_score": res_ns, "results_mots": res_mots}
with tempfile.TemporaryDirectory() as d:
    for key, fn, arg in (("txt", "write_results", res), ("txt_no_score", "write_results_no_score", res_ns), ("txt_mots", "write_results_mots", res_mots)):
        p = os.path.join(d, key)
        ns[fn](p, arg)
        out[key] = open(p).read()
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "results_txt.json"), "w"), indent=0)
print("wrote results_txt.json:", {k: len(v) for k, v in out.items()})
