"""Golden vectors for the letterbox preprocessing: outputs of the reference's own recipe (cv2.cvtColor + cv2.resize INTER_LINEAR
+ pad 114, external/lib/test/tracker/unicorn_sot.py:114-123) computed with the cv2 of the build container on small seeded
images.  Run in the build container:  python tests/golden/make_letterbox_golden.py"""
import os

import cv2
import numpy as np

CASES = [((24, 32), (40, 64)), ((45, 35), (32, 32)), ((19, 27), (48, 80)), ((32, 32), (32, 32)), ((17, 50), (40, 64))]
out = {"cv2_version": np.array(cv2.__version__)}
rng = np.random.default_rng(7)
for i, ((h, w), size) in enumerate(CASES):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    r = min(size[0] / h, size[1] / w)
    rs = cv2.resize(cv2.cvtColor(img, cv2.COLOR_RGB2BGR), (int(w * r), int(h * r)), interpolation=cv2.INTER_LINEAR)
    pad = np.full((size[0], size[1], 3), 114, np.uint8)
    pad[:int(h * r), :int(w * r)] = rs
    out[f"img{i}"], out[f"out{i}"], out[f"size{i}"] = img, pad, np.array(size)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "letterbox.npz"), **out)
print("wrote letterbox.npz with", len(CASES), "cases, cv2", cv2.__version__)
