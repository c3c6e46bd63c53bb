"""The synthetic VOS sequence shared by tests/golden/make_golden_vos.py (reference run) and the tests that replay it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
from unicorn_b200.synthetic import make_video  # noqa: E402

H0, W0, NF, NEW_AT = 256, 400, 5, 2


def make_sequence():
    """RGB uint8 frames [NF,H0,W0,3], boxes xywh [NF,3,4], label map (uint8 [H0,W0]) of the frame where object "3" appears."""
    frames, boxes = make_video(NF, H0, W0, seed=7, n_obj=3)
    rgb = frames.permute(0, 2, 3, 1).round().clamp(0, 255).to(torch.uint8).numpy()
    xywh = boxes.clone()
    xywh[..., 2:] -= xywh[..., :2]
    lab = np.zeros((H0, W0), dtype=np.uint8)
    x1, y1, x2, y2 = boxes[NEW_AT, 2].int().tolist()
    lab[y1:y2, x1:x2] = 3
    return rgb, xywh, lab


def prep_frame(img, size):
    """PreprocessorX.process through the bit-exact cv2 restatement (oracle/preprocess_oracle.py) -> fp32 [1,3,H,W]."""
    from preprocess_oracle import letterbox
    return torch.from_numpy(letterbox(img, size, swap_rb=True)[0]).permute(2, 0, 1)[None].float()


def box_xyxy(xywh, r):
    b = torch.as_tensor(xywh, dtype=torch.float32).clone()
    b[2:] += b[:2]
    return b * r
