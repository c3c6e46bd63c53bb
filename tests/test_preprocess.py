"""Letterbox preprocessing: the numpy oracle is pinned bit for bit to cv2 (the reference's third-party resize) and to the
committed fixture; the CUDA kernel (uc_letterbox_u8) is bit-exact against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
SHAPES = [((480, 640), (800, 1280)), ((1080, 1920), (800, 1280)), ((375, 1242), (800, 1280)), ((500, 333), (800, 1280)),
          ((37, 53), (320, 320)), ((800, 1280), (800, 1280)), ((300, 300), (320, 320)), ((2160, 3840), (1536, 2048))]


def test_oracle_matches_golden_fixture():
    import preprocess_oracle as po
    g = np.load(os.path.join(ROOT, "tests", "golden", "letterbox.npz"))
    i = 0
    while f"img{i}" in g:
        out, _ = po.letterbox(g[f"img{i}"], tuple(int(v) for v in g[f"size{i}"]), swap_rb=True)
        assert np.array_equal(out, g[f"out{i}"]), f"case {i}"
        i += 1
    assert i >= 5


@pytest.mark.parametrize("hw,size", SHAPES[:6])
def test_oracle_matches_cv2(hw, size):
    cv2 = pytest.importorskip("cv2")
    import preprocess_oracle as po
    rng = np.random.default_rng(hash(hw) % 1000)
    img = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
    r = min(size[0] / hw[0], size[1] / hw[1])
    ref = cv2.resize(cv2.cvtColor(img, cv2.COLOR_RGB2BGR), (int(hw[1] * r), int(hw[0] * r)), interpolation=cv2.INTER_LINEAR)
    out, r2 = po.letterbox(img, size, swap_rb=True)
    assert r2 == r
    assert np.array_equal(out[:ref.shape[0], :ref.shape[1]], ref)
    assert (out[ref.shape[0]:] == 114).all() and (out[:, ref.shape[1]:] == 114).all()


@pytest.mark.gpu
@pytest.mark.parametrize("hw,size", SHAPES)
@pytest.mark.parametrize("swap", [True, False])
def test_letterbox_kernel_bit_exact(hw, size, swap):
    import preprocess_oracle as po
    from unicorn_b200 import ops
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
    ref, r = po.letterbox(img, size, swap_rb=swap)
    out, r2 = ops.letterbox_u8(torch.from_numpy(img).cuda(), size, swap_rb=swap)
    assert r2 == r
    assert torch.equal(out.cpu(), torch.from_numpy(ref)[None])


@pytest.mark.gpu
def test_sot_device_preprocessing_equals_host_path():
    """UnicornSOTTrack with device_preproc=True (raw frame uploaded, letterbox on the GPU) returns exactly the boxes of the
    cv2 host path."""
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.sot import UnicornSOTTrack
    from unicorn_b200.weights import make_state_dict
    name = "unicorn_track_tiny"
    eng = UnicornEngine(make_state_dict(name, 0), name)
    rng = np.random.default_rng(11)
    frames = [rng.integers(0, 256, (240, 400, 3), dtype=np.uint8) for _ in range(3)]
    for f in frames:
        f[60:150, 100:220] = (np.linspace(0, 255, 120)[None, :, None] * np.ones((90, 1, 3))).astype(np.uint8)
    res = []
    for dev_pp in (False, True):
        trk = UnicornSOTTrack(eng, (320, 320), use_graph=False, device_preproc=dev_pp)
        trk.initialize(frames[0], {"init_bbox": [100, 60, 120, 90]})
        res.append([trk.track(f)["target_bbox"] for f in frames[1:]])
    assert res[0] == res[1], res
