"""LayerNorm folded into pwconv1 (engine ln_fold=True; round-2 item, written without GPU access): kernel-level check of the
statistics + folded epilogue against fp32, and the tiny SOT frame against the oracle with the tolerances of test_engine_gpu."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
dev = "cuda"


@pytest.mark.parametrize("C,H,W", [(192, 40, 56), (768, 25, 40), (96, 20, 28)])
def test_dwconv_stats_and_folded_pwconv1(C, H, W):
    from unicorn_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, H, W, C, generator=g).to(dev).bfloat16()
    wd = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    bd, lw, lb = (torch.randn(C, generator=g).to(dev) for _ in range(3))
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dev)
    b1 = torch.randn(4 * C, generator=g).to(dev)
    st = torch.zeros(H * W, 2, dtype=torch.int64, device=dev)
    t = ops.dwconv7(x, ops.pack_dw_weight(wd), bd, ln_stats=st)
    tf = t.float().reshape(-1, C)
    assert torch.allclose(st[:, 0].double() / 4194304.0, tf.double().sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[:, 1].double() / 4194304.0, (tf.double() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    w1f = ops.pack_conv_weight((w1 * lw[None, :])[:, :, None, None])
    s1 = w1f.float().sum(dim=(1, 2)).contiguous()
    c1 = (w1 @ lb + b1).contiguous()
    got = ops.conv2d(t, w1f, 1, 1, bias=c1, act=ops.ACT_GELU, row_stats=st, col_s=s1, row_eps=1e-6)
    ref = F.gelu(F.linear(F.layer_norm(tf, (C,), lw, lb, 1e-6), w1, b1))
    err = (got.float().reshape(-1, 4 * C) - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err


def test_sot_frame_with_ln_fold_vs_oracle():
    import unicorn_oracle as orc
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.sot import UnicornSOTTrack
    from unicorn_b200.synthetic import make_video
    from unicorn_b200.weights import make_state_dict
    name = "unicorn_track_tiny"
    sd = make_state_dict(name, 0)
    frames, boxes = make_video(3, 320, 320, seed=0)
    o = orc.SOTOracle(sd, name)
    o.initialize(frames[0:1], boxes[0, 0])
    st = {}
    o.track(frames[2:3], st)
    res = []
    for graph in (False, True):
        eng = UnicornEngine(sd, name, ln_fold=True)
        trk = UnicornSOTTrack(eng, (320, 320), use_graph=graph, full_nms=True)
        trk.initialize_tensor(frames[0:1], boxes[0, 0])
        if graph:
            trk.track_tensor(frames[1:2])
        dets, n = trk.track_tensor(frames[2:3])
        res.append((dets.clone(), n))
        last = trk.last
        rel = lambda a, b: ((a.float().cpu() - b).abs().max() / (b.abs().max() + 1e-12)).item()  # noqa: E731
        nchw = lambda t: t.float().permute(0, 3, 1, 2)  # noqa: E731
        assert rel(nchw(last["feat"]), st["feat"]) < 4e-2
        assert rel(nchw(last["fpn"][2]), st["fpn"][2]) < 8e-2
        assert (last["priors"][0].cpu() - st["coarse"][0]).abs().max().item() < 6e-2
        assert (last["head"].cpu()[..., 4:] - st["head"][..., 4:]).abs().max().item() < 5e-2
    assert res[0][1] == res[1][1] and torch.equal(res[0][0], res[1][0])  # eager == graph, bit for bit
