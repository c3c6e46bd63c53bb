"""Unit parity of every non-GEMM kernel against plain torch fp32 on the same (rounded) operands."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def G(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def close(got, ref, rel, name=""):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    assert err <= rel * scale + 1e-6, f"{name}: max err {err:.4g} vs scale {scale:.4g} (rel tol {rel})"


@pytest.mark.parametrize("C0,H,W", [(96, 64, 96), (192, 32, 52)])
def test_stem(C0, H, W):
    from unicorn_b200 import ops
    g = G(1)
    img = (torch.rand(2, 3, H, W, generator=g) * 255).to(dev)
    w = (torch.randn(C0, 3, 4, 4, generator=g) / 7).to(dev)
    b, lw, lb = (torch.randn(C0, generator=g).to(dev) for _ in range(3))
    out = ops.stem_ln(img, ops.pack_stem_weight(w), b, lw, lb)
    x = F.conv2d(img, w, b, stride=4).permute(0, 2, 3, 1)
    ref = F.layer_norm(x, (C0,), lw, lb, 1e-6)
    close(out, ref, 6e-3, "stem")


@pytest.mark.parametrize("C,H,W,B", [(96, 20, 28, 2), (192, 17, 23, 2), (256, 10, 16, 2), (1536, 5, 9, 2), (384, 8, 8, 2),
                                     # the four tile shapes of the fused kernel: 16x8, 8x8, 8x4, 4x2 pixels per CTA
                                     (192, 64, 120, 2), (384, 100, 160, 1), (768, 60, 100, 2), (768, 50, 80, 1), (1536, 25, 40, 1)])
def test_dwconv_ln(C, H, W, B):
    from unicorn_b200 import ops
    g = G(2)
    x = torch.randn(B, H, W, C, generator=g).to(dev).bfloat16()
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    b, lw, lb = (torch.randn(C, generator=g).to(dev) for _ in range(3))
    out = ops.dwconv7_ln(x, ops.pack_dw_weight(w), b, lw, lb)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=3, groups=C).permute(0, 2, 3, 1)
    ref = F.layer_norm(y, (C,), lw, lb, 1e-6)
    close(out, ref, 6e-3, "dwconv_ln")


@pytest.mark.parametrize("C,H,W", [(96, 20, 28), (192, 17, 23), (256, 10, 16), (1536, 5, 9), (384, 33, 40),
                                   # ConvNeXt-L layer shapes at 800x1280 (TMA kernel: 16x4-pixel x 64-channel items, edge tiles, 24 chunks)
                                   (768, 50, 80), (192, 200, 320), (1536, 25, 40), (256, 100, 160), (104, 9, 3)])
def test_dwconv_tiled(C, H, W):
    from unicorn_b200 import ops
    g = G(21)
    x = torch.randn(2, H, W, C, generator=g).to(dev).bfloat16()
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    out = ops.dwconv7(x, ops.pack_dw_weight(w), b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=3, groups=C).permute(0, 2, 3, 1)
    close(out, ref, 5e-3, "dwconv tiled")
    # tiles handed out by an atomic work counter (the engine's mode) instead of the static round-robin: same bits; and with the
    # LayerNorm statistics: sum / sum of squares over C of the stored values, fixed point 2^22
    st = torch.zeros(2 * H * W, 2, dtype=torch.int64, device=dev)
    out2 = ops.dwconv7(x, ops.pack_dw_weight(w), b, ln_stats=st, work_counter=torch.zeros(1, dtype=torch.int32, device=dev))
    assert torch.equal(out, out2)
    of = out.float().reshape(-1, C).double()
    assert torch.allclose(st[:, 0].double() / 4194304.0, of.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[:, 1].double() / 4194304.0, (of ** 2).sum(1), rtol=1e-5, atol=1e-3)
    # the tensor-core kernel (Toeplitz blocks on mma.sync; taps in bf16): against the same convolution with the taps it uses, and
    # within bf16-tap rounding of the fp32-tap one; static and counter schedules give the same bits
    wf = ops.pack_dw_weight_mma(w, b)
    om = ops.dwconv7_mma(x, wf)
    refb = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=3, groups=C).permute(0, 2, 3, 1)
    close(om, refb, 5e-3, "dwconv mma (bf16 taps)")
    close(om, ref, 8e-3, "dwconv mma vs fp32 taps")
    om2 = ops.dwconv7_mma(x, wf, work_counter=torch.zeros(1, dtype=torch.int32, device=dev))
    assert torch.equal(om, om2)


@pytest.mark.parametrize("C,M", [(192, 128), (192, 1000), (96, 777), (192, 64000), (96, 40000), (384, 100), (384, 16000), (384, 49152), (256, 300), (256, 16000)])
def test_convnext_mlp_fused(C, M):
    """uc_convnext_mlp (LayerNorm -> pwconv1 -> GELU -> pwconv2 -> layer scale -> residual in one launch, convnext.py:45-52) against
    the same chain in fp32 torch on the bf16 inputs / weights, and against the unfused kernels of the library."""
    from unicorn_b200 import ops
    g = G(5)
    t = (torch.randn(M, C, generator=g) * 1.5 + 0.3).to(dev).bfloat16()
    x = torch.randn(M, C, generator=g).to(dev).bfloat16()
    lw, lb = (1 + 0.2 * torch.randn(C, generator=g)).to(dev), (0.1 * torch.randn(C, generator=g)).to(dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dev)
    b1 = (0.1 * torch.randn(4 * C, generator=g)).to(dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dev)
    b2 = (0.1 * torch.randn(C, generator=g)).to(dev)
    gamma = (0.5 * torch.randn(C, generator=g)).to(dev)
    w1f = (w1 * lw[None, :]).bfloat16().contiguous()
    c1 = (w1 @ lb + b1).contiguous()
    w2b = w2.bfloat16().contiguous()
    out = ops.convnext_mlp(t, w1f, c1, w2b, b2, gamma, x.clone())
    tn = F.layer_norm(t.float(), (C,), None, None, 1e-6).bfloat16().float()
    hid = F.gelu(tn @ w1f.float().t() + c1).bfloat16().float()
    ref = x.float() + gamma * (hid @ w2b.float().t() + b2)
    close(out, ref, 6e-3, "fused convnext mlp")
    # rows past M in the last tile are not written; the call is deterministic
    assert torch.equal(out, ops.convnext_mlp(t, w1f, c1, w2b, b2, gamma, x.clone()))
    # the unfused library path (uc_layernorm with affine, two uc_conv2d): same result within bf16 rounding of the intermediate maps
    tl = ops.layernorm(t, lw, lb, 1e-6)
    ref2 = x.float() + gamma * (F.gelu(tl.float() @ w1.bfloat16().float().t() + b1).bfloat16().float() @ w2b.float().t() + b2)
    close(out, ref2, 2e-2, "fused vs unfused formulation")


@pytest.mark.parametrize("C", [96, 100, 192, 256, 384, 768, 1536, 2048])  # 100: not a multiple of 8 -> the 32-bit-access kernel
def test_layernorm(C):
    from unicorn_b200 import ops
    g = G(3)
    x = torch.randn(333, C, generator=g).to(dev).bfloat16()
    r = torch.randn(333, C, generator=g).to(dev).bfloat16()
    lw, lb = (torch.randn(C, generator=g).to(dev) for _ in range(2))
    close(ops.layernorm(x, lw, lb, 1e-6), F.layer_norm(x.float(), (C,), lw, lb, 1e-6), 6e-3, "ln")
    close(ops.layernorm(x, lw, lb, 1e-5, res=r), F.layer_norm(x.float() + r.float(), (C,), lw, lb, 1e-5), 6e-3, "ln+res")
    big = torch.zeros(333, C + 64, device=dev, dtype=torch.bfloat16)
    ops.layernorm(x, lw, lb, 1e-6, out=big[:, 32:32 + C])
    close(big[:, 32:32 + C], F.layer_norm(x.float(), (C,), lw, lb, 1e-6), 6e-3, "ln slice")
    assert (big[:, :32] == 0).all() and (big[:, 32 + C:] == 0).all()


def test_conv_gn_silu_prior():
    from unicorn_b200 import ops
    g = G(4)
    B, H, W, Cin, C = 1, 20, 28, 384, 256
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(C, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    gw, gb, beta = (torch.randn(C, generator=g).to(dev) for _ in range(3))
    prior = torch.rand(B, H, W, generator=g).to(dev)
    stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.int64)
    y = ops.conv2d(x, ops.pack_conv_weight(w), 1, 1, gn_stats=stats, gn_groups=16)
    pos = torch.randn(B, H, W, C, generator=g).to(dev).bfloat16()
    q = torch.empty_like(y)
    out = ops.groupnorm_apply(y.clone(), stats, gw, gb, 16, 1e-3, ops.ACT_SILU, prior=prior, beta=beta, add2=pos, out2=q)
    pre = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float())
    ref = F.silu(F.group_norm(pre, 16, gw, gb, 1e-3)) + prior[:, None] * beta.view(1, -1, 1, 1)
    close(out, ref.permute(0, 2, 3, 1), 1e-2, "gn silu prior")
    close(q, ref.permute(0, 2, 3, 1) + pos.float(), 1.2e-2, "gn second output")


def test_data_movement():
    from unicorn_b200 import ops
    g = G(5)
    x = torch.randn(1, 6, 10, 64, generator=g).to(dev).bfloat16()
    big = torch.zeros(1, 12, 20, 160, device=dev, dtype=torch.bfloat16)
    ops.copy_upsample(x, big[..., 32:96], 2)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(big[..., 32:96].float(), ref) and (big[..., :32] == 0).all() and (big[..., 96:] == 0).all()
    y = torch.randn(1, 5, 7, 256, generator=g).to(dev).bfloat16()
    ps = ops.pixel_shuffle2(y)
    assert torch.equal(ps.float(), F.pixel_shuffle(y.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
    m = torch.rand(1, 2, 100, 160, generator=g).to(dev)
    for f in (2, 4):
        close(ops.bilinear(m, 100 // f, 160 // f, float(f), float(f)), F.interpolate(m, scale_factor=1 / f, mode="bilinear", align_corners=False), 1e-6, "bilinear down")
    tab = torch.rand(1, 256, 40, 40, generator=g).to(dev)
    close(ops.bilinear(tab, 50, 80), F.interpolate(tab, (50, 80), mode="bilinear", align_corners=False), 1e-6, "bilinear up")
    lab = torch.zeros(1, 1, 320, 320, device=dev)
    lab[..., 33:201, 57:170] = 1
    close(ops.bilinear(lab, 40, 40, 8.0, 8.0), F.interpolate(lab, scale_factor=1 / 8, mode="bilinear", align_corners=False), 1e-6, "label/8")
    a = torch.randn(100, 256, generator=g).to(dev).bfloat16()
    b = torch.randn(100, 256, generator=g).to(dev).bfloat16()
    assert torch.equal(ops.add(a, b), (a.float() + b.float()).bfloat16())
    img = torch.randn(2, 24, 9, 13, generator=g).to(dev)
    nh = ops.nchw_to_nhwc(img)
    assert torch.equal(nh, img.permute(0, 2, 3, 1).bfloat16())
    assert torch.equal(ops.nhwc_to_nchw(nh), nh.float().permute(0, 3, 1, 2))


def _msda_ref(value, shapes, loc, attn):
    """ms_deform_attn_core_pytorch semantics (grid_sample bilinear/zeros/align_corners False)."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, _, L_, P_, _ = loc.shape
    vals = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * loc - 1
    outs = []
    for lid, (h, w) in enumerate(shapes):
        v = vals[lid].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, h, w)
        gr = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        outs.append(F.grid_sample(v, gr, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    return (torch.stack(outs, dim=-2).flatten(-2) * a).sum(-1).view(N_, M_ * D_, Lq_).transpose(1, 2).contiguous()


def test_msda_reference_known_answer():
    """Shapes and seed of the reference's only known-answer test, unicorn/models/ops/test.py:21-56."""
    from unicorn_b200 import ops
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = sum(h * w for h, w in shapes.tolist())
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D, device=dev) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2, device=dev)
    attn = torch.rand(N, Lq, M, L, P, device=dev) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out = ops.msda_forward(value, shapes, lsi, loc, attn)
    ref = _msda_ref(value, shapes.tolist(), loc, attn)
    assert torch.allclose(out, ref, rtol=1e-2, atol=1e-3)  # the reference's own fp32 tolerance
    assert (out - ref).abs().max().item() < 1e-6


def test_msda_random_and_fused():
    from unicorn_b200 import ops
    g = G(7)
    M, D, L, P = 8, 32, 2, 4
    hw = [(13, 21), (13, 21)]
    S = sum(h * w for h, w in hw)
    value = torch.randn(1, S, M, D, generator=g).to(dev)
    loc = (torch.rand(1, S, M, L, P, 2, generator=g) * 1.3 - 0.15).to(dev)  # some samples fall outside
    attn = torch.softmax(torch.randn(1, S, M, L * P, generator=g), -1).view(1, S, M, L, P).to(dev)
    shapes = torch.as_tensor(hw, dtype=torch.long, device=dev)
    lsi = torch.as_tensor([0, hw[0][0] * hw[0][1]], dtype=torch.long, device=dev)
    out = ops.msda_forward(value, shapes, lsi, loc, attn)
    close(out, _msda_ref(value, hw, loc, attn), 1e-5, "msda f32")
    # fused form: raw offsets + logits, reference points from the pixel grid
    off = (torch.randn(S, M, L, P, 2, generator=g) * 3).to(dev)
    logit = torch.randn(S, M, L * P, generator=g).to(dev)
    offlog = torch.cat([off.reshape(S, -1), logit.reshape(S, -1)], 1).contiguous()
    vb = value[0].reshape(S, M * D).bfloat16().contiguous()
    fused = ops.msda_fused(vb, offlog, hw, M, P)
    refs = []
    for (h, w) in hw:
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    ref_pts = torch.cat(refs, 0).to(dev)  # (S,2)
    norm = torch.tensor([[w, h] for (h, w) in hw], dtype=torch.float32, device=dev)
    loc2 = ref_pts[None, :, None, None, None, :] + off[None] / norm[None, None, None, :, None, :]
    attn2 = torch.softmax(logit, -1).view(1, S, M, L, P)
    ref = _msda_ref(vb.float().view(1, S, M, D), hw, loc2, attn2)
    close(fused, ref[0], 6e-3, "msda fused")


@pytest.mark.parametrize("n_ref,n_cur,n_obj,dt", [(1600, 1600, 1, torch.float16), (1000, 777, 3, torch.bfloat16), (4000, 4000, 8, torch.float16)])
def test_corr(n_ref, n_cur, n_obj, dt):
    from unicorn_b200 import ops
    g = G(8)
    k = (torch.randn(n_ref, 128, generator=g) * 0.6).to(dev).to(dt)
    q = (torch.randn(n_cur, 128, generator=g) * 0.6).to(dev).to(dt)
    v = torch.rand(n_obj, n_ref, generator=g).to(dev)
    out = ops.corr_propagate(k, q, v)
    S = k.float() @ q.float().t()
    ref = v @ torch.softmax(S, dim=0)
    close(out, ref, 2e-4, "corr")


def test_postprocess_matches_torchvision():
    import torchvision
    from unicorn_b200 import ops
    g = G(9)
    for A, ncls in ((2100, 1), (5000, 8), (21000, 1)):
        cx = torch.rand(A, generator=g) * 1280
        cy = torch.rand(A, generator=g) * 800
        w = torch.rand(A, generator=g) * 200 + 20
        h = torch.rand(A, generator=g) * 200 + 20
        obj = torch.rand(A, generator=g)
        cls = torch.rand(A, ncls, generator=g)
        pred = torch.cat([cx[:, None], cy[:, None], w[:, None], h[:, None], obj[:, None], cls], 1).to(dev).contiguous()
        ws = ops.PostWorkspace(A, dev)
        dets, cnt = ops.postprocess_device(pred, ncls, 0.3, 0.65, ws)
        n = int(cnt.item())
        got = dets[:n]
        box = torch.stack([pred[:, 0] - pred[:, 2] / 2, pred[:, 1] - pred[:, 3] / 2, pred[:, 0] + pred[:, 2] / 2, pred[:, 1] + pred[:, 3] / 2], 1)
        cc, cp = pred[:, 5:].max(1)
        sc = pred[:, 4] * cc
        m = sc >= 0.3
        keep = torchvision.ops.batched_nms(box[m], sc[m], cp[m], 0.65)
        ref = torch.cat([box[m], pred[m, 4:5], cc[m, None], cp[m, None].float()], 1)[keep]
        assert n == ref.shape[0], (A, n, ref.shape[0])
        assert torch.allclose(got, ref, rtol=0, atol=1e-4), (got - ref).abs().max()


def test_head_decode():
    from unicorn_b200 import ops
    g = G(10)
    hw = [(8, 12), (4, 6), (2, 3)]
    ncls = 3
    ro = [torch.randn(h * w, 8, generator=g).to(dev) for h, w in hw]
    cl = [torch.randn(h * w, 8, generator=g).to(dev) for h, w in hw]
    out = ops.head_decode(ro, cl, hw, (8, 16, 32), ncls)
    rows = []
    for (h, w), s, r, c in zip(hw, (8, 16, 32), ro, cl):
        yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grid = torch.stack((xv, yv), 2).view(-1, 2).float().to(dev)
        rows.append(torch.cat([(r[:, :2] + grid) * s, torch.exp(r[:, 2:4]) * s, torch.sigmoid(r[:, 4:5]), torch.sigmoid(c[:, :ncls])], 1))
    close(out[0], torch.cat(rows, 0), 1e-6, "decode")


def test_postprocess_max_keep_is_a_prefix():
    from unicorn_b200 import ops
    g = G(11)
    A = 6000
    pred = torch.cat([torch.rand(A, 1, generator=g) * 1280, torch.rand(A, 1, generator=g) * 800, torch.rand(A, 2, generator=g) * 200 + 20,
                      torch.rand(A, 2, generator=g)], 1).to(dev).contiguous()
    ws = ops.PostWorkspace(A, dev)
    full, cnt = ops.postprocess_device(pred, 1, 0.2, 0.65, ws)
    n_full = int(cnt.item())
    full = full[:n_full].clone()
    for k in (1, 3, 300):
        ws2 = ops.PostWorkspace(A, dev)
        part, c2 = ops.postprocess_device(pred, 1, 0.2, 0.65, ws2, max_keep=k)
        assert int(c2.item()) == min(k, n_full)
        assert torch.equal(part[:min(k, n_full)], full[:min(k, n_full)])


def test_msda_dropin_module_signature():
    """The reference's autograd wrapper calls MSDA.ms_deform_attn_forward(value, shapes, lsi, loc, attn, im2col_step)
    (ops/functions/ms_deform_attn_func.py:22-28)."""
    from unicorn_b200 import compat
    compat.install()
    import MultiScaleDeformableAttention as MSDA
    g = G(12)
    hw = [(9, 7), (5, 4)]
    S = sum(h * w for h, w in hw)
    value = torch.randn(2, S, 4, 16, generator=g).to(dev)
    loc = torch.rand(2, 11, 4, 2, 3, 2, generator=g).to(dev)
    attn = torch.softmax(torch.randn(2, 11, 4, 6, generator=g), -1).view(2, 11, 4, 2, 3).to(dev)
    shapes = torch.as_tensor(hw, dtype=torch.long, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    out = MSDA.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
    close(out, _msda_ref(value, hw, loc, attn), 1e-5, "msda drop-in")
    with pytest.raises(RuntimeError):
        MSDA.ms_deform_attn_forward(value.cpu(), shapes, lsi, loc, attn, 64)
    with pytest.raises(NotImplementedError):
        MSDA.ms_deform_attn_backward()


def test_stem_uint8_hwc_equals_float_nchw():
    from unicorn_b200 import ops
    g = G(13)
    C0, H, W = 96, 64, 96
    u8 = (torch.rand(1, H, W, 3, generator=g) * 255).to(torch.uint8).to(dev)
    f32 = u8.float().permute(0, 3, 1, 2).contiguous()
    w = (torch.randn(C0, 3, 4, 4, generator=g) / 7).to(dev)
    b, lw, lb = (torch.randn(C0, generator=g).to(dev) for _ in range(3))
    a = ops.stem_ln(u8, ops.pack_stem_weight(w), b, lw, lb)
    c = ops.stem_ln(f32, ops.pack_stem_weight(w), b, lw, lb)
    assert torch.equal(a, c)
