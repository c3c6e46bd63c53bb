"""Seeded synthetic video: uniform-noise background plus one textured rectangle moving linearly,
already in the network's input format (PreprocessorX output: fp32 BGR in [0,255], 114-padded letterbox is a no-op
because frames are generated at the test size)."""
import torch


def make_video(n_frames, H, W, seed=0, n_obj=1):
    """Returns (frames [n,3,H,W] fp32 CPU, boxes [n, n_obj, 4] xyxy)."""
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    bg = torch.rand(3, H, W, generator=g) * 255.0
    frames = torch.empty(n_frames, 3, H, W)
    boxes = torch.zeros(n_frames, n_obj, 4)
    objs = []
    for o in range(n_obj):
        bw = int(W * (0.12 + 0.1 * torch.rand(1, generator=g).item()))
        bh = int(H * (0.15 + 0.1 * torch.rand(1, generator=g).item()))
        tex = torch.rand(3, bh, bw, generator=g) * 255.0
        # a smooth pattern on top of the noise so the object is distinguishable from the background
        yy = torch.linspace(0, 3.14159 * 3, bh)[:, None]
        xx = torch.linspace(0, 3.14159 * 3, bw)[None, :]
        tex = 0.5 * tex + 0.5 * (127.5 + 127.5 * torch.sin(yy + o) * torch.cos(xx))[None]
        x0 = torch.rand(1, generator=g).item() * (W - bw) * 0.5
        y0 = torch.rand(1, generator=g).item() * (H - bh) * 0.5
        vx = (W - bw) * 0.4 / max(n_frames - 1, 1)
        vy = (H - bh) * 0.4 / max(n_frames - 1, 1)
        objs.append((tex, bw, bh, x0, y0, vx, vy))
    for t in range(n_frames):
        f = bg.clone()
        f += (torch.rand(3, H, W, generator=g) - 0.5) * 8.0  # per-frame sensor noise
        for o, (tex, bw, bh, x0, y0, vx, vy) in enumerate(objs):
            x, y = int(round(x0 + vx * t)), int(round(y0 + vy * t))
            f[:, y:y + bh, x:x + bw] = tex
            boxes[t, o] = torch.tensor([x, y, x + bw, y + bh], dtype=torch.float32)
        frames[t] = f.clamp_(0, 255)
    return frames, boxes


def make_detections(n_frames=25, n_obj=14, seed=0, dim=128, W=1280.0, H=800.0):
    """Seeded synthetic per-frame detections for association tests: objects move linearly, each carries a noisy copy
    of its identity embedding; scores vary so that the tracker's thresholds (0.5 / 0.8) are exercised; a few
    duplicates and low-score clutter boxes are added.  Returns [(boxes [N,5] x1y1x2y2s, feats [N,dim]), ...]."""
    g = torch.Generator(device="cpu").manual_seed(4000 + seed)
    ident = torch.randn(n_obj, dim, generator=g) * 1.5
    pos = torch.rand(n_obj, 2, generator=g) * torch.tensor([W * 0.7, H * 0.7])
    vel = (torch.rand(n_obj, 2, generator=g) - 0.5) * 16
    size = torch.rand(n_obj, 2, generator=g) * 80 + 40
    out = []
    for t in range(n_frames):
        present = torch.rand(n_obj, generator=g) > 0.12
        p = pos + vel * t + torch.randn(n_obj, 2, generator=g) * 1.5
        boxes = torch.cat([p, p + size], 1)
        scores = (0.45 + 0.55 * torch.rand(n_obj, generator=g)).clamp(max=0.99)
        feats = ident + 0.25 * torch.randn(n_obj, dim, generator=g)
        b, s, f = boxes[present], scores[present], feats[present]
        # duplicate of the first present box (slightly shifted, lower score) and two clutter boxes
        if b.size(0) > 0:
            b = torch.cat([b, b[:1] + 3.0, torch.rand(2, 2, generator=g).repeat(1, 2) * 400 + torch.tensor([0, 0, 60.0, 60.0])])
            s = torch.cat([s, s[:1] * 0.9, torch.tensor([0.2, 0.35])])
            f = torch.cat([f, f[:1] + 0.1 * torch.randn(1, dim, generator=g), torch.randn(2, dim, generator=g)])
        out.append((torch.cat([b, s[:, None]], 1).float(), f.float()))
    return out
