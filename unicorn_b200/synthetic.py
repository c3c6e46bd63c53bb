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
