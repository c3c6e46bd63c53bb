"""CPU oracle of the input preprocessing — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).

Restates PreprocessorX.process (external/lib/test/tracker/unicorn_sot.py:114-123) and preproc
(unicorn/data/data_augment.py:194-214): r = min(Hin/h, Win/w); cv2.resize(img, (int(w*r), int(h*r)), INTER_LINEAR) into the
top-left corner of a frame filled with 114 (SOT: after cv2.COLOR_RGB2BGR).

cv2.resize is a third-party dependency of the reference (opencv-python, unpinned in requirements.txt); its 8-bit INTER_LINEAR
is restated here from OpenCV's published algorithm (modules/imgproc/src/resize.cpp: resizeGeneric_ with HResizeLinear /
VResizeLinear<uchar,int,short,FixedPtCast<..., INTER_RESIZE_COEF_BITS*2>>, INTER_RESIZE_COEF_BITS = 11):
  * per axis: f = float((d + 0.5) * scale - 0.5), s = floor(f), f -= s, scale = 1 / (dsize / ssize) in double;
  * x axis: s < 0 -> (s, f) = (0, 0); s >= W-1 -> (W-1, 0); y axis: f is kept, the source rows s and s+1 are clamped;
  * coefficients (short) rint((1-f) * 2048), rint(f * 2048);
  * horizontal pass h = S[sx]*a0 + S[sx+1]*a1 (int32); vertical (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.
Pinned: tests/test_preprocess.py checks it bit for bit against the cv2 installed in the image (4.13) and against the
committed fixture tests/golden/letterbox.npz (generated with that cv2 by tests/golden/make_letterbox_golden.py)."""
import numpy as np


def _axis(ssize, dsize, clamp):
    scale = 1.0 / (float(dsize) / ssize)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= ssize - 1
        f[hi], s[hi] = 0, ssize - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    c1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, ssize - 1), np.clip(s + 1, 0, ssize - 1), c0, c1


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    H, W = src.shape[:2]
    if (dh, dw) == (H, W):
        return src.copy()
    sx0, sx1, a0, a1 = _axis(W, dw, True)
    sy0, sy1, b0, b1 = _axis(H, dh, False)
    S = src.astype(np.int32)
    hor = S[:, sx0] * a0[None, :, None] + S[:, sx1] * a1[None, :, None]
    h0, h1 = hor[sy0], hor[sy1]
    out = (((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox(img, input_size, swap_rb, pad=114):
    """-> (uint8 [Hin, Win, 3], r).  swap_rb=True is the SOT/VOS preprocessor (RGB in, BGR out); False is preproc()."""
    h, w = img.shape[:2]
    r = min(input_size[0] / h, input_size[1] / w)
    rh, rw = int(h * r), int(w * r)
    out = np.full((input_size[0], input_size[1], 3), pad, dtype=np.uint8)
    rs = resize_linear_u8(img[:, :, ::-1] if swap_rb else img, rw, rh)
    out[:rh, :rw] = rs
    return out, r
