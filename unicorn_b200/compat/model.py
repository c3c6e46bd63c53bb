"""Reference-shaped facade over UnicornEngine: the call conventions the reference's drivers use on `Unicorn`
(unicorn/models/unicorn.py:60-139; SURVEY.md 8b "Python model API to keep"), with the reference's tensor formats at the
boundary (NCHW fp32 in, NCHW fp32 out) so that `unicorn_sot.py` / `mot_evaluator.py`-style code runs on the B200 path
unchanged:

    model = UnicornB200Model(state_dict, "unicorn_track_large")
    fpn_outs, seq_dict = model(imgs=x, mode="backbone")                         # unicorn.py:97-101
    f0, f1 = model(seq_dict0=a, seq_dict1=b, mode="interaction")               # unicorn.py:102-110
    emb = model(feat=f1, mode="upsample")                                      # unicorn.py:111-113
    out, seq_dict = model(imgs=x, mode="whole")                                # unicorn.py:133-139
    out = model.head(fpn_outs, prior_pyramid, mode="sot" | "mot")             # unicorn_head.py:249-336
    dets = postprocess(out, num_classes, conf_thre, nms_thre)                  # utils/boxes.py:33-77

Inside, everything runs on the engine's NHWC bf16 kernels; the conversions at the boundary are exact in the outbound
direction (bf16 -> fp32) and in the inbound direction for tensors that came out of this facade (their fp32 values are bf16
numbers).  `seq_dict` holds plain tensors (keys feat, pos, h, w) and survives copy.deepcopy (mot_evaluator.py:1015).  The
fused fast paths (correlation without the N x N matrix, CUDA-graph frames) live in the driver classes
(unicorn_b200.sot / mot / vos), which is where a per-frame loop should go; this class is the drop-in for code that calls
the model stage by stage.  There is no CPU path: tensors must be CUDA tensors on the engine's device."""

from .. import ops
from ..engine import UnicornEngine


def _nhwc(x):
    assert x.is_cuda and x.dim() == 4, "UnicornB200Model: CUDA NCHW tensors only (no CPU fallback)"
    return ops.nchw_to_nhwc(x.float().contiguous())


class _Head:
    def __init__(self, model):
        self._m = model
        self.decode_in_inference = True

    def __call__(self, fpn_outs, prior_ms=None, mode="sot"):
        """fpn_outs: 3 NCHW fp32 maps; prior_ms: 3 fp32 maps (1,1,h,w) or None / all-zero for "mot".  -> (1, A, 5+ncls)."""
        e = self._m.engine
        fpn = [_nhwc(t) for t in fpn_outs]
        pri = None
        if prior_ms is not None and mode == "sot":
            pri = [p.float().reshape(1, p.shape[-2], p.shape[-1]).contiguous() for p in prior_ms]
        return e.head(fpn, pri, mode).clone()


class UnicornB200Model:
    def __init__(self, state_dict, cfg_name, device="cuda"):
        self.engine = UnicornEngine(state_dict, cfg_name, device=device)
        self.head = _Head(self)
        self.num_classes = self.engine.ncls

    def eval(self):
        return self

    def cuda(self, *a, **k):
        return self

    def half(self):  # the reference's --fp16 switch: the engine already computes in bf16 / fp16
        return self

    def _backbone(self, imgs):
        e = self.engine
        assert imgs.is_cuda and imgs.dim() == 4 and imgs.shape[0] == 1, "one frame per call (the reference's tracking drivers use batch 1)"
        e.begin_frame()
        fpn, seq = e.backbone(imgs.float().contiguous(), tag="compat")
        h, w = seq["h"], seq["w"]
        seq_dict = {"feat": ops.nhwc_to_nchw(seq["feat"]), "pos": e.pos_tokens(h, w)[1].clone(), "h": h, "w": w}
        return fpn, seq_dict

    def __call__(self, imgs=None, seq_dict0=None, seq_dict1=None, feat=None, mode="backbone", **unused):
        e = self.engine
        if mode == "backbone":
            fpn, seq_dict = self._backbone(imgs)
            return tuple(ops.nhwc_to_nchw(t) for t in fpn), seq_dict
        if mode == "interaction":
            f0, f1 = e.interaction(_nhwc(seq_dict0["feat"]), _nhwc(seq_dict1["feat"]))
            return ops.nhwc_to_nchw(f0), ops.nhwc_to_nchw(f1)
        if mode == "upsample":
            return ops.nhwc_to_nchw(e.upsample(_nhwc(feat), "compat"))
        if mode == "whole":  # backbone + head with zero priors, MOT prediction set (unicorn.py:133-139)
            fpn, seq_dict = self._backbone(imgs)
            return e.head(fpn, None, "mot").clone(), seq_dict
        raise ValueError(f"UnicornB200Model: unsupported mode {mode!r} (inference modes: backbone, interaction, upsample, whole)")


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45):
    """unicorn.utils.postprocess (utils/boxes.py:33-77): list with one (M,7) tensor of rows
    (x1,y1,x2,y2,obj_conf,class_conf,class_pred), descending score, or None when nothing passes — on the GPU."""
    out = []
    for p in prediction:
        p = p.float().contiguous()
        ws = ops.PostWorkspace(p.shape[0], p.device)
        dets, cnt = ops.postprocess_device(p, num_classes, conf_thre, nms_thre, ws)
        n = int(cnt.item())
        out.append(dets[:n].clone() if n > 0 else None)
    return out
