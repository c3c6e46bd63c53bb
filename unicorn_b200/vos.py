"""VOS per-frame driver on the B200 engine — mirrors external/lib/test/tracker/unicorn_vos.py (track :71-127,
get_mask_results :129-155, get_det_results :157-201) for objects given in the first frame: one backbone pass and one
fused correlation per frame (all objects' label maps are propagated by a single uc_corr_propagate launch, n_obj <= 8),
then per object: prior pyramid -> mask head -> NMS -> dynamic-conv mask of the best instance."""
import torch

from . import ops
from .engine import UnicornEngine
from .sot import get_label_map


class UnicornVOSTrack:
    def __init__(self, engine: UnicornEngine, input_size, conf=0.001, nms=0.65, max_inst=3, d_rate=2):
        assert engine.cfg["mask"], "VOS needs a *_mask model"
        self.eng, self.input_size = engine, tuple(input_size)
        self.conf, self.nms, self.max_inst, self.d_rate = conf, nms, max_inst, d_rate
        H, W = self.input_size
        A = (H // 8) * (W // 8) + (H // 16) * (W // 16) + (H // 32) * (W // 32)
        self.ws = ops.PostWorkspace(A, engine.dev)
        self.img_in = torch.empty(1, 3, H, W, dtype=torch.float32, device=engine.dev)

    def initialize_tensor(self, ref_frame, boxes_xyxy):
        """boxes_xyxy: dict obj_id -> box in resized-image coordinates (unicorn_vos.py:60-66)."""
        e = self.eng
        H, W = self.input_size
        self.img_in.copy_(ref_frame)
        e.begin_frame()
        _, seq = e.backbone(self.img_in, tag="ref")
        self.ref_feat = seq["feat"]
        self.obj_ids = list(boxes_xyxy.keys())
        assert 1 <= len(self.obj_ids) <= 8
        maps = [ops.bilinear(get_label_map(boxes_xyxy[o], H, W, e.dev), H // 8, W // 8, 8.0, 8.0).reshape(1, -1) for o in self.obj_ids]
        self.lbs_pre = torch.cat(maps, 0).contiguous()
        torch.cuda.synchronize()

    def track_tensor(self, cur_frame):
        """Returns {obj_id: (det_row [7] cpu or None, mask fp32 [H,W] device or None)} for the best instance."""
        e = self.eng
        H, W = self.input_size
        self.img_in.copy_(cur_frame)
        e.begin_frame()
        fpn, seq = e.backbone(self.img_in, tag="cur")
        f_pre, f_cur = e.interaction(self.ref_feat, seq["feat"])
        e_pre, e_cur = e.upsample(f_pre, "embp"), e.upsample(f_cur, "embc")
        K = len(self.obj_ids)
        hh, ww = H // 8, W // 8
        coarse = ops.corr_propagate(e_pre.view(-1, 128), e_cur.view(-1, 128), self.lbs_pre, out=e.buf("vos.coarse", (K, hh * ww), torch.float32))
        mf, um = e.mask_branch(fpn)  # identical for every object: computed once
        out = {}
        self.last = dict(mask_feats=mf, up_masks=um, coarse=coarse, per_obj={})
        for i, oid in enumerate(self.obj_ids):
            c0 = coarse[i:i + 1].view(1, hh, ww)
            pri = (c0, ops.bilinear(c0, hh // 2, ww // 2, 2.0, 2.0), ops.bilinear(c0, hh // 4, ww // 4, 4.0, 4.0))
            head = e.head(fpn, pri, "sot", with_masks=True)
            ops.postprocess_device(head[0], 1, self.conf, self.nms, self.ws, max_keep=self.max_inst)
            hw = [(t.shape[1], t.shape[2]) for t in e.dyn_levels]
            masks = ops.dynamic_masks(mf, um, e.dyn_levels, hw, self.ws, 1, up_rate=8 // self.d_rate, d_rate=self.d_rate)
            n = int(self.ws.count.item())
            self.last["per_obj"][oid] = dict(head=head.clone(), dyn=[t.clone() for t in e.dyn_levels])
            out[oid] = (self.ws.dets[0].cpu(), masks[0].clone()) if n > 0 else (None, None)
        return out
