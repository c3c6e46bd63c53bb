"""MOTS per-frame driver (UNTESTED ON A GPU — written after the round-1 GPU budget was spent; see tests/test_mots_gpu.py):
the per-frame body of MOTEvaluator.evaluate_omni_mots (unicorn/evaluators/mot_evaluator.py:776-897) on the B200 engine:
whole-mode detector with the CondInst controllers -> NMS -> dynamic-conv masks of the kept detections -> embedding
sampling -> QuasiDenseEmbedTracker.match(return_index=True) -> masks of the tracked boxes in ascending-id order,
overlap free, area filter, RLE (results.mots_frame_result)."""
import torch
import torch.nn.functional as F

from . import ops
from .engine import UnicornEngine
from .results import mots_frame_result
from .tracker import QuasiDenseEmbedTracker


class UnicornMOTSTracker:
    def __init__(self, engine: UnicornEngine, input_size, conf=0.01, nms=0.7, score_thr=0.1, max_dets=64, mask_thres=0.3, d_rate=2,
                 min_box_area=100, tracker=None):
        assert engine.cfg["mask"], "MOTS needs a *_mask model"
        self.eng, self.input_size = engine, tuple(input_size)
        self.conf, self.nms, self.score_thr, self.max_dets = conf, nms, score_thr, max_dets
        self.mask_thres, self.d_rate, self.min_box_area = mask_thres, d_rate, min_box_area
        self.tracker = tracker or QuasiDenseEmbedTracker(device=engine.dev)
        H, W = self.input_size
        A = (H // 8) * (W // 8) + (H // 16) * (W // 16) + (H // 32) * (W // 32)
        self.ws = ops.PostWorkspace(A, engine.dev)
        self.img_in = torch.empty(1, 3, H, W, dtype=torch.float32, device=engine.dev)
        self.feats = torch.zeros(max_dets, 128, dtype=torch.float32, device=engine.dev)
        self.frame_id = 0
        self._prev_feat = torch.zeros(1, H // 16, W // 16, engine.dims[2], dtype=torch.bfloat16, device=engine.dev)
        self._has_prev = torch.zeros(1, dtype=torch.int32, device=engine.dev)
        self.last = {}

    def step_tensor(self, frame, img_h, img_w):
        """frame: preprocessed fp32 [1,3,H,W]; (img_h, img_w): original image size.  Returns the tuple write_results_mots()
        consumes for this frame: (frame_id, ids (1-based), cat_id, img_h, img_w, rles)."""
        e = self.eng
        H, W = self.input_size
        self.frame_id += 1
        self.img_in.copy_(frame, non_blocking=True)
        e.begin_frame()
        fpn, seq = e.backbone(self.img_in, tag="mots%d" % (self.frame_id & 1))
        out = e.head(fpn, None, "mot", with_masks=True)
        dets, cnt = ops.postprocess_device(out[0], e.ncls, self.conf, self.nms, self.ws)
        mf, um = e.mask_branch(fpn)
        hw = [(t.shape[1], t.shape[2]) for t in e.dyn_levels]
        masks = ops.dynamic_masks(mf, um, e.dyn_levels, hw, self.ws, self.max_dets, up_rate=8 // self.d_rate, d_rate=self.d_rate)
        ops.copy_rows_if(self._has_prev, seq["feat"], self._prev_feat, invert=True)  # first frame with detections: pre_dict = cur_dict (:812-813)
        _, f_cur = e.interaction(self._prev_feat, seq["feat"])
        emb = e.upsample(f_cur, "mots.emb")
        ops.sample_embed(emb, dets, self.max_dets, 8.0, count=cnt, out=self.feats)
        ops.copy_rows_if(cnt, seq["feat"], self._prev_feat)  # pre_dict advances only on frames with detections (:803,818)
        self._has_prev.bitwise_or_((cnt > 0).to(torch.int32))
        n = min(int(cnt.item()), self.max_dets)
        d, f = dets[:n].cpu(), self.feats[:n].cpu()
        scale = min(H / float(img_h), W / float(img_w))
        # masks at the original image scale, thresholded (:804-805)
        m = F.interpolate(masks[:n, None], scale_factor=1 / scale, mode="bilinear", align_corners=False)[:, 0, :img_h, :img_w] > self.mask_thres
        scores = d[:, 4] * d[:, 5]
        keep = scores > self.score_thr
        boxes = torch.cat([d[keep, :4] / scale, scores[keep, None]], 1)
        m, f = m[keep.to(m.device)], f[keep]
        self.last = dict(dets=d, masks=masks[:n], head=out, mask_feats=mf, up_masks=um, dyn=[t for t in e.dyn_levels])
        if n == 0:  # outputs[0] is None: no tracking for this frame (mot_evaluator.py:803)
            return self.frame_id, [], 2, img_h, img_w, []
        ob, _, oid, idx = self.tracker.match(boxes, torch.ones(boxes.size(0)), f, self.frame_id, return_index=True)
        m = m[idx.to(m.device)]
        valid = oid > -1
        return mots_frame_result(self.frame_id, ob[valid], oid[valid], m[valid.to(m.device)].cpu(), img_h, img_w, self.min_box_area)
