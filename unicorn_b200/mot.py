"""MOT per-frame driver on the B200 engine — the per-frame body of MOTEvaluator.evaluate_omni
(unicorn/evaluators/mot_evaluator.py:985-1057): `model(imgs, mode="whole")` -> postprocess -> score filter ->
interaction with the previous frame -> embedding upsample (current frame only) -> embedding sampling at the box
centres -> QuasiDenseEmbedTracker.match.  The reference's per-box Python grid_sample loop, deepcopy of the frame
dict and empty_cache() calls are gone; the previous frame's projected tokens are kept in the encoder's token buffer
(rows of level 0) by swapping two token buffers instead of re-projecting."""
import torch

from . import ops
from .engine import UnicornEngine
from .tracker import QuasiDenseEmbedTracker


class UnicornMOTTracker:
    def __init__(self, engine: UnicornEngine, input_size, conf=0.01, nms=0.7, score_thr=0.1, max_dets=1024, tracker=None):
        self.eng, self.input_size = engine, tuple(input_size)
        self.conf, self.nms, self.score_thr, self.max_dets = conf, nms, score_thr, max_dets
        self.tracker = tracker or QuasiDenseEmbedTracker(device=engine.dev)
        H, W = self.input_size
        A = (H // 8) * (W // 8) + (H // 16) * (W // 16) + (H // 32) * (W // 32)
        self.ws = ops.PostWorkspace(A, engine.dev)
        self.img_in = torch.empty(1, 3, H, W, dtype=torch.float32, device=engine.dev)
        self.frame_id = 0
        self._prev_feat = None

    def step_tensor(self, frame, scale=1.0):
        """frame: preprocessed fp32 [1,3,H,W].  Returns (bboxes [n,5] in original-image coordinates, ids [n])."""
        e = self.eng
        self.frame_id += 1
        self.img_in.copy_(frame, non_blocking=True)
        e.begin_frame()
        tag = "mot%d" % (self.frame_id & 1)  # two buffer sets: the previous frame's s16 feature must survive
        fpn, seq = e.backbone(self.img_in, tag=tag)
        out = e.head(fpn, None, "mot")  # whole mode: zero priors (unicorn.py:133-139)
        dets, cnt = ops.postprocess_device(out[0], e.ncls, self.conf, self.nms, self.ws)
        prev = self._prev_feat if self._prev_feat is not None else seq["feat"]  # frame 1: pre_dict = cur_dict (:1014-1015)
        _, f_cur = e.interaction(prev, seq["feat"])
        emb = e.upsample(f_cur, "mot.emb")
        self._prev_feat = seq["feat"]
        n_max = self.max_dets
        feats = ops.sample_embed(emb, dets, n_max, 8.0, count=cnt)
        n = min(int(cnt.item()), n_max)
        d = dets[:n].cpu()
        f = feats[:n].cpu()
        scores = d[:, 4] * d[:, 5]
        keep = scores > self.score_thr  # :1008-1012
        boxes = torch.cat([d[keep, :4] / scale, scores[keep, None]], 1)
        labels = torch.ones(boxes.size(0))  # :1013 (all labels = 1)
        self.last = dict(dets=d, feats=f, embed=emb, head=out)
        if boxes.size(0) == 0:
            return torch.zeros(0, 5), torch.zeros(0, dtype=torch.long)
        ob, _, oid = self.tracker.match(boxes, labels, f[keep], self.frame_id)
        valid = oid > -1  # :1047-1053
        ob, oid = ob[valid], oid[valid]
        order = oid.sort()[1]
        return ob[order], oid[order]
