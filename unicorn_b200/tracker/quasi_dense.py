"""Quasi-dense embedding tracker on the B200 path — same constructor and `match` signature as the reference's
unicorn/tracker/quasi_dense_embed_tracker.py (QuasiDenseEmbedTracker.match :137-212, update_memo :47-102).

The two dense pieces of the association — the pairwise IoU matrices (duplicate removal, backdrop NMS) and the
bi-softmax embedding similarity E·Mᵀ — run as sm_100a kernels (uc_box_iou, uc_bisoftmax); the memo is kept as
stacked tensors instead of a dict of dicts; the greedy assignment (inherently sequential, ≤ a few hundred rows)
stays on the host, as in the reference.  Inputs may be CPU or CUDA tensors; outputs are CPU tensors like the
reference's (the eval loop writes them to text files)."""
import torch

from .. import ops
from ._stream import assoc_stream


class QuasiDenseEmbedTracker:
    def __init__(self, init_score_thr=0.8, obj_score_thr=0.5, match_score_thr=0.5, memo_tracklet_frames=30,
                 memo_backdrop_frames=1, memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3,
                 nms_class_iou_thr=0.7, with_cats=True, match_metric="bisoftmax", device="cuda"):
        assert 0 <= memo_momentum <= 1.0 and memo_tracklet_frames >= 0 and memo_backdrop_frames >= 0
        if match_metric != "bisoftmax":
            raise NotImplementedError("only the bisoftmax metric (the reference default) is implemented")
        self.init_score_thr, self.obj_score_thr, self.match_score_thr = init_score_thr, obj_score_thr, match_score_thr
        self.memo_tracklet_frames, self.memo_backdrop_frames, self.memo_momentum = memo_tracklet_frames, memo_backdrop_frames, memo_momentum
        self.nms_conf_thr, self.nms_backdrop_iou_thr, self.nms_class_iou_thr = nms_conf_thr, nms_backdrop_iou_thr, nms_class_iou_thr
        self.with_cats = with_cats
        self.dev = torch.device(device)
        self.num_tracklets = 0
        # tracklet memo, one row per live id (insertion order == the reference's dict order)
        self.t_ids = torch.zeros(0, dtype=torch.long)
        self.t_box = torch.zeros(0, 5)
        self.t_emb = None
        self.t_lab = torch.zeros(0)
        self.t_last = torch.zeros(0, dtype=torch.long)
        self.t_vel = torch.zeros(0, 5)
        self.t_acc = torch.zeros(0, dtype=torch.long)
        self.backdrops = []  # newest first: (boxes, embeds, labels)

    @property
    def empty(self):
        return self.t_ids.numel() == 0

    # ---------------------------------------------------------------------------------------------- device helpers
    def _iou(self, a, b):
        if a.size(0) == 0 or b.size(0) == 0:
            return torch.zeros(a.size(0), b.size(0))
        with torch.cuda.stream(assoc_stream(self.dev)):  # not behind the next frame's kernels on the main stream
            return ops.box_iou(a.to(self.dev, torch.float32).contiguous(), b.to(self.dev, torch.float32).contiguous()).cpu()

    def _scores(self, embeds, labels, m_embeds, m_labels):
        with torch.cuda.stream(assoc_stream(self.dev)):
            e = embeds.to(self.dev, torch.float32).contiguous()
            m = m_embeds.to(self.dev, torch.float32).contiguous()
            ld = labels.to(self.dev, torch.float32).contiguous() if self.with_cats else None
            lm = m_labels.to(self.dev, torch.float32).contiguous() if self.with_cats else None
            return ops.bisoftmax(e, m, ld, lm).cpu()

    # ---------------------------------------------------------------------------------------------- match
    def match(self, bboxes, labels, track_feats, frame_id, asso_tau=-1, return_index=False):
        # the caller's label tensor (dtype, device) is what comes back, like the reference which only indexes it; the float copy
        # below feeds the class gate of the score kernel and the memo
        labels_in, out_dev = labels, bboxes.device
        bboxes, labels, track_feats = bboxes.detach().cpu().float(), labels.detach().cpu().float(), track_feats.detach().cpu().float()
        order = bboxes[:, -1].sort(descending=True)[1]
        bboxes, labels, embeds = bboxes[order], labels[order], track_feats[order]
        labels_in = labels_in[order.to(labels_in.device)]
        n = bboxes.size(0)
        # duplicate removal: a box is dropped if ANY higher-scored box (kept or not) overlaps it above its threshold
        valids = torch.ones(n, dtype=torch.bool)
        if n > 1:
            iou = self._iou(bboxes[:, :4], bboxes[:, :4])
            thr = torch.where(bboxes[:, -1] < self.obj_score_thr, torch.tensor(self.nms_backdrop_iou_thr), torch.tensor(self.nms_class_iou_thr))
            over = torch.tril(iou > thr[:, None], diagonal=-1)  # row i vs columns < i
            valids = ~over.any(dim=1)
        bboxes, labels, embeds = bboxes[valids], labels[valids], embeds[valids]
        labels_in = labels_in[valids.to(labels_in.device)]
        n = bboxes.size(0)
        ids = torch.full((n,), -1, dtype=torch.long)
        if n > 0 and not self.empty:
            m_emb = torch.cat([self.t_emb] + [b[1] for b in self.backdrops])
            m_lab = torch.cat([self.t_lab] + [b[2] for b in self.backdrops])
            m_ids = torch.cat([self.t_ids] + [torch.full((b[1].size(0),), -1, dtype=torch.long) for b in self.backdrops])
            scores = self._scores(embeds, labels, m_emb, m_lab).numpy().copy()
            det_score = bboxes[:, -1].numpy()
            for i in range(n):  # greedy, in detection-score order; a claimed memo column is zeroed for everybody else
                j = int(scores[i].argmax())
                conf = scores[i, j]
                tid = int(m_ids[j])
                if conf > self.match_score_thr and tid > -1:
                    if det_score[i] > self.obj_score_thr:
                        ids[i] = tid
                        scores[:i, j] = 0
                        scores[i + 1:, j] = 0
                    elif conf > self.nms_conf_thr:
                        ids[i] = -2
        new = (ids == -1) & (bboxes[:, 4] > self.init_score_thr)
        n_new = int(new.sum())
        ids[new] = torch.arange(self.num_tracklets, self.num_tracklets + n_new, dtype=torch.long)
        self.num_tracklets += n_new
        self._update_memo(ids, bboxes, embeds, labels, frame_id)
        if return_index:
            return bboxes.to(out_dev), labels_in, ids.to(out_dev), valids.to(out_dev)
        return bboxes.to(out_dev), labels_in, ids.to(out_dev)

    # ---------------------------------------------------------------------------------------------- memo
    def _update_memo(self, ids, bboxes, embeds, labels, frame_id):
        if self.t_emb is None:
            self.t_emb = torch.zeros(0, embeds.size(1))
        pos = {int(t): k for k, t in enumerate(self.t_ids.tolist())}
        add, rows, slots = [], [], []
        for r, tid in enumerate(ids.tolist()):
            if tid < 0:
                continue
            k = pos.get(tid)
            if k is None:
                add.append(r)
            else:
                rows.append(r)
                slots.append(k)
        if rows:  # every id occurs once per frame: the memo rows are updated together (same arithmetic as the per-track loop)
            r, k = torch.tensor(rows, dtype=torch.long), torch.tensor(slots, dtype=torch.long)
            vel = (bboxes[r] - self.t_box[k]) / (frame_id - self.t_last[k]).float()[:, None]
            self.t_box[k] = bboxes[r]
            self.t_emb[k] = (1 - self.memo_momentum) * self.t_emb[k] + self.memo_momentum * embeds[r]
            self.t_last[k] = frame_id
            self.t_lab[k] = labels[r]
            acc = self.t_acc[k].float()[:, None]
            self.t_vel[k] = (self.t_vel[k] * acc + vel) / (acc + 1)
            self.t_acc[k] += 1
        if add:
            a = torch.tensor(add, dtype=torch.long)
            self.t_ids = torch.cat([self.t_ids, ids[a]])
            self.t_box = torch.cat([self.t_box, bboxes[a]])
            self.t_emb = torch.cat([self.t_emb, embeds[a]])
            self.t_lab = torch.cat([self.t_lab, labels[a]])
            self.t_last = torch.cat([self.t_last, torch.full((len(add),), frame_id, dtype=torch.long)])
            self.t_vel = torch.cat([self.t_vel, torch.zeros(len(add), bboxes.size(1))])
            self.t_acc = torch.cat([self.t_acc, torch.zeros(len(add), dtype=torch.long)])
        # backdrops: unmatched (-1) boxes not overlapped (> thr) by any earlier box of this frame
        bd = torch.nonzero(ids == -1, as_tuple=False).squeeze(1)
        if bd.numel():
            iou = self._iou(bboxes[bd, :4], bboxes[:, :4])
            col = torch.arange(bboxes.size(0))[None, :]
            hit = ((iou > self.nms_backdrop_iou_thr) & (col < bd[:, None])).any(dim=1)
            bd = bd[~hit]
        self.backdrops.insert(0, (bboxes[bd], embeds[bd], labels[bd]))
        alive = (frame_id - self.t_last) < self.memo_tracklet_frames
        if not bool(alive.all()):
            self.t_ids, self.t_box, self.t_emb, self.t_lab = self.t_ids[alive], self.t_box[alive], self.t_emb[alive], self.t_lab[alive]
            self.t_last, self.t_vel, self.t_acc = self.t_last[alive], self.t_vel[alive], self.t_acc[alive]
        if len(self.backdrops) > self.memo_backdrop_frames:
            self.backdrops.pop()
