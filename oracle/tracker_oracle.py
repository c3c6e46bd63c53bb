"""CPU oracle for the MOT association step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain torch restatement of unicorn/tracker/quasi_dense_embed_tracker.py (QuasiDenseEmbedTracker.match :137-212,
update_memo :47-102, memo :104-135) and of the per-frame glue of unicorn/evaluators/mot_evaluator.py:1005-1045
(score filter, interaction with the previous frame, embedding sampling at box centres, match).
Pinned: tests/golden/make_golden_tracker.py runs the UNMODIFIED reference tracker class on seeded detection
sequences and stores the ids (tests/golden/qd_tracker.npz); tests/test_tracker_oracle.py re-checks this file.
"""
import torch
import torch.nn.functional as F


def box_iou(a, b):
    """torchvision.ops.box_iou (the reference's choice, quasi_dense_embed_tracker.py:6)."""
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area1[:, None] + area2[None, :] - inter)


class QDTrackerOracle:
    """State: tracklets {id: bbox(5), embed, label, last_frame, velocity, acc_frame}; backdrops list (newest first)."""

    def __init__(self, init_score_thr=0.8, obj_score_thr=0.5, match_score_thr=0.5, memo_tracklet_frames=30,
                 memo_backdrop_frames=1, memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3,
                 nms_class_iou_thr=0.7, with_cats=True):  # quasi_dense_embed_tracker.py:11-22
        self.p = dict(init=init_score_thr, obj=obj_score_thr, match=match_score_thr, tf=memo_tracklet_frames,
                      bf=memo_backdrop_frames, mom=memo_momentum, nmsc=nms_conf_thr, bd_iou=nms_backdrop_iou_thr,
                      cls_iou=nms_class_iou_thr, cats=with_cats)
        self.num_tracklets = 0
        self.tracklets = {}
        self.backdrops = []

    def _memo(self):  # :104-135
        boxes = [v["bbox"][None] for v in self.tracklets.values()]
        embeds = [v["embed"][None] for v in self.tracklets.values()]
        labels = [v["label"].view(1) for v in self.tracklets.values()]
        ids = [torch.tensor(list(self.tracklets.keys()), dtype=torch.long)]
        for bd in self.backdrops:
            boxes.append(bd["bboxes"]); embeds.append(bd["embeds"]); labels.append(bd["labels"])
            ids.append(torch.full((bd["embeds"].size(0),), -1, dtype=torch.long))
        return torch.cat(boxes), torch.cat(labels), torch.cat(embeds), torch.cat(ids)

    def _filter_and_scores(self, bboxes, labels, track_feats):
        """:139-170 without touching the state: score-sorted, duplicate-filtered detections and their bi-softmax score matrix
        against the memo (None when there is nothing to match)."""
        P = self.p
        order = bboxes[:, -1].sort(descending=True)[1]  # :139-142
        bboxes, labels, embeds = bboxes[order], labels[order], track_feats[order]
        valid = torch.ones(bboxes.size(0), dtype=torch.bool)  # :145-152 duplicate removal
        ious = box_iou(bboxes[:, :4], bboxes[:, :4])
        for i in range(1, bboxes.size(0)):
            thr = P["bd_iou"] if bboxes[i, -1] < P["obj"] else P["cls_iou"]
            if (ious[i, :i] > thr).any():
                valid[i] = False
        bboxes, labels, embeds = bboxes[valid], labels[valid], embeds[valid]
        scores, m_ids = None, None
        if bboxes.size(0) > 0 and self.tracklets:  # :161 (`empty` looks at tracklets only)
            m_boxes, m_labels, m_embeds, m_ids = self._memo()
            feats = embeds @ m_embeds.t()  # :166-170 bi-softmax
            scores = (feats.softmax(dim=1) + feats.softmax(dim=0)) / 2
            if P["cats"]:
                scores = scores * (labels.view(-1, 1) == m_labels.view(1, -1)).float()
        return bboxes, labels, embeds, scores, m_ids

    def decision_margin(self, bboxes, labels, track_feats):
        """(score matrix, smallest distance of a row's decision from flipping): per detection row the lead of its best memo entry over
        the second best and the distance of the best score from the thresholds it is compared with (:188-199) — what a perturbation
        of the scores has to exceed to change an id.  (None, inf) when nothing is matched."""
        P = self.p
        _, _, _, scores, _ = self._filter_and_scores(bboxes, labels, track_feats)
        if scores is None or scores.numel() == 0:
            return None, float("inf")
        top = scores.topk(min(2, scores.shape[1]), dim=1)[0]
        lead = top[:, 0] - (top[:, 1] if top.shape[1] > 1 else 0.0)
        thr = torch.minimum((top[:, 0] - P["match"]).abs(), (top[:, 0] - P["nmsc"]).abs())
        return scores, float(torch.minimum(lead, thr).min())

    def match(self, bboxes, labels, track_feats, frame_id):
        P = self.p
        bboxes, labels, embeds, scores, m_ids = self._filter_and_scores(bboxes, labels, track_feats)
        ids = torch.full((bboxes.size(0),), -1, dtype=torch.long)
        if scores is not None:
            for i in range(bboxes.size(0)):  # :188-199 greedy in score order with column zeroing
                conf, j = torch.max(scores[i], dim=0)
                tid = m_ids[j]
                if conf > P["match"]:
                    if tid > -1:
                        if bboxes[i, -1] > P["obj"]:
                            ids[i] = tid
                            scores[:i, j] = 0
                            scores[i + 1:, j] = 0
                        elif conf > P["nmsc"]:
                            ids[i] = -2
        new = (ids == -1) & (bboxes[:, 4] > P["init"])  # :200-206
        n_new = int(new.sum())
        ids[new] = torch.arange(self.num_tracklets, self.num_tracklets + n_new, dtype=torch.long)
        self.num_tracklets += n_new
        self._update(ids, bboxes, embeds, labels, frame_id)
        return bboxes, labels, ids

    def _update(self, ids, bboxes, embeds, labels, frame_id):  # :47-102
        P = self.p
        for tid, bbox, embed, label in zip(ids.tolist(), bboxes, embeds, labels):
            if tid < 0:
                continue
            t = self.tracklets.get(tid)
            if t is not None:
                vel = (bbox - t["bbox"]) / (frame_id - t["last_frame"])
                t["bbox"] = bbox
                t["embed"] = (1 - P["mom"]) * t["embed"] + P["mom"] * embed
                t["last_frame"] = frame_id
                t["label"] = label
                t["velocity"] = (t["velocity"] * t["acc_frame"] + vel) / (t["acc_frame"] + 1)
                t["acc_frame"] += 1
            else:
                self.tracklets[tid] = dict(bbox=bbox, embed=embed, label=label, last_frame=frame_id,
                                           velocity=torch.zeros_like(bbox), acc_frame=0)
        bd = torch.nonzero(ids == -1, as_tuple=False).squeeze(1)
        ious = box_iou(bboxes[bd, :4], bboxes[:, :4])
        keep = [int(ind) for i, ind in enumerate(bd) if not (ious[i, :ind] > P["bd_iou"]).any()]
        keep = torch.tensor(keep, dtype=torch.long)
        self.backdrops.insert(0, dict(bboxes=bboxes[keep], embeds=embeds[keep], labels=labels[keep]))
        for k in [k for k, v in self.tracklets.items() if frame_id - v["last_frame"] >= P["tf"]]:
            self.tracklets.pop(k)
        if len(self.backdrops) > P["bf"]:
            self.backdrops.pop()


def sample_embeddings(embed, bboxes, img_size, s=8):
    """mot_evaluator.py:1024-1034: embed (1,C,h,w), bboxes (N,4) xyxy in network-input pixels -> (N,C)."""
    cx, cy = (bboxes[:, 0] + bboxes[:, 2]) / 2 / s - 0.5, (bboxes[:, 1] + bboxes[:, 3]) / 2 / s - 0.5
    cx = (torch.clamp(cx, min=0, max=img_size[1] // s - 1) / (img_size[1] // s - 1) - 0.5) * 2.0
    cy = (torch.clamp(cy, min=0, max=img_size[0] // s - 1) / (img_size[0] // s - 1) - 0.5) * 2.0
    grids = torch.stack([cx, cy], dim=-1)
    feats = [F.grid_sample(embed, g.view(1, 1, 1, 2), mode="bilinear", padding_mode="border", align_corners=False).squeeze()
             for g in grids]
    return torch.stack(feats, 0) if feats else torch.zeros((0, embed.size(1)))
