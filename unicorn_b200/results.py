"""Result formats of the MOT / MOTS evaluators (SURVEY.md 8(f) "next" row 2): the txt writers of
unicorn/evaluators/mot_evaluator.py:37-72, the overlap-free mask post-processing of :858-866 and the COCO run-length
encoding the MOTS writer stores (pycocotools.mask.encode on a Fortran-ordered mask, :884-888).

pycocotools is a third-party dependency of the reference that is absent from this image: `rle_encode` restates the
published COCO mask API (cocoapi/common/maskApi.c: rleEncode + rleToString — column-major run lengths starting with the
zero run, then 5 data bits per character with a continuation bit, chars offset by 48, counts after the second stored as
differences to the count two positions earlier).  Its parity is UNPINNED (no pycocotools here, no vectors in the reference);
tests check the round trip with `rle_decode` and a hand-computed example."""
import numpy as np
import torch


def _fmt(v, nd):
    return round(float(v), nd)


def write_results(filename, results):
    """MOT-challenge txt (mot_evaluator.py:50-60): results = [(frame_id, tlwhs, track_ids, scores), ...]; rows with id < 0 skipped."""
    with open(filename, "w") as f:
        for frame_id, tlwhs, track_ids, scores in results:
            for (x1, y1, w, h), tid, s in zip(tlwhs, track_ids, scores):
                if tid < 0:
                    continue
                f.write(f"{frame_id},{tid},{_fmt(x1, 1)},{_fmt(y1, 1)},{_fmt(w, 1)},{_fmt(h, 1)},{_fmt(s, 2)},-1,-1,-1\n")


def write_results_no_score(filename, results):
    """mot_evaluator.py:63-72: results = [(frame_id, tlwhs, track_ids), ...]."""
    with open(filename, "w") as f:
        for frame_id, tlwhs, track_ids in results:
            for (x1, y1, w, h), tid in zip(tlwhs, track_ids):
                if tid < 0:
                    continue
                f.write(f"{frame_id},{tid},{_fmt(x1, 1)},{_fmt(y1, 1)},{_fmt(w, 1)},{_fmt(h, 1)},-1,-1,-1,-1\n")


def write_results_mots(filename, results):
    """MOTS txt (mot_evaluator.py:37-47): results = [(frame_id, track_ids, cat_id, H, W, rles), ...]; ids are offset by 2000."""
    with open(filename, "w") as f:
        for frame_id, track_ids, cat_id, H, W, rles in results:
            for tid, rle in zip(track_ids, rles):
                if tid < 0:
                    continue
                f.write(f"{frame_id} {2000 + tid} {cat_id} {H} {W} {rle}\n")


def overlap_free(masks):
    """mot_evaluator.py:858-866: instance n keeps only the pixels no earlier instance (ascending track id) claimed.
    masks: bool [N,H,W] tensor -> bool [N,H,W] (one cumulative OR instead of the reference's Python loop)."""
    if masks.size(0) == 0:
        return masks
    m = masks.bool()
    claimed_before = (torch.cumsum(m.to(torch.int32), 0) - m.to(torch.int32)) > 0
    return m & ~claimed_before


def rle_encode(mask):
    """COCO compressed RLE string of a binary mask [H,W] (what pycocotools.mask.encode(np.asfortranarray(m))["counts"] holds)."""
    flat = np.asarray(mask, dtype=bool).reshape(-1, order="F")
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(bounds).tolist()
    if flat.size and flat[0]:
        counts = [0] + counts  # the first run counts zeros
    out = []
    for i, c in enumerate(counts):
        x = int(c) - (int(counts[i - 2]) if i > 2 else 0)
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def rle_decode(s, H, W):
    """Inverse of rle_encode: compressed string -> bool [H,W]."""
    counts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    flat = np.zeros(H * W, dtype=bool)
    pos, val = 0, False
    for c in counts:
        flat[pos:pos + c] = val
        pos += c
        val = not val
    return flat.reshape(H, W, order="F")


def mots_frame_result(frame_id, boxes, ids, masks, img_h, img_w, min_box_area=100, cat_id=2):
    """Host half of one MOTS frame after association (mot_evaluator.py:846-897): boxes [n,5] (x1,y1,x2,y2,score) and ids [n]
    as returned by QuasiDenseEmbedTracker.match (valid ids only, any order), masks bool [n,H,W] in the same order.
    Sorts by ascending id, makes the masks overlap free in that order, drops boxes with area <= min_box_area, encodes the
    survivors and returns the tuple write_results_mots() consumes: (frame_id, ids + 1, cat_id, img_h, img_w, rles)."""
    ids = torch.as_tensor(ids).long()
    boxes = torch.as_tensor(boxes, dtype=torch.float32)
    order = ids.sort()[1]
    ids, boxes, masks = ids[order], boxes[order], masks[order]
    free = overlap_free(masks).cpu().numpy() if masks.size(0) else None
    out_ids, rles = [], []
    for i in range(boxes.size(0)):
        tid = int(ids[i])
        if tid < 0:
            continue
        x1, y1, x2, y2 = boxes[i, :4].tolist()
        if (x2 - x1) * (y2 - y1) > min_box_area:
            rles.append(rle_encode(free[i]))
            out_ids.append(tid + 1)  # 1-based ids for the MOTS files
    return frame_id, out_ids, cat_id, img_h, img_w, rles
