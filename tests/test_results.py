"""MOT / MOTS result formats (unicorn_b200/results.py) against the text the reference's own writer functions produce
(tests/golden/results_txt.json, tests/golden/make_golden_results.py), the overlap-free mask rule against the reference's
loop restated literally, and the COCO RLE (parity unpinned: no pycocotools here) through its round trip."""
import json
import os

import numpy as np
import torch

from unicorn_b200 import results as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "results_txt.json")))


def test_writers_match_reference_text(tmp_path):
    for key, fn, arg in (("txt", R.write_results, G["results"]), ("txt_no_score", R.write_results_no_score, G["results_no_score"]),
                         ("txt_mots", R.write_results_mots, G["results_mots"])):
        p = tmp_path / key
        fn(str(p), [tuple(r) for r in arg])
        assert p.read_text() == G[key], key


def test_overlap_free_equals_reference_loop():
    g = torch.Generator().manual_seed(0)
    masks = torch.rand(6, 40, 56, generator=g) > 0.6
    new = masks.clone()
    prev = masks[0].clone()
    for n in range(1, masks.size(0)):  # mot_evaluator.py:861-865
        new[n] = torch.logical_and(torch.logical_not(prev), masks[n])
        prev = torch.logical_or(prev, masks[n])
    assert torch.equal(R.overlap_free(masks), new)
    assert R.overlap_free(masks[:0]).shape[0] == 0
    assert (R.overlap_free(masks).sum(0) <= 1).all()  # every pixel belongs to at most one instance


def test_rle_hand_example_and_round_trip():
    assert R.rle_encode(np.array([[0, 1], [1, 1]])) == "13"          # column-major runs: 1 zero, 3 ones
    assert R.rle_encode(np.zeros((10, 10))) == "T3"                    # one run of 100: 100 = 4 + 3*32 -> 'T' (continued), '3'
    assert R.rle_encode(np.ones((1, 3))) == "03"                       # starts with an empty zero run
    rng = np.random.default_rng(0)
    for shape, p in (((37, 53), 0.5), ((720, 1280), 0.02), ((64, 64), 0.97), ((5, 7), 0.0), ((5, 7), 1.0)):
        m = rng.random(shape) < p
        if p == 0.02:  # blobs: long runs with positive and negative deltas
            m = np.zeros(shape, bool)
            for _ in range(12):
                y, x, h, w = rng.integers(0, 600), rng.integers(0, 1100), rng.integers(5, 120), rng.integers(5, 180)
                m[y:y + h, x:x + w] = True
        s = R.rle_encode(m)
        assert all(48 <= ord(c) < 112 for c in s)
        assert np.array_equal(R.rle_decode(s, *shape), m)


def test_mots_frame_result_equals_reference_steps():
    """mots_frame_result against the reference's per-frame steps written out literally (mot_evaluator.py:850-897: sort by id,
    overlap-free loop, min_box_area filter, RLE of the Fortran-ordered mask, 1-based ids)."""
    g = torch.Generator().manual_seed(3)
    n, H, W = 7, 48, 64
    ids = torch.tensor([5, 2, 9, 0, 7, 3, 1])
    boxes = torch.rand(n, 5, generator=g) * torch.tensor([30.0, 20.0, 40.0, 40.0, 1.0])
    boxes[:, 2:4] += boxes[:, :2] + 15.0          # every box at least 15 x 15
    boxes[3, 2:4] = boxes[3, :2] + 2.0            # area 4: filtered out by min_box_area
    masks = torch.rand(n, H, W, generator=g) > 0.55
    frame = R.mots_frame_result(11, boxes, ids, masks, 720, 1280, min_box_area=100)
    # literal restatement
    _, inds = ids.sort(descending=False)
    s_ids, s_boxes, s_masks = ids[inds], boxes[inds], masks[inds]
    new = s_masks.clone()
    prev = s_masks[0].clone()
    for k in range(1, n):
        new[k] = torch.logical_and(torch.logical_not(prev), s_masks[k])
        prev = torch.logical_or(prev, s_masks[k])
    exp_ids, exp_rle = [], []
    for i in range(n):
        x1, y1, x2, y2, _ = s_boxes[i].tolist()
        if (x2 - x1) * (y2 - y1) > 100:
            exp_rle.append(R.rle_encode(np.asfortranarray(new[i].numpy())))
            exp_ids.append(int(s_ids[i]) + 1)
    assert frame == (11, exp_ids, 2, 720, 1280, exp_rle)
    assert 1 not in frame[1] and len(frame[1]) == n - 1     # id 0 (+1) was the tiny box
    for rle, i in zip(frame[5], [j for j in range(n) if j != 0]):
        assert np.array_equal(R.rle_decode(rle, H, W), new[i].numpy())
