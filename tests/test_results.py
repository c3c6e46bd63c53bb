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
