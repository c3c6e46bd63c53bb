"""The weight interface (names, shapes, order) equals the reference state_dict manifests dumped from the reference itself."""
import json
import os

import pytest
import torch

from unicorn_b200.weights import CONFIGS, make_state_dict, param_shapes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_param_shapes_match_reference_manifest(name):
    man = json.load(open(os.path.join(GOLD, f"manifest_{name}.json")))
    mine = [(k, list(v)) for k, v in param_shapes(name).items()]
    assert mine == list(man.items())


def test_param_counts():
    n = lambda name: sum(int(torch.tensor(s).prod()) for s in param_shapes(name).values())  # noqa: E731
    assert n("unicorn_track_tiny") == 59449305      # 59.4 M
    assert n("unicorn_track_large") == 261223737    # 261.2 M
    assert n("unicorn_track_large_mask") == 266247762


def test_seeded_weights_are_deterministic_and_perturbed():
    a, b = make_state_dict("unicorn_track_tiny", 0), make_state_dict("unicorn_track_tiny", 0)
    c = make_state_dict("unicorn_track_tiny", 1)
    k = "transformer.encoder.layers.0.self_attn.sampling_offsets.weight"
    assert all(torch.equal(a[n], b[n]) for n in a)
    assert not torch.equal(a[k], c[k])
    assert a[k].abs().max() > 0  # zero-init in the reference (ms_deform_attn.py:63): perturbed so MSDA depends on data
    assert a["head.obj_preds_sot.0.bias"].mean() > -6
