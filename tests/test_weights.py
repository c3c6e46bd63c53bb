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


def test_load_checkpoint_formats_and_validation(tmp_path):
    """The reference checkpoint layout ({"model": state_dict}, optional DDP `module.` prefix, fp16 tensors) is accepted and
    validated against the parameter table; wrong shapes / missing keys fail loudly."""
    import pytest
    from unicorn_b200.weights import check_state_dict, load_checkpoint
    name = "unicorn_track_tiny"
    sd = make_state_dict(name, 0)
    p = tmp_path / "best_ckpt.pth"
    torch.save({"model": {"module." + k: v.half() for k, v in sd.items()}, "start_epoch": 3}, p)
    got = load_checkpoint(str(p), name)
    assert list(got) == list(sd) and all(v.dtype == torch.float32 for v in got.values())
    assert torch.equal(got["backbone.backbone.stages.0.0.gamma"], sd["backbone.backbone.stages.0.0.gamma"].half().float())
    assert list(load_checkpoint(sd, name)) == list(sd)                     # bare state_dict
    m = make_state_dict("unicorn_track_tiny_mask", 0)                      # mask models: head.mask_head.* buffers are not parameters
    assert len(load_checkpoint(m, "unicorn_track_tiny_mask")) == len(m)
    assert len(load_checkpoint({k: v for k, v in m.items() if "mask_head" not in k}, "unicorn_track_tiny_mask")) == len(m) - 2
    bad = dict(sd)
    k0 = "backbone.backbone.stages.1.0.pwconv1.weight"
    bad[k0] = bad[k0][:-1]
    del bad["head.stems.0.conv.weight"]
    bad["extra.weight"] = torch.zeros(1)
    with pytest.raises(ValueError, match="does not match"):
        load_checkpoint(bad, name)
    missing, unexpected, mismatched = check_state_dict(bad, name, strict=False)
    assert missing == ["head.stems.0.conv.weight"] and unexpected == ["extra.weight"] and mismatched[0][0] == k0


def test_committed_tuning_tables_are_well_formed():
    """unicorn_b200/tuned/*.json: layer-shape keys (12 '|'-separated fields) -> an N tile the C side accepts."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "unicorn_b200", "tuned", "*.json"))
    assert files
    valid = {0, 16, 32, 64, 96, 128, 192, 256, 1128, 1192, 1256}
    for f in files:
        tab = json.load(open(f))
        assert tab, f
        for k, v in tab.items():
            assert len(k.split("|")) == 12 and v in valid, (f, k, v)
