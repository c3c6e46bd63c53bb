"""Dump the reference models' state_dict key->shape manifests (run in the build container only).
They pin unicorn_b200.weights.param_shapes() to the reference's weight interface (SURVEY.md §5 checkpoint)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402

out_dir = os.path.dirname(os.path.abspath(__file__))
for name in ("unicorn_track_tiny", "unicorn_track_large", "unicorn_track_large_mot_challenge",
             "unicorn_track_tiny_mask", "unicorn_track_large_mask"):
    _, m = ref_import.get_model(name)
    sd = m.state_dict()
    man = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(out_dir, f"manifest_{name}.json"), "w") as f:
        json.dump(man, f, indent=0, sort_keys=False)
    print(name, len(man), sum(v.numel() for v in sd.values()) / 1e6, "M")
