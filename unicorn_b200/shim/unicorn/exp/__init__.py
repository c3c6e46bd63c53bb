"""unicorn.exp — get_exp / Exp for the tracking configs of the per-frame path (reference: unicorn/exp/build.py:35-50,
unicorn/exp/unicorn_track.py:30-122, unicorn_track_mask.py:31-46, exps/default/unicorn_track_*.py).

get_exp(exp_file, exp_name) keeps the reference's signature: the config is identified by the file's base name (the file itself is
not executed — the reference's Exp classes build PyTorch training objects that do not exist here)."""
import os

from unicorn_b200.compat.model import UnicornB200Model
from unicorn_b200.weights import CONFIGS


class Exp:
    """Attributes read by the inference drivers (unicorn_sot.py:18-25, unicorn_vos.py:19-31, tools/track_omni.py:150-201)."""

    def __init__(self, exp_name):
        if exp_name not in CONFIGS:
            raise KeyError(f"unicorn_b200 has no config {exp_name!r} (known: {sorted(CONFIGS)})")
        cfg = CONFIGS[exp_name]
        self.exp_name = exp_name
        self.num_classes = cfg["num_classes"]       # unicorn_track.py:36 (8) / *_mot_challenge.py:18 (1)
        self.backbone_name = "convnext_tiny" if "tiny" in exp_name else "convnext_large"
        self.normalize = False                      # unicorn_track.py:76
        self.test_size = (800, 1280)                # unicorn_track.py:104
        self.input_size = (800, 1280)
        self.test_conf = 0.001                      # unicorn_track.py (YOLOX default)
        self.nmsthre = 0.65
        self.grid_sample = False
        self.output_dir = "./Unicorn_outputs"
        self.mask = cfg["mask"]
        if cfg["mask"]:
            self.use_raft = True                    # unicorn_track_mask.py:44
            self.d_rate = 2                         # unicorn_track_mask.py:45
            self.ctrl_loc = "reg"                   # unicorn_track_mask.py:38
        self.model = None

    def get_model(self, load_pretrain=True):
        """exp/unicorn_track.py:115-193.  Returns the B200 model shell; weights come from load_state_dict (there is no
        Unicorn_outputs/<pretrain>/best_ckpt.pth lookup: pass load_pretrain=False like the reference's inference drivers)."""
        if load_pretrain:
            raise RuntimeError("get_model(load_pretrain=True) would read a COCO-pretrained checkpoint for training; the inference "
                               "drivers call get_model(load_pretrain=False) and load_state_dict the tracking checkpoint")
        if self.model is None:
            self.model = UnicornB200Model(None, self.exp_name)
        return self.model

    def merge(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
            if hasattr(self, k):
                src = getattr(self, k)
                if src is not None and not isinstance(v, type(src)):
                    try:
                        v = type(src)(v)
                    except Exception:
                        import ast
                        v = ast.literal_eval(v)
                setattr(self, k, v)


ExpTrack = ExpTrackMask = Exp


def get_exp(exp_file=None, exp_name=None):
    """unicorn/exp/build.py:35-50."""
    assert exp_file is not None or exp_name is not None, "plz provide exp file or exp name."
    name = os.path.basename(exp_file).split(".")[0] if exp_file is not None else exp_name
    return Exp(name)
