"""unicorn.utils — the inference helpers the per-frame drivers import (reference: unicorn/utils/__init__.py, boxes.py)."""
from .boxes import postprocess, postprocess_inst, xyxy2xywh  # noqa: F401


def fuse_model(model):
    """unicorn/utils/model_utils.py fuse_model: Conv+BN fusion — the tracking models carry GroupNorm and the engine already fuses
    what can be fused; a no-op kept for driver compatibility (tools/track_omni.py --fuse)."""
    return model


def get_model_info(model, tsize):
    return f"UnicornB200Model({getattr(model, 'cfg_name', '?')}) test size {tuple(tsize)}"
