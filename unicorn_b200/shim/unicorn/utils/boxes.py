"""unicorn.utils.boxes — postprocess / postprocess_inst on the GPU (reference: unicorn/utils/boxes.py:33-77, :80-152)."""
from unicorn_b200.compat.model import postprocess, postprocess_inst  # noqa: F401


def xyxy2xywh(bboxes):
    """unicorn/utils/boxes.py xyxy2xywh (in place, like the reference)."""
    bboxes[:, 2] = bboxes[:, 2] - bboxes[:, 0]
    bboxes[:, 3] = bboxes[:, 3] - bboxes[:, 1]
    return bboxes
