"""One high-priority CUDA stream per device for the association kernels (IoU, bi-softmax): when the MOT driver has
already enqueued the next frame on the main stream, the tracker's tiny launches must not queue behind it."""
import torch

_STREAMS = {}


def assoc_stream(device):
    d = torch.device(device)
    key = d.index if d.index is not None else torch.cuda.current_device()
    s = _STREAMS.get(key)
    if s is None:
        s = _STREAMS[key] = torch.cuda.Stream(device=key, priority=-1)
    return s
