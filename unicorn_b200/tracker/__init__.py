from .quasi_dense import QuasiDenseEmbedTracker  # noqa: F401
from .byte_tracker import BYTETracker  # noqa: F401
