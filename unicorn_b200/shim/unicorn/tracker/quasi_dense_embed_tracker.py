"""unicorn.tracker.quasi_dense_embed_tracker (reference: unicorn/tracker/quasi_dense_embed_tracker.py:9-212)."""
from unicorn_b200.tracker.quasi_dense import QuasiDenseEmbedTracker  # noqa: F401
