from .quasi_dense import QuasiDenseEmbedTracker  # noqa: F401
