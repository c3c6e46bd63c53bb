"""unicorn_b200 — B200-native (sm_100a) implementation of Unicorn's per-frame inference hot path."""
__version__ = "0.1.0"
