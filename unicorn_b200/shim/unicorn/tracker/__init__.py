"""unicorn.tracker — association (reference: unicorn/tracker/byte_tracker.py, quasi_dense_embed_tracker.py)."""
