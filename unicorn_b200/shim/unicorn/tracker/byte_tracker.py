"""unicorn.tracker.byte_tracker (reference: unicorn/tracker/byte_tracker.py:13-144 STrack, :147-296 BYTETracker)."""
from unicorn_b200.tracker.byte_tracker import BYTETracker, STrack  # noqa: F401
