import os
import sys


def install():
    """Make `import MultiScaleDeformableAttention` resolve to the unicorn_b200 operator."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
