"""The C-ABI library loads without a GPU, exports every symbol include/unicorn_b200.h declares, and the product path
fails loudly (no CPU / PyTorch fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "unicorn_b200.h")).read()
    return sorted(set(re.findall(r"UC_API\s+[\w\s\*]+?\b(uc_\w+)\s*\(", txt)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("uc_conv2d", "uc_msda_forward_f32", "uc_corr_propagate", "uc_postprocess", "uc_dwconv7", "uc_layernorm"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from unicorn_b200 import _lib
    lib = _lib.lib()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/unicorn_b200.h but not exported"
    assert lib.uc_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from unicorn_b200 import _lib
    from unicorn_b200.engine import UnicornEngine
    from unicorn_b200.weights import make_state_dict
    lib = _lib.lib()
    assert lib.uc_check_device() != 0
    with pytest.raises(_lib.UnicornB200Error):
        UnicornEngine(make_state_dict("unicorn_track_tiny", 0), "unicorn_track_tiny", device="cpu")
    d = _lib.UcConv2d()
    assert lib.uc_conv2d(ctypes.byref(d), None) != 0  # argument validation, no launch
