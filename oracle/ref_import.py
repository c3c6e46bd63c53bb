"""Import the UNMODIFIED reference (MasterBin-IIAU/Unicorn at /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Usable only in the build container (where /root/reference
exists); nothing on the product path, in `-m gpu` tests, smoke() or bench.py imports this.
It is used by tests/golden/make_golden.py to (a) pin oracle/unicorn_oracle.py against the
reference's own modules and (b) generate the committed golden fixtures.

Recipe follows SURVEY.md Appendix A: stub the absent third-party imports, swap the CUDA-only
MSDeformAttnFunction for the reference's own pure-PyTorch `ms_deform_attn_core_pytorch`
(unicorn/models/ops/functions/ms_deform_attn_func.py:41-61), and redirect hard-coded
device="cuda" (deformable_transformer.py:71, unicorn.py:136-138) to the CPU.
"""
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"


class _AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for kk, vv in list(self.items()):
            if isinstance(vv, dict) and not isinstance(vv, _AttrDict):
                self[kk] = _AttrDict(vv)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        self[k] = v


class _CfgNode(_AttrDict):
    def clone(self):
        import copy
        return copy.deepcopy(self)

    def defrost(self):
        pass

    def freeze(self):
        pass

    def merge_from_file(self, *a, **k):
        pass

    def merge_from_list(self, *a, **k):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()

    def forward(self, x):
        return x


_installed = False


def install():
    """Register stubs and make `import unicorn` resolve to /root/reference/unicorn."""
    global _installed
    if _installed:
        return
    _installed = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _mod("thop", profile=lambda *a, **k: (0, 0))
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=_DropPath, trunc_normal_=nn.init.trunc_normal_,
         to_2tuple=lambda x: (x, x) if not isinstance(x, tuple) else x)
    _mod("easydict", EasyDict=_AttrDict)
    _mod("yacs")
    _mod("yacs.config", CfgNode=_CfgNode)
    _mod("lap")
    _mod("cython_bbox", bbox_overlaps=None)
    _mod("MultiScaleDeformableAttention")
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocotools.mask", "motmetrics"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _mod(name, COCO=object, COCOeval=object)
    # device="cuda" -> cpu
    for fn_name in ("zeros", "full", "tensor", "ones", "arange", "linspace", "empty"):
        orig = getattr(torch, fn_name)

        def wrapped(*a, __orig=orig, **k):
            if str(k.get("device", "")) .startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, fn_name, wrapped)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import unicorn  # noqa: F401  (the reference package)
    from unicorn.models.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
    import unicorn.models.ops.modules.ms_deform_attn as mda

    class _Fn:
        @staticmethod
        def apply(value, shapes, level_start, loc, w, step):
            return ms_deform_attn_core_pytorch(value, shapes, loc, w)
    mda.MSDeformAttnFunction = _Fn


def get_model(exp_name, seed=0):
    """exp_name e.g. 'unicorn_track_tiny'; returns reference model in eval mode on CPU."""
    install()
    from unicorn.exp import get_exp
    exp = get_exp(f"{REF_ROOT}/exps/default/{exp_name}.py", None)
    torch.manual_seed(seed)
    model = exp.get_model(load_pretrain=False).eval()
    return exp, model
