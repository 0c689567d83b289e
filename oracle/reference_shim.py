"""Import the *unmodified* reference modules: from /root/reference when it is mounted (build container), else from
`oracle/_ref/` - the same files byte-compiled by `oracle/build_ref.py` (git-ignored, travels to the GPU box).

TEST / BASELINE INFRASTRUCTURE ONLY.  Callers must check `available()` first.  Two shims are needed:
  * `timm.models.layers.trunc_normal_` (timm is not installed) -> torch.nn.init.trunc_normal_
  * `get_grid` hard-codes `.cuda()` (model/Transolver_Structured_Mesh_2D.py:189,195): on a CPU-only
    host `torch.Tensor.cuda` is patched to identity while the model is constructed.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_SRC_ROOT = os.environ.get("TBNS_REFERENCE_ROOT", "/root/reference")
_PYC_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _root():
    if os.path.isfile(os.path.join(_SRC_ROOT, "model", "Physics_Attention.py")):
        return _SRC_ROOT
    if os.path.isfile(os.path.join(_PYC_ROOT, "model", "Physics_Attention.refbin")):
        return _PYC_ROOT
    return None


REF_ROOT = _root()


def available() -> bool:
    return REF_ROOT is not None


def kind() -> str:
    """'source' (/root/reference mounted), 'compiled' (oracle/_ref) or 'absent'"""
    return "absent" if REF_ROOT is None else ("source" if REF_ROOT == _SRC_ROOT else "compiled")


def _install_shims():
    sys.dont_write_bytecode = True  # the mount is read-only
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers
    if REF_ROOT is None:
        raise ImportError("reference not available: neither /root/reference nor oracle/_ref (run oracle/build_ref.py)")
    if REF_ROOT == _PYC_ROOT:
        if not any(isinstance(f, _CompiledRefFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _CompiledRefFinder(_PYC_ROOT))
    elif REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


class _CompiledRefFinder:
    """meta-path finder for the byte-compiled reference: `model`, `utils` (namespace packages in the reference), their
    modules and `model_dict` resolve to oracle/_ref/**/<name>.refbin (the bytes of a .pyc, loaded sourceless)"""
    PACKAGES = ("model", "utils")

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery as M
        import importlib.util as U
        parts = fullname.split(".")
        if fullname in self.PACKAGES and os.path.isdir(os.path.join(self.root, fullname)):
            spec = M.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [os.path.join(self.root, fullname)]
            return spec
        if parts[0] in self.PACKAGES or fullname == "model_dict":
            f = os.path.join(self.root, *parts) + ".refbin"
            if os.path.isfile(f):
                return U.spec_from_file_location(fullname, f, loader=M.SourcelessFileLoader(fullname, f))
        return None


@contextlib.contextmanager
def cpu_cuda_identity(force: bool = False):
    """Make `.cuda()` a no-op when there is no GPU (reference hard-codes it in get_grid); force=True does so even when a GPU
    is visible (CPU baseline on the GPU box: the model must stay on the host)."""
    if torch.cuda.is_available() and not force:
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def physics_attention():
    _install_shims()
    import importlib
    return importlib.import_module("model.Physics_Attention")


def transolver_2d():
    _install_shims()
    import importlib
    return importlib.import_module("model.Transolver_Structured_Mesh_2D")


def transolver_irregular():
    _install_shims()
    import importlib
    return importlib.import_module("model.Transolver_Irregular_Mesh")


def sol_2d():
    _install_shims()
    import importlib
    return importlib.import_module("model.SOL_Transolver_Structured_Mesh_2D")


def testloss():
    _install_shims()
    import importlib
    return importlib.import_module("utils.testloss")
