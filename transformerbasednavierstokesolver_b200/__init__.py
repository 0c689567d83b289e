"""B200-native (sm_100a) Physics-Attention path for Transolver Navier-Stokes solvers.

Drop-in for `model/Physics_Attention.py`, `model/Transolver_*`, `model_dict.py` of
OnurBasci/TransformerBasedNavierStokeSolver: same constructors, `forward` signatures, `state_dict` keys
and registry names; the compute runs in hand-written CUDA kernels behind the C ABI in include/tbns.h.
"""
from . import _lib  # noqa: F401
from .config import get_default_precision, set_default_precision  # noqa: F401

__all__ = ["set_default_precision", "get_default_precision"]
