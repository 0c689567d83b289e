"""ctypes binding of libtbns.so (the C ABI declared in include/tbns.h).

The library is built in-tree (`make -C transformerbasednavierstokesolver_b200/csrc`, or
`__graft_entry__.build()`).  There is NO fallback: if the shared object is missing or a call fails the
caller gets an exception — the product path never silently runs anything else.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtbns.so")

TBNS_PREC_FP32 = 0
TBNS_PREC_BF16 = 1
TBNS_PREC_FP32_EXACT = 2

_fp = C.c_void_p  # device pointers travel as integers
_ll = C.c_longlong
_i = C.c_int


class GemmDesc(C.Structure):
    """mirror of `tbns_gemm_desc` (include/tbns.h) — field order and types must match exactly."""
    _fields_ = [
        ("M", _i), ("N", _i), ("K", _i), ("batch", _i),
        ("sA", _ll), ("sB", _ll), ("sC", _ll), ("sR", _ll), ("sAux", _ll),
        ("A", _fp), ("lda", _ll), ("a_kind", _i),
        ("B", _fp), ("ldb", _ll), ("b_kind", _i),
        ("C", _fp), ("ldc", _ll),
        ("conv_mode", _i), ("Hg", _i), ("Wg", _i), ("Cin", _i), ("flip", _i),
        ("bias", _fp),
        ("residual", _fp), ("ldr", _ll),
        ("act", _i),
        ("aux_out", _fp), ("aux_in", _fp), ("ldaux", _ll),
        ("precision", _i),
        ("split_k", _i), ("ws", _fp),
        ("scatter", _i), ("I", _i), ("taps", _i), ("Cx", _fp), ("Cfx", _fp),
    ]


class TcDesc(C.Structure):
    """mirror of `tbns_tc_desc`"""
    _fields_ = [
        ("A16", _fp), ("Bimg", _i), ("Hg", _i), ("Wg", _i), ("Cin", _i), ("taps", _i), ("flip", _i),
        ("W16", _fp), ("N", _i), ("w_batched", _i),
        ("bias", _fp), ("act", _i),
        ("aux_out", _fp), ("aux_in", _fp), ("ldaux", _ll),
        ("residual", _fp), ("ldr", _ll),
        ("C", _fp), ("ldc", _ll),
        ("C16", _fp), ("ldc16", _ll),
        ("round_tf32", _i), ("aux_bf16", _i),
        ("ln_gamma", _fp), ("ln_beta", _fp), ("ln_out16", _fp), ("ln_mean", _fp), ("ln_rstd", _fp), ("ln_eps", C.c_float),
    ]


class TcWgradDesc(C.Structure):
    """mirror of `tbns_tc_wgrad_desc`"""
    _fields_ = [
        ("A16", _fp), ("Ma", _i),
        ("B16", _fp), ("Nb", _i),
        ("Bimg", _i), ("Hg", _i), ("Wg", _i), ("taps", _i), ("batched", _i),
        ("split_k", _i), ("ws", _fp),
        ("C", _fp), ("ldc", _ll), ("sC", _ll),
        ("scatter", _i), ("I", _i), ("Cx", _fp), ("Cfx", _fp),
    ]


class TbnsError(RuntimeError):
    pass


_SIGS = {
    "tbns_last_error": (C.c_char_p, []),
    "tbns_version": (_i, []),
    "tbns_device_ok": (_i, []),
    "tbns_sm_count": (_i, []),
    "tbns_gemm": (_i, [C.POINTER(GemmDesc), _fp]),
    "tbns_gemm_tc_supported": (_i, [_i, _i, _i]),
    "tbns_gemm_tc": (_i, [C.POINTER(TcDesc), _fp]),
    "tbns_gemm_tc_wgrad_supported": (_i, [_i, _i, _i]),
    "tbns_gemm_tc_wgrad": (_i, [C.POINTER(TcWgradDesc), _fp]),
    "tbns_cast_bf16": (_i, [_fp, _fp, _ll, _fp]),
    "tbns_layernorm_fwd": (_i, [_fp] * 7 + [_i, _i, C.c_float, _fp]),
    "tbns_layernorm_bwd_ws_floats": (C.c_size_t, [_i]),
    "tbns_layernorm_bwd": (_i, [_fp] * 12 + [_i, _i, _fp]),
    "tbns_layernorm_bwd_supported16": (_i, [_i]),
    "tbns_layernorm_bwd16": (_i, [_fp] * 10 + [_i, _i, _fp]),
    "tbns_layernorm_bwd_ctas": (_i, [_i]),
    "tbns_ln_linear1_supported": (_i, [_i]),
    "tbns_ln_linear1_fwd": (_i, [_fp] * 8 + [_i, _i, C.c_float, _fp]),
    "tbns_ln_linear1_fwd_strided": (_i, [_fp] * 6 + [_ll] + [_fp] * 2 + [_i, _i, C.c_float, _fp]),
    "tbns_ln_linear1_bwd": (_i, [_fp] * 11 + [_i, _i, _fp]),
    "tbns_pack_inputs": (_i, [_fp, _i, _fp, _ll, _i, _fp, _ll, _i, _fp, _i, _ll, _i, _fp]),
    "tbns_pack_proj_weights": (_i, [_fp] * 7 + [_i, _i, _i, _fp]),
    "tbns_pack_proj_weights16": (_i, [_fp] * 9 + [_i, _i, _i, _fp]),
    "tbns_cast_bf16_pair": (_i, [_fp] * 3 + [_i, _i, _i, _fp]),
    "tbns_slice_groups": (_i, [_i, _i, _i]),
    "tbns_pa_slice_fwd": (_i, [_fp] * 7 + [_i] * 6 + [_fp]),
    "tbns_pa_slice_tc_supported": (_i, [_i, _i]),
    "tbns_pa_slice_fwd_tc": (_i, [_fp] * 6 + [_i] * 6 + [_fp]),
    "tbns_pa_slice_bwd_tc": (_i, [_fp] * 10 + [_i] * 6 + [_fp]),
    "tbns_pa_proj_bias_grad": (_i, [_fp] * 6 + [_i] * 5 + [_fp]),
    "tbns_pa_token_attn_fwd": (_i, [_fp, _i] + [_fp] * 15 + [_i] * 5 + [_fp]),
    "tbns_pa_token_attn_bwd": (_i, [_fp] * 16 + [_i] * 5 + [_fp]),
    "tbns_pa_slice_bwd": (_i, [_fp] * 12 + [_i] * 6 + [_fp]),
    "tbns_pa_dtau_finish": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _fp]),
    "tbns_reduce_rows": (_i, [_fp, _fp, _i, _ll, _fp]),
    "tbns_adamw_flat": (_i, [_fp] * 5 + [_ll, _fp]),
    "tbns_colsum_ws_floats": (C.c_size_t, [_ll]),
    "tbns_colsum": (_i, [_fp, _ll, _fp, _fp, _i, _i, _fp]),
    "tbns_colsum_bf16": (_i, [_fp, _ll, _fp, _fp, _i, _i, _fp]),
}

EXPORTS = tuple(_SIGS.keys())

_lib = None


def load():
    """dlopen libtbns.so and type its entry points.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tbns_last_error().decode("utf-8", "replace")
        raise TbnsError(f"{what} failed (rc={rc}): {msg}")
