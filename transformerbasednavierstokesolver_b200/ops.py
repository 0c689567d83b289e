"""Host-side orchestration of the Physics-Attention path: autograd Functions over the libtbns C ABI.

Every tensor handed to the library is a contiguous fp32 CUDA tensor; PyTorch only owns the memory, the
stream and autograd bookkeeping.  Stage order and math follow the reference op by op:

  forward   model/Physics_Attention.py:88-119 (structured) / :31-57 (irregular)
  backward  SURVEY.md §8 (a-bwd); restated and checked in oracle/physics_attention.py

No CPU path exists: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as ct
import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GemmDesc, TBNS_PREC_BF16, TBNS_PREC_FP32, TBNS_PREC_FP32_EXACT, check

# "fp32": 3xTF32 split products on tcgen05 (fp32-class error, gemm_x3.cu); "fp32_exact": fp32 FMA on the SIMT engine
PRECISIONS = {"fp32": TBNS_PREC_FP32, "bf16": TBNS_PREC_BF16, "fp32_exact": TBNS_PREC_FP32_EXACT}
_SM_COUNT = {}


def _num_sms() -> int:
    """multiProcessorCount of the current device (148 on a B200), queried once per device; the B200 figure when no device is
    visible (the host-side schedulers are unit-tested on CPU-only machines)"""
    if not torch.cuda.is_available():
        return 148
    dev = torch.cuda.current_device()
    n = _SM_COUNT.get(dev)
    if n is None:
        n = _SM_COUNT[dev] = int(torch.cuda.get_device_properties(dev).multi_processor_count)
    return n


def _on_device(fn):
    """run an autograd-Function forward/backward with the tensors' device current (kernels launch on THAT device's current
    stream even when the caller's current device is another one)"""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kw)
                break
        return fn(*args, **kw)
    return wrapped

# bookkeeping for bench.py: number of libtbns kernels launched, and (when PROFILE is a dict) CUDA-event pairs around
# tagged launches, recorded on the launching stream
LAUNCHES = 0
PROFILE = None


def _count(n: int):
    global LAUNCHES
    LAUNCHES += n


# TBNS_NVTX=1: every tagged stage of the block (proj_fprop, slice_fwd, token_attn_fwd, deslice_out, mlp_fc1, ...) is wrapped in
# an NVTX range, so timelines taken with the CUDA profilers name the stages of model/Physics_Attention.py:88-119
_NVTX = os.environ.get("TBNS_NVTX", "0") == "1"


class _Timed:
    def __init__(self, tag):
        self.tag = tag if (PROFILE is not None and tag is not None) else None
        self.nvtx = tag if (_NVTX and tag is not None) else None

    def __enter__(self):
        if self.nvtx is not None:
            torch.cuda.nvtx.range_push(self.nvtx)
        if self.tag is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.tag is not None:
            self.e1.record()
            PROFILE.setdefault(self.tag, []).append((self.e0, self.e1))
        if self.nvtx is not None:
            torch.cuda.nvtx.range_pop()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("transformerbasednavierstokesolver_b200 has no CPU path: tensors must live on a CUDA device")
        if t.dtype != torch.float32:
            raise RuntimeError(f"expected float32 tensor, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("expected a contiguous tensor")


def _split_k(M: int, N: int, K: int, batch: int = 1) -> int:
    """pick a split so that small-output / long-K contractions (wgrad) still fill the 148 SMs."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * batch
    sms = _num_sms()
    if tiles >= sms:
        return 1
    s = (2 * sms + tiles - 1) // tiles
    s = min(s, max(1, K // 256))
    return max(1, min(s, 64))


def gemm(*, M, N, K, A, lda, a_kind, B, ldb, b_kind, C=None, ldc=0, batch=1, sA=0, sB=0, sC=0, sR=0, sAux=0,
         conv_mode=0, Hg=0, Wg=0, Cin=0, flip=0, bias=None, residual=None, ldr=0, act=0, aux_out=None, aux_in=None,
         ldaux=0, precision=TBNS_PREC_FP32, split_k=1, scatter=None, I=0, taps=0, tag=None):
    """thin wrapper over `tbns_gemm` (include/tbns.h)."""
    lib = _lib.load()
    d = GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.sA, d.sB, d.sC, d.sR, d.sAux = sA, sB, sC, sR, sAux
    d.A, d.lda, d.a_kind = _p(A), lda, a_kind
    d.B, d.ldb, d.b_kind = _p(B), ldb, b_kind
    d.C, d.ldc = _p(C), ldc
    d.conv_mode, d.Hg, d.Wg, d.Cin, d.flip = conv_mode, Hg, Wg, Cin, flip
    d.bias = _p(bias)
    d.residual, d.ldr = _p(residual), ldr
    d.act, d.aux_out, d.aux_in, d.ldaux = act, _p(aux_out), _p(aux_in), ldaux
    d.precision = precision
    ws = None
    if split_k > 1:
        ws = torch.empty(split_k * batch * M * N, device=A.device, dtype=torch.float32)
    d.split_k, d.ws = split_k, _p(ws)
    if scatter is not None:
        d.scatter, d.I, d.taps = 1, I, taps
        d.Cx, d.Cfx = _p(scatter[0]), _p(scatter[1])
    with _Timed(tag):
        check(lib.tbns_gemm(ct.byref(d), _stream()), "tbns_gemm")
    _count(2 if split_k > 1 else 1)


def reduce_rows(inp: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    out = torch.empty(cols, device=inp.device, dtype=torch.float32)
    with _Timed("reduce_rows"):
        check(_lib.load().tbns_reduce_rows(_p(inp), _p(out), rows, cols, _stream()), "tbns_reduce_rows")
    _count(1)
    return out


def colsum(inp: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(cols, device=inp.device, dtype=torch.float32)
    ws = torch.empty(lib.tbns_colsum_ws_floats(cols), device=inp.device, dtype=torch.float32)
    with _Timed("colsum"):
        check(lib.tbns_colsum(_p(inp), cols, _p(out), _p(ws), rows, cols, _stream()), "tbns_colsum")
    _count(2)
    return out


def colsum_bf16(inp16: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(cols, device=inp16.device, dtype=torch.float32)
    ws = torch.empty(lib.tbns_colsum_ws_floats(cols), device=inp16.device, dtype=torch.float32)
    with _Timed("colsum"):
        check(lib.tbns_colsum_bf16(_p(inp16), cols, _p(out), _p(ws), rows, cols, _stream()), "tbns_colsum_bf16")
    _count(2)
    return out


def cast_bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy (round to nearest even) through libtbns"""
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    with _Timed("cast_bf16"):
        check(_lib.load().tbns_cast_bf16(_p(t), _p(out), t.numel(), _stream()), "tbns_cast_bf16")
    _count(1)
    return out


# ------------------------------------------------------------------------------------------------
# Derived weight copies (bf16 operands, packed projection weights) are cached per master tensor.  A cache entry is valid
# while (a) the master's `_version` / address are unchanged, (b) the global WEIGHT GENERATION is unchanged and (c) it was
# built in the same capture context.  (b) exists because CUDA-graph replays of the optimizer (and `.data` writes) change
# the masters without touching `_version`: whoever mutates parameters behind autograd's back calls
# `invalidate_weight_caches()` (train.GraphedTrainStep does after every optimizer replay, train.broadcast_parameters too).
# (c) makes every graph capture re-record the cast / pack kernels (so replays refresh the copies from the current masters)
# and keeps eager calls from reusing tensors that live in a graph's private pool.
# ------------------------------------------------------------------------------------------------
_W16_CACHE = {}
_WEIGHT_GEN = 0
_CAP_EPOCH = 0
_WAS_CAPTURING = False


def invalidate_weight_caches():
    """call after parameters were modified in a way autograd's version counters do not see (graph replays, `.data` writes)"""
    global _WEIGHT_GEN
    _WEIGHT_GEN += 1


# Fused / foreach optimizers (torch.optim.AdamW(fused=True), the reference's AdamW on CUDA) update parameters WITHOUT bumping
# `_version`, so every optimizer step of any torch optimizer starts a new weight generation.
from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook  # noqa: E402

_reg_post_hook(lambda *_a, **_k: invalidate_weight_caches())


def begin_capture():
    """call right before a new CUDA-graph capture of code that uses this module: derived weights are rebuilt inside it"""
    global _CAP_EPOCH
    _CAP_EPOCH += 1


def cache_context():
    """(weight generation, capture tag): tag 0 = eager, otherwise the epoch of the capture in progress"""
    global _CAP_EPOCH, _WAS_CAPTURING
    cap = torch.cuda.is_current_stream_capturing()
    if cap and not _WAS_CAPTURING:
        _CAP_EPOCH += 1   # a capture nobody announced with begin_capture()
    _WAS_CAPTURING = cap
    return _WEIGHT_GEN, (_CAP_EPOCH if cap else 0)


def _w16_cached(W: torch.Tensor, key_extra, build) -> torch.Tensor:
    """bf16 operand copies of weights, cached per tensor OBJECT (weak reference: a recycled address or id never matches)
    until the fp32 master changes (see the validity rules above)."""
    key = (id(W),) + key_extra
    ctx = cache_context()
    hit = _W16_CACHE.get(key)
    if hit is not None and hit[0]() is W and hit[1] == W._version and hit[2] == W.data_ptr() and hit[4] == ctx:
        return hit[3]
    with torch.no_grad():
        w16 = build()
    if len(_W16_CACHE) > 4096:
        _W16_CACHE.clear()
    _W16_CACHE[key] = (weakref.ref(W), W._version, W.data_ptr(), w16, ctx)
    return w16


def weight_bf16_pair(W: torch.Tensor, Kp: Optional[int] = None):
    """(bf16 copy [R, Kp], transposed bf16 copy [Kp, R]) of a Linear weight [R, K], K zero-padded to Kp (tensor-core
    contractions take K in units of 64): the K-major operands of y = x W^T and dx = dy W, produced by ONE kernel launch and
    cached together."""
    R_, K = W.shape
    Kp = K if Kp is None else Kp

    def build():
        Wc = W.detach().contiguous()
        w16 = torch.empty(R_, Kp, device=W.device, dtype=torch.bfloat16)
        wt16 = torch.empty(Kp, R_, device=W.device, dtype=torch.bfloat16)
        check(_lib.load().tbns_cast_bf16_pair(_p(Wc), _p(w16), _p(wt16), R_, K, Kp, _stream()), "tbns_cast_bf16_pair")
        _count(1)
        return w16, wt16
    return _w16_cached(W, ("pair", Kp), build)


def weight_bf16(W: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """bf16 (optionally transposed) copy of a weight matrix as a K-major tensor-core operand."""
    return weight_bf16_pair(W)[1 if transpose else 0]


def weight_bf16_padded(W: torch.Tensor, Kp: int, transpose: bool = False) -> torch.Tensor:
    """like weight_bf16 for a weight [R, K] whose K is zero-padded to Kp: [R, Kp] or, transposed, [Kp, R]."""
    return weight_bf16_pair(W, Kp)[1 if transpose else 0]


def tc_supported(Cin: int, N: int, taps: int) -> bool:
    return bool(_lib.load().tbns_gemm_tc_supported(Cin, N, taps))


def ln_fusable(N: int) -> bool:
    """the producer GEMM can run the consumer's LayerNorm in its epilogue when one output tile holds whole rows"""
    return N in (128, 256)


def gemm_tc(A16, W16, C, bias, Bimg, Hg, Wg, Cin, N, taps=1, flip=0, tag=None, *, C16=None, w_batched=0, act=0, aux_out=None,
            aux_in=None, residual=None, round_tf32=0, aux_bf16=0, ln=None):
    """tcgen05 implicit GEMM, K-major operands (include/tbns.h: tbns_gemm_tc).  C fp32 and/or C16 bf16 outputs [.., N].
    ln = (gamma, beta, eps): also run LayerNorm over the output rows in the epilogue -> returns (y16, mean, rstd)."""
    d = _lib.TcDesc()
    ln_out = None
    if ln is not None:
        rows = Bimg * Hg * Wg
        y16 = torch.empty(rows, N, device=C.device, dtype=torch.bfloat16)
        mean = torch.empty(rows, device=C.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=C.device, dtype=torch.float32)
        g_, b_ = ln[0].detach().contiguous(), ln[1].detach().contiguous()
        d.ln_gamma, d.ln_beta, d.ln_out16, d.ln_mean, d.ln_rstd, d.ln_eps = _p(g_), _p(b_), _p(y16), _p(mean), _p(rstd), float(ln[2])
        ln_out = (y16, mean, rstd)
    d.A16, d.Bimg, d.Hg, d.Wg, d.Cin, d.taps, d.flip = _p(A16), Bimg, Hg, Wg, Cin, taps, flip
    d.W16, d.N, d.w_batched = _p(W16), N, w_batched
    d.bias, d.act = _p(bias), act
    d.aux_out, d.aux_in, d.ldaux = _p(aux_out), _p(aux_in), N
    d.residual, d.ldr = _p(residual), N
    d.C, d.ldc = _p(C), N
    d.C16, d.ldc16 = _p(C16), N
    d.round_tf32 = round_tf32
    d.aux_bf16 = aux_bf16
    with _Timed(tag):
        check(_lib.load().tbns_gemm_tc(ct.byref(d), _stream()), "tbns_gemm_tc")
    _count(1)
    return ln_out


def _wgrad_split(tiles: int, kblocks: int) -> int:
    """split-K factor of the token-contraction kernel: one CTA per SM (197 KB of shared memory each), so pick the smallest
    split whose CTA count fills whole waves of 148 SMs (>= 95 %), keeping >= 4 k-blocks per CTA."""
    best, best_eff = 1, 0.0
    sms = _num_sms()
    for s in range(1, max(1, min(sms, kblocks // 4)) + 1):
        ctas = tiles * s
        eff = ctas / (((ctas + sms - 1) // sms) * sms)
        if eff >= 0.95:
            return s
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
    return best


def wgrad_supported(Ma: int, Nb: int, taps: int) -> bool:
    return bool(_lib.load().tbns_gemm_tc_wgrad_supported(Ma, Nb, taps))


def gemm_tc_wgrad(A16, B16, Bimg, Hg, Wg, Ma, Nb, taps=1, batched=0, C=None, scatter=None, I=0, tag=None):
    """tcgen05 token-contraction GEMM, MN-major operands (include/tbns.h: tbns_gemm_tc_wgrad):
    D[(tap,a), n] = sum_token A16[shift(token,tap), a] * B16[token, n]; result in C ([batch][taps*Ma][Nb]) or scattered."""
    BN = 256 if Nb % 256 == 0 else (128 if Nb % 128 == 0 else 64)
    batch = Bimg if batched else 1
    tiles = taps * (Ma // 128) * (Nb // BN) * batch
    kblocks = ((Hg * Wg + 63) // 64) * (1 if batched else Bimg)
    split = _wgrad_split(tiles, kblocks)
    ws = torch.empty(split * batch * taps * Ma * Nb, device=A16.device, dtype=torch.float32)
    d = _lib.TcWgradDesc()
    d.A16, d.Ma, d.B16, d.Nb = _p(A16), Ma, _p(B16), Nb
    d.Bimg, d.Hg, d.Wg, d.taps, d.batched = Bimg, Hg, Wg, taps, batched
    d.split_k, d.ws = split, _p(ws)
    d.C, d.ldc, d.sC = _p(C), Nb, taps * Ma * Nb
    if scatter is not None:
        d.scatter, d.I, d.Cx, d.Cfx = 1, I, _p(scatter[0]), _p(scatter[1])
    with _Timed(tag):
        check(_lib.load().tbns_gemm_tc_wgrad(ct.byref(d), _stream()), "tbns_gemm_tc_wgrad")
    _count(2)


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float = 1e-5, want32: bool = True, want16: bool = False):
    """returns (y fp32 | None, y16 bf16 | None, mean, rstd)"""
    _chk(x, gamma, beta)
    C_ = x.shape[-1]
    rows = x.numel() // C_
    y = torch.empty_like(x) if want32 else None
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want16 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    with _Timed("layernorm_fwd"):
        check(_lib.load().tbns_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(y16), _p(mean), _p(rstd), rows, C_, eps, _stream()),
              "tbns_layernorm_fwd")
    _count(1)
    return y, y16, mean, rstd


def layernorm_bwd16(dy16, x, mean, rstd, gamma, dres=None, want16: bool = True):
    """layernorm_bwd for a bf16 incoming gradient -> (dx, dx16 | None, dgamma, dbeta, colsum(dx), keep-alive)"""
    _chk(x, mean, rstd, gamma, dres)
    lib = _lib.load()
    C_ = x.shape[-1]
    rows = x.numel() // C_
    dx = torch.empty_like(x)
    dx16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want16 else None
    ws = torch.empty(lib.tbns_layernorm_bwd_ws_floats(C_), device=x.device, dtype=torch.float32)
    with _Timed("layernorm_bwd"):
        check(lib.tbns_layernorm_bwd16(_p(dy16), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dx16), None, _p(ws),
                                       rows, C_, _stream()), "tbns_layernorm_bwd16")
    _count(1)
    # dgamma | dbeta | colsum(dx) are parameter gradients: their fixed-order reduction over the per-CTA partials runs on the
    # side stream (the callers join before they return; `ws` must stay referenced until then - it is returned for that)
    ctas = lib.tbns_layernorm_bwd_ctas(rows)
    with _OnSide():
        out3 = reduce_rows(ws, ctas, 3 * C_).view(3, C_)
    return dx, dx16, out3[0], out3[1], out3[2], ws


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None, want16: bool = False, want_sum: bool = False):
    """returns (dx, dx16 | None, dgamma, dbeta[, colsum(dx)]); dx = LN'(dy) + dres"""
    _chk(dy, x, mean, rstd, gamma, dres)
    lib = _lib.load()
    C_ = x.shape[-1]
    rows = x.numel() // C_
    dx = torch.empty_like(x)
    dx16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want16 else None
    out3 = torch.empty(3 if want_sum else 2, C_, device=x.device, dtype=torch.float32)   # contiguous: one reduction launch
    dg, db = out3[0], out3[1]
    dsum = out3[2] if want_sum else None
    ws = torch.empty(lib.tbns_layernorm_bwd_ws_floats(C_), device=x.device, dtype=torch.float32)
    with _Timed("layernorm_bwd"):
        check(lib.tbns_layernorm_bwd(_p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dx16), _p(dg), _p(db),
                                     _p(dsum), _p(ws), rows, C_, _stream()), "tbns_layernorm_bwd")
    _count(2 if want_sum else 3)
    if want_sum:
        return dx, dx16, dg, db, dsum
    return dx, dx16, dg, db


# ------------------------------------------------------------------------------------------------
# Weight-gradient work (token-contraction GEMMs, their split-K reductions, bias column sums, the small deterministic
# reductions) is off the critical path of backward: it is issued on a side stream, forked after its inputs are ready and
# joined before the autograd Function returns, so it fills the SMs that the latency-bound kernels of the data-gradient
# chain leave idle.  Under CUDA-graph capture the fork/join become parallel branches of the graph.
# Discipline: tensors read by side-stream kernels stay referenced until _join_side(); tensors allocated while the side
# stream is current belong to its allocator pool.  TBNS_SIDE_STREAM=0 runs everything in order on one stream.
# ------------------------------------------------------------------------------------------------
_SIDE_STREAMS = {}
_USE_SIDE = os.environ.get("TBNS_SIDE_STREAM", "1") != "0"


def _side():
    dev = torch.cuda.current_device()
    s = _SIDE_STREAMS.get(dev)
    if s is None:
        s = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


class _OnSide:
    """launches inside the block go to the side stream, ordered after everything issued so far on the current stream"""

    def __enter__(self):
        self.ctx = None
        if _USE_SIDE:
            side = _side()
            side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _join_side():
    if _USE_SIDE:
        torch.cuda.current_stream().wait_stream(_side())


_PREFETCH_WEIGHTS = os.environ.get("TBNS_SIDE_PREFETCH", "1") != "0"


# Gradient destinations.  train.FlatGradients.begin() arms every parameter with (view of the flat gradient buffer, token of
# this backward pass).  A forward stage notes the destinations of its large weights in its ctx; the matching backward lets the
# wgrad kernels write straight into the view and returns an alias of it, so the end-of-backward gather has nothing to copy for
# that parameter (cfg 1: 42 of the 45 MB).  A view is claimed at most once per pass - a parameter used by several calls
# (unrolled training) gets the first gradient in place and autograd accumulates the others into it.
_DIRECT_GRADS = os.environ.get("TBNS_DIRECT_GRADS", "1") != "0"


def _grad_dst(param):
    return getattr(param, "_tbns_grad_dst", None) if _DIRECT_GRADS else None


def _claim_grad(dst, shape, **kw):
    """-> tensor the weight gradient of this stage should be written to"""
    if dst is not None:
        view, token = dst
        if tuple(view.shape) == tuple(shape) and getattr(view, "_tbns_claimed", None) is not token:
            view._tbns_claimed = token
            return view.detach()      # fresh tensor object on the same memory: autograd adopts it as p.grad without a copy
    return torch.empty(shape, **kw)


def refresh_weights_ahead(owner, blocks, prec: int) -> bool:
    """Issue the derived-weight refresh of ALL blocks (packed bf16 projection operands, bf16 pairs of the MLP weights) on the
    side stream at the start of a forward pass, so that it overlaps the preprocess MLP instead of sitting in front of every
    block's first GEMM (24 small launches per step at cfg 1).  The later per-block lookups hit the cache.  Returns True when
    work was forked: the caller joins (`_join_side()`) before the first block.  `owner` remembers the cache context it
    prefetched for, so calls within the same weight generation (unrolled / rollout loops) cost nothing."""
    if prec != TBNS_PREC_BF16 or not _USE_SIDE or not _PREFETCH_WEIGHTS:
        return False
    ctx = cache_context()
    if getattr(owner, "_tbns_prefetched_ctx", None) == ctx:
        return False
    with _OnSide():
        for b in blocks:
            attn = getattr(b, "Attn", None)
            mlp = getattr(b, "mlp", None)
            if attn is None or mlp is None or not hasattr(attn, "_packed_weights"):
                continue
            attn._packed_weights(prec)
            for lin in (mlp.linear_pre[0], mlp.linear_post):
                if lin.weight.is_cuda and tc_supported(lin.weight.shape[1], lin.weight.shape[0], 1):
                    weight_bf16_pair(lin.weight)
    owner._tbns_prefetched_ctx = ctx
    return True


# Side products of a fused backward stage (a bf16 copy of the gradient it returns, its column sums = the bias gradient of
# the layer that produced the stream) are handed to the stage autograd runs next as an ATTRIBUTE OF THE RETURNED TENSOR
# OBJECT.  PyTorch preserves the Python object of a live tensor, so the next custom Function's backward receives the very
# same object when - and only when - autograd passes the gradient through untouched; whenever it builds a new tensor
# (accumulation of several paths, hooks, casts) the attribute is simply absent and the consumer recomputes what it needs.
# Nothing is keyed by address and nothing outlives the gradient tensor itself.
HANDOFF_HITS = 0      # diagnostics (tests assert that the fused chain really hands over)
HANDOFF_MISSES = 0


def _stash_grad16(t: torch.Tensor, t16: Optional[torch.Tensor] = None, colsum_: Optional[torch.Tensor] = None):
    if t16 is not None or colsum_ is not None:
        t._tbns_side = (t16, colsum_, t._version)
    return t


def _begin_forward():
    """kept as the single hook every forward stage calls (no global state left to reset)"""
    return None


def _take_grad16(t: torch.Tensor):
    """-> (bf16 copy | None, column sums | None) attached by the producer of gradient tensor `t`"""
    global HANDOFF_HITS, HANDOFF_MISSES
    side = getattr(t, "_tbns_side", None)
    if side is not None and side[2] == t._version:
        HANDOFF_HITS += 1
        return side[0], side[1]
    HANDOFF_MISSES += 1
    return None, None


# The LayerNorm a stage starts with may already have been computed by the producer of its input (fused into that GEMM's
# epilogue, gemm_tc(ln=...)): the result travels as an attribute of the producer's output tensor and is used only when it was
# computed with exactly this stage's affine parameters.
LN_FUSED_HITS = 0


def _attach_ln(t: torch.Tensor, ln_out, gamma, beta, eps):
    if ln_out is not None:
        t._tbns_ln = (ln_out, gamma.data_ptr(), gamma._version, beta.data_ptr(), beta._version, float(eps), t._version, _WEIGHT_GEN)
    return t


def _take_ln(t: torch.Tensor, gamma, beta, eps):
    """-> (y16 [rows, C] bf16, mean, rstd) computed by the producer of `t` with these LayerNorm parameters, or None"""
    global LN_FUSED_HITS
    side = getattr(t, "_tbns_ln", None)
    if side is not None and side[1:] == (gamma.data_ptr(), gamma._version, beta.data_ptr(), beta._version, float(eps), t._version,
                                         _WEIGHT_GEN):
        LN_FUSED_HITS += 1
        return side[0]
    return None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, x, gamma, beta, eps):
        _begin_forward()
        x = x.contiguous()
        y, _, mean, rstd = layernorm_fwd(x, gamma.contiguous(), beta.contiguous(), eps)
        ctx.save_for_backward(x, mean, rstd, gamma)
        return y

    @staticmethod
    @_on_device
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        dx, _, dg, db = layernorm_bwd(dy.contiguous(), x, mean, rstd, gamma.contiguous())
        _stash_grad16(dx)
        return dx, dg, db, None


# ------------------------------------------------------------------------------------------------
# projection weight packing (cached by the modules; see model/Physics_Attention.py in this package)
# ------------------------------------------------------------------------------------------------
def pack_proj_weights(Wx, bx, Wfx, bfx, bf16_only: bool = False):
    """nn.Conv2d [I,C,3,3] / nn.Linear [I,C] pair -> (Wf [2I, taps*C], Wd [C, taps*2I], bcat [2I], Wf16, Wd16); one launch.
    bf16_only (the module runs the tensor-core path): only the bf16 operands are produced and returned in the Wf / Wd slots
    as well."""
    _chk(Wx, bx, Wfx, bfx)
    I, C_ = Wx.shape[0], Wx.shape[1]
    taps = 9 if Wx.dim() == 4 else 1
    dev = Wx.device
    bcat = torch.empty(2 * I, device=dev, dtype=torch.float32)
    want_f = tc_supported(C_, 2 * I, taps)
    want_d = tc_supported(2 * I, C_, taps)
    bf16_only = bf16_only and want_f and want_d
    Wf = None if bf16_only else torch.empty(2 * I, taps * C_, device=dev, dtype=torch.float32)
    Wd = None if bf16_only else torch.empty(C_, taps * 2 * I, device=dev, dtype=torch.float32)
    Wf16 = torch.empty(2 * I, taps * C_, device=dev, dtype=torch.bfloat16) if want_f else None
    Wd16 = torch.empty(C_, taps * 2 * I, device=dev, dtype=torch.bfloat16) if want_d else None
    check(_lib.load().tbns_pack_proj_weights16(_p(Wx), _p(bx), _p(Wfx), _p(bfx), _p(Wf), _p(Wd), _p(Wf16), _p(Wd16), _p(bcat), I, C_,
                                               taps, _stream()), "tbns_pack_proj_weights16")
    _count(1)
    if bf16_only:
        return Wf16, Wd16, bcat, Wf16, Wd16
    return Wf, Wd, bcat, Wf16, Wd16


def pa_tc_shapes_ok(C_, I2, HG, Cout, taps) -> bool:
    """shape rules of the tensor-core route of the whole attention module (see _pa_tc_ok)"""
    return (tc_supported(C_, I2, taps) and tc_supported(I2, C_, taps) and wgrad_supported(C_, I2, taps) and tc_supported(HG, Cout, 1)
            and tc_supported(Cout, HG, 1) and wgrad_supported(HG, Cout, 1))


# ------------------------------------------------------------------------------------------------
# Physics-Attention forward / backward
# ------------------------------------------------------------------------------------------------
def _pa_tc_ok(precision, C_, I2, HG, Cout, taps, packed16) -> bool:
    """bf16 mode runs every token contraction of the module on tcgen05 when all of them fit the tensor-core kernels'
    shape rules (channel counts multiples of 64 / 128); otherwise the whole module uses the fp32 SIMT engine with
    bf16-rounded operands (same numerics, slower)."""
    return precision == TBNS_PREC_BF16 and packed16 and pa_tc_shapes_ok(C_, I2, HG, Cout, taps)


def pa_forward(x, temperature, Wf, bcat, Ws, bs, Wq, Wk, Wv, Wo, bo, residual, heads: int,
               grid: Optional[Tuple[int, int]], precision: int, Wf16=None, x16=None, ln_next=None):
    """returns (out, saved tuple). x [B,N,C]; Wf/bcat packed projections; residual [B,N,Cout] or None.
    ln_next = (gamma, beta, eps) of the LayerNorm that consumes `out` (tensor-core route): computed in the epilogue of the
    output GEMM and attached to `out` (see _attach_ln)."""
    lib = _lib.load()
    B, N, C_ = (x if x is not None else x16).shape
    I2 = Wf.shape[0]
    I = I2 // 2
    H = heads
    D = I // H
    G = Ws.shape[0]
    Cout = Wo.shape[0]
    HG = H * G
    dev = (x if x is not None else x16).device
    f32 = dict(device=dev, dtype=torch.float32)
    bf = dict(device=dev, dtype=torch.bfloat16)
    st = _stream()
    structured = grid is not None
    taps = 9 if structured else 1
    Hg, Wg = grid if structured else (1, N)
    tc = _pa_tc_ok(precision, C_, I2, HG, Cout, taps, Wf16 is not None)
    # (1a) projections: XF = [x_mid | fx_mid]   Physics_Attention.py:94-97 / :36-39
    slice_tc = tc and bool(lib.tbns_pa_slice_tc_supported(D, G))
    # the tensor-core slice stage consumes the projections in bf16: the GEMM epilogue writes nothing else (no fp32 XF)
    XF = torch.empty(B * N, I2, **(bf if slice_tc else f32))
    if tc:
        # tensor-core path: bf16 operands through TMA, tcgen05.mma, fp32 accumulate in TMEM
        if x16 is None:
            x16 = cast_bf16(x)
        if slice_tc:
            gemm_tc(x16, Wf16, None, bcat, B, Hg, Wg, C_, I2, taps, 0, tag="proj_fprop", C16=XF)
        else:
            gemm_tc(x16, Wf16, XF, bcat, B, Hg, Wg, C_, I2, taps, 0, tag="proj_fprop")
    elif structured:
        gemm(M=B * N, N=I2, K=9 * C_, A=x, lda=C_, a_kind=0, B=Wf, ldb=9 * C_, b_kind=0, C=XF, ldc=I2, conv_mode=1, Hg=Hg,
             Wg=Wg, Cin=C_, bias=bcat, precision=precision, tag="proj_fprop")
    else:
        gemm(M=B * N, N=I2, K=C_, A=x, lda=C_, a_kind=0, B=Wf, ldb=C_, b_kind=0, C=XF, ldc=I2, bias=bcat, precision=precision,
             tag="proj_fprop")
    # (1b,1c) slice weights + partial slice tokens   :98-101 / :40-42
    groups = lib.tbns_slice_groups(B, N, H)
    w = None if tc else torch.empty(B, N, HG, **f32)
    w16 = torch.empty(B, N, HG, **bf) if tc else None     # tensor-core mode keeps only the bf16 copy
    part = torch.empty(B * H * groups * G * (D + 1), **f32)
    if slice_tc:
        with _Timed("slice_fwd"):
            check(lib.tbns_pa_slice_fwd_tc(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(w16), _p(part), B, N, H, D, G, int(structured), st),
                  "tbns_pa_slice_fwd_tc")
    else:
        with _Timed("slice_fwd"):
            check(lib.tbns_pa_slice_fwd(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(w), _p(w16), _p(part), B, N, H, D, G,
                                        int(structured), st), "tbns_pa_slice_fwd")
    _count(2)  # + token_attn_fwd below
    # (2) token normalisation + attention among slice tokens + fold of to_out   :102-111 / :43-52
    s = torch.empty(B, H, G, **f32)
    Tt, tok, q, k, v, O = (torch.empty(B, H, G, D, **f32) for _ in range(6))
    A = torch.empty(B, H, G, G, **f32)
    P = torch.empty(B, HG, Cout, **f32)
    P16 = torch.empty(B, HG, Cout, **bf) if tc else None    # K-major operand of dw = dOut.P^T
    PT16 = torch.empty(B, Cout, HG, **bf) if tc else None   # K-major operand of out = w.P
    with _Timed("token_attn_fwd"):
      check(lib.tbns_pa_token_attn_fwd(_p(part), groups, _p(Wq), _p(Wk), _p(Wv), _p(Wo), _p(s), _p(Tt), _p(tok), _p(q), _p(k), _p(v),
                                     _p(A), _p(O), _p(P), _p(P16), _p(PT16), B, H, D, G, Cout, st), "tbns_pa_token_attn_fwd")  # noqa: E111
    # (3) deslice (+) to_out (+ bias, + residual)   :116-119 / :55-57
    out = torch.empty(B, N, Cout, **f32)
    if tc:
        fuse = ln_next is not None and ln_fusable(Cout)
        ln_out = gemm_tc(w16, PT16, out, bo, B, 1, N, HG, Cout, w_batched=1, residual=residual, tag="deslice_out",
                         ln=(ln_next if fuse else None))
        if fuse:
            _attach_ln(out, ln_out, *ln_next)
        return out, (XF, w16, s, tok, q, k, v, A, O, P16, x16)
    gemm(M=N, N=Cout, K=HG, A=w, lda=HG, a_kind=0, B=P, ldb=Cout, b_kind=1, C=out, ldc=Cout, batch=B, sA=N * HG,
         sB=HG * Cout, sC=N * Cout, sR=N * Cout, bias=bo, residual=residual, ldr=Cout, precision=precision, tag="deslice_out")
    return out, (XF, w, s, tok, q, k, v, A, O, P, x)


def pa_backward(dout, xshape, temperature, Wd, Wx_shape, Ws, bs, Wq, Wk, Wv, Wo, saved, heads: int,
                grid: Optional[Tuple[int, int]], precision: int, Wd16=None, dout16=None, dbo=None, dx_bf16: bool = False,
                grad_dst=(None, None)):
    """returns dx and the parameter gradients in reference (state_dict) layouts.  `saved` is pa_forward's tuple; its last
    entry is the module input (fp32 in SIMT mode, its bf16 copy in tensor-core mode)."""
    lib = _lib.load()
    XF, w, s, tok, q, k, v, A, O, P, xs = saved
    x = xs
    B, N, C_ = xshape
    I2 = XF.shape[1]
    I = I2 // 2
    H = heads
    D = I // H
    G = Ws.shape[0]
    Cout = Wo.shape[0]
    HG = H * G
    dev = dout.device
    f32 = dict(device=dev, dtype=torch.float32)
    st = _stream()
    structured = grid is not None
    taps = 9 if structured else 1
    Hg, Wg = grid if structured else (1, N)
    groups = lib.tbns_slice_groups(B, N, H)
    tc = w.dtype == torch.bfloat16   # forward ran the tensor-core path and kept bf16 copies

    # (3') deslice (+) to_out backward
    if dbo is None:
        dbo = colsum(dout, B * N, Cout)
    dP = torch.empty(B, HG, Cout, **f32)
    slice_tc = tc and bool(lib.tbns_pa_slice_tc_supported(D, G))
    dw = None if slice_tc else torch.empty(B, N, HG, **f32)
    dw16 = torch.empty(B, N, HG, device=dev, dtype=torch.bfloat16) if slice_tc else None   # the tensor-core slice kernel reads bf16
    if tc:
        if dout16 is None:
            dout16 = cast_bf16(dout)
        with _OnSide():   # dw is not needed before the slice backward: it runs beside dP -> token attention backward
            gemm_tc(dout16, P, dw, None, B, 1, N, Cout, HG, w_batched=1, C16=dw16, tag="deslice_dw")
        gemm_tc_wgrad(w, dout16, B, 1, N, HG, Cout, batched=1, C=dP, tag="deslice_dP")
    else:
        gemm(M=HG, N=Cout, K=N, A=w, lda=HG, a_kind=1, B=dout, ldb=Cout, b_kind=1, C=dP, ldc=Cout, batch=B, sA=N * HG, sB=N * Cout,
             sC=HG * Cout, precision=precision, split_k=_split_k(HG, Cout, N, B), tag="deslice_dP")
        gemm(M=N, N=HG, K=Cout, A=dout, lda=Cout, a_kind=0, B=P, ldb=Cout, b_kind=0, C=dw, ldc=HG, batch=B, sA=N * Cout,
             sB=HG * Cout, sC=N * HG, precision=precision, tag="deslice_dw")
    # (2') token attention backward
    dTt = torch.empty(B, H, G, D, **f32)
    ds = torch.empty(B, H, G, **f32)
    dWqkv_part = torch.empty(B * H, 3 * D * D, **f32)
    dWo_part = torch.empty(B, Cout * I, **f32)
    with _Timed("token_attn_bwd"):
      check(lib.tbns_pa_token_attn_bwd(_p(dP), _p(Wq), _p(Wk), _p(Wv), _p(Wo), _p(s), _p(tok), _p(q), _p(k), _p(v), _p(A), _p(O),
                                     _p(dTt), _p(ds), _p(dWqkv_part), _p(dWo_part), B, H, D, G, Cout, st), "tbns_pa_token_attn_bwd")
    _count(3)  # token_attn_bwd, slice_bwd, dtau_finish
    with _OnSide():
        dWqkv = reduce_rows(dWqkv_part, B * H, 3 * D * D).view(3, D, D)
        dWo = reduce_rows(dWo_part, B, Cout * I).view(Cout, I)
    # (1') slice backward (+ bias gradients of the projections)
    dXF = None if tc else torch.empty(B * N, I2, **f32)
    dXF16 = torch.empty(B * N, I2, device=dev, dtype=torch.bfloat16) if tc else None
    dWs_part = torch.empty(B * H * groups, G * (D + 1), **f32)
    dtau_part = torch.empty(B * H * groups, **f32)
    if tc:
        _join_side()   # dw
    if slice_tc:
        with _Timed("slice_bwd"):
            check(lib.tbns_pa_slice_bwd_tc(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(dw16), _p(dTt), _p(ds), _p(dXF16), _p(dWs_part),
                                           _p(dtau_part), B, N, H, D, G, int(structured), st), "tbns_pa_slice_bwd_tc")
        with _OnSide():
            # projection-bias gradients from token-reduced quantities: db_x = (sum_t dL).Ws, db_fx = (sum_t w).dTt
            dbx, dbfx = torch.empty(I, **f32), torch.empty(I, **f32)
            check(lib.tbns_pa_proj_bias_grad(_p(dWs_part), _p(Ws), _p(s), _p(dTt), _p(dbx), _p(dbfx), B, H, D, G, groups, _stream()),
                  "tbns_pa_proj_bias_grad")
            _count(1)
    else:
        dbcat_part = torch.empty(B * groups, H * 2 * D, **f32)
        with _Timed("slice_bwd"):
            check(lib.tbns_pa_slice_bwd(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(dw), _p(dTt), _p(ds), _p(dXF), _p(dXF16),
                                        _p(dWs_part), _p(dtau_part), _p(dbcat_part), B, N, H, D, G, int(structured), st),
                  "tbns_pa_slice_bwd")
        dbc = reduce_rows(dbcat_part, B * groups, H * 2 * D).view(H, 2, D)
        dbx, dbfx = dbc[:, 0, :].reshape(I), dbc[:, 1, :].reshape(I)
    with _OnSide():
        dWsb = reduce_rows(dWs_part, B * H * groups, G * (D + 1)).view(G, D + 1)
        dWs, dbs = dWsb[:, :D].contiguous(), dWsb[:, D].contiguous()
        dtemp = torch.empty(H, **f32)
        check(lib.tbns_pa_dtau_finish(_p(dtau_part), _p(temperature), _p(dtemp), B, H, groups, int(structured), _stream()),
              "tbns_pa_dtau_finish")
    # (1a') projections: dgrad, wgrad (scattered straight into Conv2d / Linear weight layout)
    # dx_bf16 (tensor-core route, caller feeds a LayerNorm backward): the data gradient is written in bf16 only
    dx_bf16 = dx_bf16 and tc
    dx = torch.empty(B, N, C_, device=dev, dtype=torch.bfloat16) if dx_bf16 else torch.empty(B, N, C_, **f32)
    if tc:
        with _OnSide():
            dWx = _claim_grad(grad_dst[0], Wx_shape, **f32)
            dWfx = _claim_grad(grad_dst[1], Wx_shape, **f32)
            gemm_tc_wgrad(xs, dXF16, B, Hg, Wg, C_, I2, taps=taps, scatter=(dWx, dWfx), I=I, tag="proj_wgrad")
        if dx_bf16:
            gemm_tc(dXF16, Wd16, None, None, B, Hg, Wg, I2, C_, taps, 1, C16=dx, tag="proj_dgrad")
        else:
            gemm_tc(dXF16, Wd16, dx, None, B, Hg, Wg, I2, C_, taps, 1, tag="proj_dgrad")
        keep = (dXF16, xs, dWs_part, dtau_part, dWqkv_part, dWo_part, dTt, s, dw16, dP)
        return dx, dict(temperature=dtemp.view(1, H, 1, 1), Wx=dWx, bx=dbx, Wfx=dWfx, bfx=dbfx, Ws=dWs, bs=dbs,
                        Wq=dWqkv[0], Wk=dWqkv[1], Wv=dWqkv[2], Wo=dWo, bo=dbo, _keep=keep)
    dWx = torch.empty(Wx_shape, **f32)
    dWfx = torch.empty(Wx_shape, **f32)
    if structured:
        gemm(M=B * N, N=C_, K=9 * I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=dx, ldc=C_, conv_mode=1, Hg=Hg, Wg=Wg,
             Cin=I2, flip=1, precision=precision, tag="proj_dgrad")
        gemm(M=9 * C_, N=I2, K=B * N, A=x, lda=C_, a_kind=1, B=dXF, ldb=I2, b_kind=1, conv_mode=2, Hg=Hg, Wg=Wg, Cin=C_,
             precision=precision, split_k=_split_k(9 * C_, I2, B * N), scatter=(dWx, dWfx), I=I, taps=9, tag="proj_wgrad")
    else:
        gemm(M=B * N, N=C_, K=I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=I2, b_kind=0, C=dx, ldc=C_, precision=precision,
             tag="proj_dgrad")
        gemm(M=C_, N=I2, K=B * N, A=x, lda=C_, a_kind=1, B=dXF, ldb=I2, b_kind=1, Cin=C_, precision=precision,
             split_k=_split_k(C_, I2, B * N), scatter=(dWx, dWfx), I=I, taps=1, tag="proj_wgrad")
    keep = (dXF, dWs_part, dtau_part, dWqkv_part, dWo_part, dTt, s, dw, dP)   # read by side-stream launches: alive until the join
    return dx, dict(temperature=dtemp.view(1, H, 1, 1), Wx=dWx, bx=dbx, Wfx=dWfx, bfx=dbfx, Ws=dWs, bs=dbs,
                    Wq=dWqkv[0], Wk=dWqkv[1], Wv=dWqkv[2], Wo=dWo, bo=dbo, _keep=keep)


# ------------------------------------------------------------------------------------------------
# auto-encoder variant (reference Physics_Attention_Structured_Mesh_2D_Auto_Encoder, model/Physics_Attention.py:122-227):
# the same kernels with the pipeline cut after the token stage (encode) / restarted at the deslice (decode).
# The slice weights are a differentiable fp32 tensor here ([B,N,H*G]; the reference caches [B,H,N,G]), so the slice stage
# always runs the exact SIMT kernels; projections and deslice use the tensor-core kernels in bf16 mode when shapes allow.
# The token-attention backward from dO works on G x D tensors per (batch, head) and is plain torch.
# ------------------------------------------------------------------------------------------------
EPS_NORM = 1e-5


def pa_encode_forward(x, temperature, Wf, bcat, Ws, bs, Wq, Wk, Wv, Wo, heads: int, grid, precision: int, Wf16=None):
    """-> (O [B,H,G,D], w [B,N,H*G] fp32, saved)   (encode(), reference :185-212)"""
    lib = _lib.load()
    B, N, C_ = x.shape
    I2 = Wf.shape[0]
    I = I2 // 2
    H = heads
    D = I // H
    G = Ws.shape[0]
    Cout = Wo.shape[0]
    HG = H * G
    f32 = dict(device=x.device, dtype=torch.float32)
    st = _stream()
    structured = grid is not None
    taps = 9 if structured else 1
    Hg, Wg = grid if structured else (1, N)
    tc = (precision == TBNS_PREC_BF16 and Wf16 is not None and tc_supported(C_, I2, taps) and tc_supported(I2, C_, taps)
          and wgrad_supported(C_, I2, taps))
    XF = torch.empty(B * N, I2, **f32)
    x16 = None
    if tc:
        x16 = cast_bf16(x)
        gemm_tc(x16, Wf16, XF, bcat, B, Hg, Wg, C_, I2, taps, 0, tag="proj_fprop")
    elif structured:
        gemm(M=B * N, N=I2, K=9 * C_, A=x, lda=C_, a_kind=0, B=Wf, ldb=9 * C_, b_kind=0, C=XF, ldc=I2, conv_mode=1, Hg=Hg,
             Wg=Wg, Cin=C_, bias=bcat, precision=precision, tag="proj_fprop")
    else:
        gemm(M=B * N, N=I2, K=C_, A=x, lda=C_, a_kind=0, B=Wf, ldb=C_, b_kind=0, C=XF, ldc=I2, bias=bcat, precision=precision,
             tag="proj_fprop")
    O, w, saved6 = _xf_encode_forward(XF, B, N, temperature, Ws, bs, Wq, Wk, Wv, Wo, heads, structured)
    return O, w, (XF, *saved6, x16 if tc else x)


def _xf_encode_forward(XF, B, N, temperature, Ws, bs, Wq, Wk, Wv, Wo, heads: int, clamp: bool):
    """slice weights (exact SIMT kernel, fp32 output) + token attention from the projected features XF [B*N, 2I]
    -> (O [B,H,G,D], w [B,N,H*G], (s, tok, q, k, v, A))"""
    lib = _lib.load()
    I = XF.shape[1] // 2
    H = heads
    D = I // H
    G = Ws.shape[0]
    Cout = Wo.shape[0]
    HG = H * G
    f32 = dict(device=XF.device, dtype=torch.float32)
    st = _stream()
    groups = lib.tbns_slice_groups(B, N, H)
    w = torch.empty(B, N, HG, **f32)
    part = torch.empty(B * H * groups * G * (D + 1), **f32)
    check(lib.tbns_pa_slice_fwd(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(w), None, _p(part), B, N, H, D, G, int(clamp), st),
          "tbns_pa_slice_fwd")
    s = torch.empty(B, H, G, **f32)
    Tt, tok, q, k, v, O = (torch.empty(B, H, G, D, **f32) for _ in range(6))
    A = torch.empty(B, H, G, G, **f32)
    P = torch.empty(B, HG, Cout, **f32)   # by-product of the token kernel (O.Wo_h^T), unused here
    check(lib.tbns_pa_token_attn_fwd(_p(part), groups, _p(Wq), _p(Wk), _p(Wv), _p(Wo), _p(s), _p(Tt), _p(tok), _p(q), _p(k), _p(v),
                                     _p(A), _p(O), _p(P), None, None, B, H, D, G, Cout, st), "tbns_pa_token_attn_fwd")
    _count(2)
    return O, w, (s, tok, q, k, v, A)


def _xf_encode_backward(dO, dw, XF, B, N, temperature, Ws, bs, Wq, Wk, Wv, saved6, heads: int, clamp: bool, want16: bool):
    """backward of _xf_encode_forward from (dO | None, dw | None) -> dXF (fp32, or bf16 when want16) and parameter gradients"""
    lib = _lib.load()
    s, tok, q, k, v, A = saved6
    I2 = XF.shape[1]
    I = I2 // 2
    H = heads
    D = I // H
    G = Ws.shape[0]
    HG = H * G
    dev = XF.device
    f32 = dict(device=dev, dtype=torch.float32)
    st = _stream()
    groups = lib.tbns_slice_groups(B, N, H)
    # token attention backward from dO (SURVEY.md §8 a-bwd; oracle token_attn_bwd without the to_out fold)
    if dO is None:
        dTt = torch.zeros(B, H, G, D, **f32)
        ds = torch.zeros(B, H, G, **f32)
        dWq, dWk, dWv = (torch.zeros(D, D, **f32) for _ in range(3))
    else:
        dO = dO.contiguous()
        dA = dO @ v.transpose(-1, -2)
        dv = A.transpose(-1, -2) @ dO
        dS = A * (dA - (dA * A).sum(-1, keepdim=True))
        scale = D ** -0.5
        dq = dS @ k * scale
        dk = dS.transpose(-1, -2) @ q * scale
        dtok = dq @ Wq + dk @ Wk + dv @ Wv
        dWq = torch.einsum("bhgi,bhgj->ij", dq, tok)
        dWk = torch.einsum("bhgi,bhgj->ij", dk, tok)
        dWv = torch.einsum("bhgi,bhgj->ij", dv, tok)
        inv = 1.0 / (s + EPS_NORM)
        dTt = (dtok * inv[..., None]).contiguous()
        ds = (-(dtok * tok).sum(-1) * inv).contiguous()
    dw = torch.zeros(B, N, HG, **f32) if dw is None else dw.contiguous()
    # slice backward (exact SIMT kernel: fp32 dw in, fp32 or bf16 dXF out)
    dXF = None if want16 else torch.empty(B * N, I2, **f32)
    dXF16 = torch.empty(B * N, I2, device=dev, dtype=torch.bfloat16) if want16 else None
    dWs_part = torch.empty(B * H * groups, G * (D + 1), **f32)
    dtau_part = torch.empty(B * H * groups, **f32)
    dbcat_part = torch.empty(B * groups, H * 2 * D, **f32)
    check(lib.tbns_pa_slice_bwd(_p(XF), _p(Ws), _p(bs), _p(temperature), _p(dw), _p(dTt), _p(ds), _p(dXF), _p(dXF16), _p(dWs_part),
                                _p(dtau_part), _p(dbcat_part), B, N, H, D, G, int(clamp), st), "tbns_pa_slice_bwd")
    dbc = reduce_rows(dbcat_part, B * groups, H * 2 * D).view(H, 2, D)
    dbx, dbfx = dbc[:, 0, :].reshape(I), dbc[:, 1, :].reshape(I)
    dWsb = reduce_rows(dWs_part, B * H * groups, G * (D + 1)).view(G, D + 1)
    dWs, dbs = dWsb[:, :D].contiguous(), dWsb[:, D].contiguous()
    dtemp = torch.empty(H, **f32)
    check(lib.tbns_pa_dtau_finish(_p(dtau_part), _p(temperature), _p(dtemp), B, H, groups, int(clamp), st), "tbns_pa_dtau_finish")
    _count(2)
    return dXF, dXF16, dict(temperature=dtemp.view(1, H, 1, 1), bx=dbx, bfx=dbfx, Ws=dWs, bs=dbs, Wq=dWq, Wk=dWk, Wv=dWv)


def pa_encode_backward(dO, dw, xshape, temperature, Wd, Wx_shape, Ws, bs, Wq, Wk, Wv, saved, heads: int, grid, precision: int,
                       Wd16=None):
    """gradients of encode w.r.t. its input and parameters from (dO [B,H,G,D] | None, dw [B,N,H*G] | None)"""
    XF, s, tok, q, k, v, A, xs = saved
    B, N, C_ = xshape
    I2 = XF.shape[1]
    I = I2 // 2
    f32 = dict(device=XF.device, dtype=torch.float32)
    structured = grid is not None
    taps = 9 if structured else 1
    Hg, Wg = grid if structured else (1, N)
    tc = xs.dtype == torch.bfloat16
    dXF, dXF16, g = _xf_encode_backward(dO, dw, XF, B, N, temperature, Ws, bs, Wq, Wk, Wv, (s, tok, q, k, v, A), heads, structured, tc)
    # projections: dgrad, wgrad
    dx = torch.empty(B, N, C_, **f32)
    dWx = torch.empty(Wx_shape, **f32)
    dWfx = torch.empty(Wx_shape, **f32)
    if tc:
        gemm_tc(dXF16, Wd16, dx, None, B, Hg, Wg, I2, C_, taps, 1, tag="proj_dgrad")
        gemm_tc_wgrad(xs, dXF16, B, Hg, Wg, C_, I2, taps=taps, scatter=(dWx, dWfx), I=I, tag="proj_wgrad")
    elif structured:
        gemm(M=B * N, N=C_, K=9 * I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=dx, ldc=C_, conv_mode=1, Hg=Hg, Wg=Wg,
             Cin=I2, flip=1, precision=precision, tag="proj_dgrad")
        gemm(M=9 * C_, N=I2, K=B * N, A=xs, lda=C_, a_kind=1, B=dXF, ldb=I2, b_kind=1, conv_mode=2, Hg=Hg, Wg=Wg, Cin=C_,
             precision=precision, split_k=_split_k(9 * C_, I2, B * N), scatter=(dWx, dWfx), I=I, taps=9, tag="proj_wgrad")
    else:
        gemm(M=B * N, N=C_, K=I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=I2, b_kind=0, C=dx, ldc=C_, precision=precision,
             tag="proj_dgrad")
        gemm(M=C_, N=I2, K=B * N, A=xs, lda=C_, a_kind=1, B=dXF, ldb=I2, b_kind=1, Cin=C_, precision=precision,
             split_k=_split_k(C_, I2, B * N), scatter=(dWx, dWfx), I=I, taps=1, tag="proj_wgrad")
    g.update(Wx=dWx, Wfx=dWfx)
    return dx, g


class PaEncodeFn(torch.autograd.Function):
    """(code, slice weights) = encode(x): both outputs are differentiable (decode / reconstruct_fx send gradient into the
    cached slice weights, reference :214-227)."""

    @staticmethod
    @_on_device
    def forward(ctx, x, temperature, Wx, bx, Wfx, bfx, Ws, bs, Wq, Wk, Wv, Wo, packed, heads, grid, precision):
        _begin_forward()
        x = x.contiguous()
        Wf, Wd, bcat, Wf16, Wd16 = packed
        _chk(x, temperature, Ws, bs, Wq, Wk, Wv, Wo, bcat)
        temperature_c = temperature.contiguous()
        O, w, saved = pa_encode_forward(x, temperature_c, Wf, bcat, Ws.contiguous(), bs.contiguous(), Wq.contiguous(), Wk.contiguous(),
                                        Wv.contiguous(), Wo.contiguous(), heads, grid, precision, Wf16)
        ctx.save_for_backward(temperature_c, Wd, Ws, bs, Wq, Wk, Wv, *saved)
        ctx.Wd16 = Wd16
        ctx.cfg = (heads, grid, precision, tuple(Wx.shape), tuple(x.shape))
        return O, w

    @staticmethod
    @_on_device
    def backward(ctx, dO, dw):
        heads, grid, precision, wshape, xshape = ctx.cfg
        temperature, Wd, Ws, bs, Wq, Wk, Wv, *saved = ctx.saved_tensors
        dx, g = pa_encode_backward(dO, dw, xshape, temperature, Wd, wshape, Ws.contiguous(), bs.contiguous(), Wq.contiguous(),
                                   Wk.contiguous(), Wv.contiguous(), tuple(saved), heads, grid, precision, ctx.Wd16)
        _stash_grad16(dx)
        return (dx, g["temperature"], g["Wx"], g["bx"], g["Wfx"], g["bfx"], g["Ws"], g["bs"], g["Wq"], g["Wk"], g["Wv"], None,
                None, None, None, None)


class PaDecodeFn(torch.autograd.Function):
    """out = to_out(deslice(code, w)) = w . (code . Wo_h^T) + bo   (decode(), reference :221-227)"""

    @staticmethod
    @_on_device
    def forward(ctx, code, w, Wo, bo, precision):
        _begin_forward()
        code, w, Wo, bo = code.contiguous(), w.contiguous(), Wo.contiguous(), bo.contiguous()
        _chk(code, w, Wo, bo)
        B, H, G, D = code.shape
        N, HG = w.shape[1], w.shape[2]
        Cout = Wo.shape[0]
        P = torch.einsum("bhgd,chd->bhgc", code, Wo.view(Cout, H, D)).reshape(B, HG, Cout).contiguous()
        tc = (precision == TBNS_PREC_BF16 and tc_supported(HG, Cout, 1) and tc_supported(Cout, HG, 1) and wgrad_supported(HG, Cout, 1))
        out = torch.empty(B, N, Cout, device=w.device, dtype=torch.float32)
        if tc:
            w16, P16 = cast_bf16(w), cast_bf16(P)
            PT16 = cast_bf16(P.transpose(1, 2).contiguous())
            gemm_tc(w16, PT16, out, bo, B, 1, N, HG, Cout, w_batched=1, tag="deslice_out")
            ctx.save_for_backward(code, Wo, w16, P16)
        else:
            gemm(M=N, N=Cout, K=HG, A=w, lda=HG, a_kind=0, B=P, ldb=Cout, b_kind=1, C=out, ldc=Cout, batch=B, sA=N * HG,
                 sB=HG * Cout, sC=N * Cout, bias=bo, precision=precision, tag="deslice_out")
            ctx.save_for_backward(code, Wo, w, P)
        ctx.precision = precision
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        code, Wo, w, P = ctx.saved_tensors
        precision = ctx.precision
        dout = dout.contiguous()
        B, H, G, D = code.shape
        N, HG = w.shape[1], w.shape[2]
        Cout = Wo.shape[0]
        f32 = dict(device=dout.device, dtype=torch.float32)
        dbo = colsum(dout, B * N, Cout)
        dP = torch.empty(B, HG, Cout, **f32)
        dw = torch.empty(B, N, HG, **f32)
        if w.dtype == torch.bfloat16:
            dout16 = cast_bf16(dout)
            gemm_tc_wgrad(w, dout16, B, 1, N, HG, Cout, batched=1, C=dP, tag="deslice_dP")
            gemm_tc(dout16, P, dw, None, B, 1, N, Cout, HG, w_batched=1, tag="deslice_dw")
        else:
            gemm(M=HG, N=Cout, K=N, A=w, lda=HG, a_kind=1, B=dout, ldb=Cout, b_kind=1, C=dP, ldc=Cout, batch=B, sA=N * HG, sB=N * Cout,
                 sC=HG * Cout, precision=precision, split_k=_split_k(HG, Cout, N, B), tag="deslice_dP")
            gemm(M=N, N=HG, K=Cout, A=dout, lda=Cout, a_kind=0, B=P, ldb=Cout, b_kind=0, C=dw, ldc=HG, batch=B, sA=N * Cout,
                 sB=HG * Cout, sC=N * HG, precision=precision, tag="deslice_dw")
        dP4 = dP.view(B, H, G, Cout)
        dcode = torch.einsum("bhgc,chd->bhgd", dP4, Wo.view(Cout, H, D))
        dWo = torch.einsum("bhgc,bhgd->chd", dP4, code).reshape(Cout, H * D)
        return dcode, dw, dWo, dbo, None


class SliceLinearFn(torch.autograd.Function):
    """project_slice: nn.Linear(slice_num, slice_num) over the slice index of the cached weights w [B,N,H*G]
    (reconstruct_fx(), reference :215); exact fp32 contraction (K = slice_num)."""

    @staticmethod
    @_on_device
    def forward(ctx, w, Wp, bp):
        _begin_forward()
        w, Wp, bp = w.contiguous(), Wp.contiguous(), bp.contiguous()
        _chk(w, Wp, bp)
        G = Wp.shape[0]
        M = w.numel() // G
        out = torch.empty_like(w)
        gemm(M=M, N=G, K=G, A=w, lda=G, a_kind=0, B=Wp, ldb=G, b_kind=0, C=out, ldc=G, bias=bp, precision=TBNS_PREC_FP32)
        ctx.save_for_backward(w, Wp)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        w, Wp = ctx.saved_tensors
        dout = dout.contiguous()
        G = Wp.shape[0]
        M = w.numel() // G
        f32 = dict(device=w.device, dtype=torch.float32)
        dw = torch.empty_like(w)
        gemm(M=M, N=G, K=G, A=dout, lda=G, a_kind=0, B=Wp, ldb=G, b_kind=1, C=dw, ldc=G, precision=TBNS_PREC_FP32)
        dWp = torch.empty(G, G, **f32)
        gemm(M=G, N=G, K=M, A=dout, lda=G, a_kind=1, B=w, ldb=G, b_kind=1, C=dWp, ldc=G, precision=TBNS_PREC_FP32,
             split_k=_split_k(G, G, M))
        dbp = colsum(dout, M, G)
        return dw, dWp, dbp


# ------------------------------------------------------------------------------------------------
# structured 3D variant (reference Physics_Attention_Structured_Mesh_3D, model/Physics_Attention.py:232-288):
# the Conv3d 3x3x3 projections run as three passes of the 2D implicit-GEMM kernels, the rest is XF -> encode -> decode.
# ------------------------------------------------------------------------------------------------
class XFEncodeFn(torch.autograd.Function):
    """(slice tokens after attention O, slice weights w) from already projected features XF [B,N,2I]"""

    @staticmethod
    @_on_device
    def forward(ctx, XF, temperature, Ws, bs, Wq, Wk, Wv, Wo, heads, clamp):
        _begin_forward()
        B, N, I2 = XF.shape
        XF2 = XF.contiguous().view(B * N, I2)
        temperature_c = temperature.contiguous()
        Ws, bs, Wq, Wk, Wv, Wo = (t.contiguous() for t in (Ws, bs, Wq, Wk, Wv, Wo))
        _chk(XF2, temperature_c, Ws, bs, Wq, Wk, Wv, Wo)
        O, w, saved6 = _xf_encode_forward(XF2, B, N, temperature_c, Ws, bs, Wq, Wk, Wv, Wo, heads, clamp)
        ctx.save_for_backward(XF2, temperature_c, Ws, bs, Wq, Wk, Wv, *saved6)
        ctx.cfg = (B, N, heads, clamp)
        return O, w

    @staticmethod
    @_on_device
    def backward(ctx, dO, dw):
        B, N, heads, clamp = ctx.cfg
        XF2, temperature, Ws, bs, Wq, Wk, Wv, *saved6 = ctx.saved_tensors
        dXF, _, g = _xf_encode_backward(dO, dw, XF2, B, N, temperature, Ws, bs, Wq, Wk, Wv, tuple(saved6), heads, clamp, False)
        return dXF.view(B, N, -1), g["temperature"], g["Ws"], g["bs"], g["Wq"], g["Wk"], g["Wv"], None, None, None


class Conv3dProjFn(torch.autograd.Function):
    """XF = [in_project_x(x) | in_project_fx(x)] for the Conv3d(3x3x3, pad 1) pair of the 3D module (reference :246-247,
    :264-267; tokens n = (h*W + w)*D + d).  The (W, D) planes are the images of the 2D implicit-GEMM kernels (B*H of them);
    the kernel's three H taps are three passes over H-shifted copies of the zero-padded input that accumulate in place
    (bias in the first pass, `residual = XF` afterwards).  packs[kh] = pack_proj_weights of the kh-th (kW, kD) weight plane."""

    @staticmethod
    @_on_device
    def forward(ctx, x, Wx, bx, Wfx, bfx, packs, grid3, precision):
        _begin_forward()
        x = x.contiguous()
        _chk(x)
        B, N, C_ = x.shape
        Hg, Wg, Dg = grid3
        I2 = packs[0][0].shape[0]
        tc = (precision == TBNS_PREC_BF16 and all(pk[3] is not None and pk[4] is not None for pk in packs)
              and tc_supported(C_, I2, 9) and tc_supported(I2, C_, 9) and wgrad_supported(C_, I2, 9))
        xp = torch.zeros(B, Hg + 2, Wg, Dg, C_, device=x.device, dtype=torch.float32)
        xp[:, 1:Hg + 1] = x.view(B, Hg, Wg, Dg, C_)
        XF = torch.empty(B * N, I2, device=x.device, dtype=torch.float32)
        planes = []
        for kh in range(3):
            A = xp[:, kh:kh + Hg].contiguous()
            if tc:
                A = cast_bf16(A)
            planes.append(A)
            Wf, _, bcat, Wf16, _ = packs[kh]
            first = kh == 0
            if tc:
                gemm_tc(A, Wf16, XF, bcat if first else None, B * Hg, Wg, Dg, C_, I2, 9, 0, residual=None if first else XF,
                        tag="proj3d_fprop")
            else:
                gemm(M=B * N, N=I2, K=9 * C_, A=A, lda=C_, a_kind=0, B=Wf, ldb=9 * C_, b_kind=0, C=XF, ldc=I2, conv_mode=1, Hg=Wg,
                     Wg=Dg, Cin=C_, bias=bcat if first else None, residual=None if first else XF, ldr=I2, precision=precision,
                     tag="proj3d_fprop")
        ctx.save_for_backward(*planes)
        ctx.packs = packs
        ctx.cfg = (B, N, C_, grid3, I2, precision, tc, tuple(Wx.shape))
        return XF.view(B, N, I2)

    @staticmethod
    @_on_device
    def backward(ctx, dXF):
        B, N, C_, (Hg, Wg, Dg), I2, precision, tc, wshape = ctx.cfg
        planes = ctx.saved_tensors
        I = I2 // 2
        dev = dXF.device
        f32 = dict(device=dev, dtype=torch.float32)
        dXF = dXF.contiguous().view(B * N, I2)
        dbcat = colsum(dXF, B * N, I2)
        dXF16 = cast_bf16(dXF) if tc else None
        dxp = torch.zeros(B, Hg + 2, Wg, Dg, C_, **f32)
        dWx = torch.empty(wshape, **f32)
        dWfx = torch.empty(wshape, **f32)
        for kh in range(3):
            _, Wd, _, _, Wd16 = ctx.packs[kh]
            dpl = torch.empty(B * N, C_, **f32)
            dWx_k = torch.empty(wshape[0], wshape[1], 3, 3, **f32)
            dWfx_k = torch.empty(wshape[0], wshape[1], 3, 3, **f32)
            if tc:
                gemm_tc(dXF16, Wd16, dpl, None, B * Hg, Wg, Dg, I2, C_, 9, 1, tag="proj3d_dgrad")
                gemm_tc_wgrad(planes[kh], dXF16, B * Hg, Wg, Dg, C_, I2, taps=9, scatter=(dWx_k, dWfx_k), I=I, tag="proj3d_wgrad")
            else:
                gemm(M=B * N, N=C_, K=9 * I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=dpl, ldc=C_, conv_mode=1, Hg=Wg,
                     Wg=Dg, Cin=I2, flip=1, precision=precision, tag="proj3d_dgrad")
                gemm(M=9 * C_, N=I2, K=B * N, A=planes[kh], lda=C_, a_kind=1, B=dXF, ldb=I2, b_kind=1, conv_mode=2, Hg=Wg, Wg=Dg,
                     Cin=C_, precision=precision, split_k=_split_k(9 * C_, I2, B * N), scatter=(dWx_k, dWfx_k), I=I, taps=9,
                     tag="proj3d_wgrad")
            dxp[:, kh:kh + Hg] += dpl.view(B, Hg, Wg, Dg, C_)
            dWx[:, :, kh] = dWx_k
            dWfx[:, :, kh] = dWfx_k
        dx = dxp[:, 1:Hg + 1].reshape(B, N, C_)
        return dx, dWx, dbcat[:I].contiguous(), dWfx, dbcat[I:].contiguous(), None, None, None


class PhysicsAttentionFn(torch.autograd.Function):
    """y = PhysicsAttention(x) (+ residual).  Parameters arrive in reference layout; `packed` holds the packed projection
    weights (fp32 and bf16 copies; non-differentiable, refreshed by the module when the masters change)."""

    @staticmethod
    @_on_device
    def forward(ctx, x, residual, temperature, Wx, bx, Wfx, bfx, Ws, bs, Wq, Wk, Wv, Wo, bo, packed, heads, grid, precision):
        _begin_forward()
        x = x.contiguous()
        if residual is not None:
            residual = residual.contiguous()
        Wf, Wd, bcat, Wf16, Wd16 = packed
        _chk(x, residual, temperature, Ws, bs, Wq, Wk, Wv, Wo, bo, bcat)
        temperature_c = temperature.contiguous()
        out, saved = pa_forward(x, temperature_c, Wf, bcat, Ws.contiguous(), bs.contiguous(), Wq.contiguous(), Wk.contiguous(),
                                Wv.contiguous(), Wo.contiguous(), bo.contiguous(), residual, heads, grid, precision, Wf16)
        ctx.save_for_backward(temperature_c, Wd, Ws, bs, Wq, Wk, Wv, Wo, *saved)
        ctx.Wd16 = Wd16
        ctx.cfg = (heads, grid, precision, tuple(Wx.shape), residual is not None, tuple(x.shape))
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        heads, grid, precision, wshape, has_res, xshape = ctx.cfg
        temperature, Wd, Ws, bs, Wq, Wk, Wv, Wo, *saved = ctx.saved_tensors
        dout = dout.contiguous()
        dx, g = pa_backward(dout, xshape, temperature, Wd, wshape, Ws.contiguous(), bs.contiguous(), Wq.contiguous(), Wk.contiguous(),
                            Wv.contiguous(), Wo.contiguous(), tuple(saved), heads, grid, precision, ctx.Wd16)
        _join_side()
        g.pop("_keep", None)
        _stash_grad16(dx)
        return (dx, dout if has_res else None, g["temperature"], g["Wx"], g["bx"], g["Wfx"], g["bfx"], g["Ws"], g["bs"], g["Wq"],
                g["Wk"], g["Wv"], g["Wo"], g["bo"], None, None, None, None)


class AttnBlockFn(torch.autograd.Function):
    """fx + PhysicsAttention(LayerNorm(fx))  — first half of Transolver_block.forward (Transolver_Structured_Mesh_2D.py:70).
    In tensor-core mode ln_1's output exists only as the bf16 TMA operand; backward fuses the residual branch into the
    LayerNorm backward kernel and hands a bf16 copy of its result to the previous block's MLP backward."""

    @staticmethod
    @_on_device
    def forward(ctx, fx, ln_w, ln_b, eps, temperature, Wx, bx, Wfx, bfx, Ws, bs, Wq, Wk, Wv, Wo, bo, packed, heads, grid, precision,
                nln_w=None, nln_b=None, nln_eps=1e-5):
        _begin_forward()
        fx = fx.contiguous()
        Wf, Wd, bcat, Wf16, Wd16 = packed
        _chk(fx, ln_w, ln_b, temperature, Ws, bs, Wq, Wk, Wv, Wo, bo, bcat)
        B, N, C_ = fx.shape
        structured = grid is not None
        tc = _pa_tc_ok(precision, C_, Wf.shape[0], heads * Ws.shape[0], Wo.shape[0], 9 if structured else 1, Wf16 is not None)
        pre = _take_ln(fx, ln_w, ln_b, eps) if tc else None
        if pre is not None:      # ln_1 was computed by the producer of fx (previous block's fc2 / the preprocess MLP)
            x1, (x1_16, mean, rstd) = None, pre
            x1_16 = x1_16.view(B, N, C_)
        else:
            x1, x1_16, mean, rstd = layernorm_fwd(fx, ln_w.contiguous(), ln_b.contiguous(), eps, want32=not tc, want16=tc)
        temperature_c = temperature.contiguous()
        out, saved = pa_forward(x1, temperature_c, Wf, bcat, Ws.contiguous(), bs.contiguous(), Wq.contiguous(), Wk.contiguous(),
                                Wv.contiguous(), Wo.contiguous(), bo.contiguous(), fx, heads, grid, precision, Wf16, x16=x1_16,
                                ln_next=((nln_w, nln_b, nln_eps) if (tc and nln_w is not None) else None))
        ctx.save_for_backward(fx, ln_w, mean, rstd, temperature_c, Wd, Ws, bs, Wq, Wk, Wv, Wo, *saved)
        ctx.Wd16 = Wd16
        ctx.cfg = (heads, grid, precision, tuple(Wx.shape), tuple(fx.shape))
        ctx.grad_dst = (_grad_dst(Wx), _grad_dst(Wfx))
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        heads, grid, precision, wshape, xshape = ctx.cfg
        fx, ln_w, mean, rstd, temperature, Wd, Ws, bs, Wq, Wk, Wv, Wo, *saved = ctx.saved_tensors
        dout = dout.contiguous()
        dout16, dsum = _take_grad16(dout)
        ln16 = bool(_lib.load().tbns_layernorm_bwd_supported16(fx.shape[-1]))
        dx1, g = pa_backward(dout, xshape, temperature, Wd, wshape, Ws.contiguous(), bs.contiguous(), Wq.contiguous(), Wk.contiguous(),
                             Wv.contiguous(), Wo.contiguous(), tuple(saved), heads, grid, precision, ctx.Wd16, dout16=dout16,
                             dbo=dsum, dx_bf16=ln16, grad_dst=ctx.grad_dst)
        if dx1.dtype == torch.bfloat16:
            dfx, dfx16, dlw, dlb, dfsum, _ln_keep = layernorm_bwd16(dx1, fx, mean, rstd, ln_w, dres=dout)   # noqa: F841 (alive until the join)
        else:
            dfx, dfx16, dlw, dlb, dfsum = layernorm_bwd(dx1, fx, mean, rstd, ln_w, dres=dout, want16=precision == TBNS_PREC_BF16,
                                                        want_sum=True)
        _join_side()                        # weight-gradient work of pa_backward overlapped the LayerNorm backward above
        g.pop("_keep", None)
        _stash_grad16(dfx, dfx16, dfsum)   # column sums of dfx = to-be bias gradient of the previous block's mlp.linear_post
        return (dfx, dlw, dlb, None, g["temperature"], g["Wx"], g["bx"], g["Wfx"], g["bfx"], g["Ws"], g["bs"], g["Wq"], g["Wk"],
                g["Wv"], g["Wo"], g["bo"], None, None, None, None, None, None, None)


# ------------------------------------------------------------------------------------------------
# LayerNorm + MLP (+ residual)   model/Transolver_Structured_Mesh_2D.py:71 with MLP :13-38 (n_layers=0)
# ------------------------------------------------------------------------------------------------
class LnMlpFn(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, fx, gamma, beta, W1, b1, W2, b2, eps, precision, nln_w=None, nln_b=None, nln_eps=1e-5):
        _begin_forward()
        fx = fx.contiguous()
        _chk(fx, gamma, beta, W1, b1, W2, b2)
        C_ = fx.shape[-1]
        M = fx.numel() // C_
        R = W1.shape[0]
        Cout = W2.shape[0]
        use_tc = (precision == TBNS_PREC_BF16 and Cout == C_ and tc_supported(C_, R, 1) and tc_supported(R, Cout, 1)
                  and wgrad_supported(Cout, R, 1) and wgrad_supported(R, C_, 1))
        got = _take_ln(fx, gamma, beta, eps) if use_tc else None
        W1_param, W2_param = W1, W2
        gamma, beta, W1, b1, W2, b2 = (t.contiguous() for t in (gamma, beta, W1, b1, W2, b2))
        if got is not None:      # ln_2 was computed in the epilogue of the attention's output GEMM
            x2, (x2_16, mean, rstd) = None, got
        else:
            x2, x2_16, mean, rstd = layernorm_fwd(fx, gamma, beta, eps, want32=not use_tc, want16=use_tc)
        pre = torch.empty(M, R, device=fx.device, dtype=torch.float32)
        if use_tc:
            # tensor-core MLP: bf16 operands; LN output and hidden activation only ever exist in bf16 (+ fp32 pre-activation
            # for GELU')
            hid16 = torch.empty(M, R, device=fx.device, dtype=torch.bfloat16)
            pre = torch.empty(M, R, device=fx.device, dtype=torch.bfloat16)   # GELU'(pre-activation), written by fc1's epilogue (act 3)
            gemm_tc(x2_16, weight_bf16(W1), None, b1, 1, 1, M, C_, R, act=3, aux_out=pre, aux_bf16=1, C16=hid16, tag="mlp_fc1")
            out = torch.empty(*fx.shape[:-1], Cout, device=fx.device, dtype=torch.float32)
            fuse = nln_w is not None and ln_fusable(Cout)
            ln_out = gemm_tc(hid16, weight_bf16(W2), out, b2, 1, 1, M, R, Cout, residual=fx, tag="mlp_fc2",
                             ln=((nln_w, nln_b, nln_eps) if fuse else None))   # + the next block's ln_1
            if fuse:
                _attach_ln(out, ln_out, nln_w, nln_b, nln_eps)
            ctx.save_for_backward(fx, gamma, W1, W2, x2_16, mean, rstd, pre, hid16)
            ctx.precision = precision
            ctx.grad_dst = (_grad_dst(W1_param), _grad_dst(W2_param))
            return out
        hid = torch.empty(M, R, device=fx.device, dtype=torch.float32)
        gemm(M=M, N=R, K=C_, A=x2, lda=C_, a_kind=0, B=W1, ldb=C_, b_kind=0, C=hid, ldc=R, bias=b1, act=1, aux_out=pre, ldaux=R,
             precision=precision)
        out = torch.empty(*fx.shape[:-1], Cout, device=fx.device, dtype=torch.float32)
        gemm(M=M, N=Cout, K=R, A=hid, lda=R, a_kind=0, B=W2, ldb=R, b_kind=0, C=out, ldc=Cout, bias=b2, residual=fx, ldr=C_,
             precision=precision)
        ctx.save_for_backward(fx, gamma, W1, W2, x2, mean, rstd, pre, hid)
        ctx.precision = precision
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        fx, gamma, W1, W2, x2, mean, rstd, pre, hid = ctx.saved_tensors
        precision = ctx.precision
        dout = dout.contiguous()
        C_ = fx.shape[-1]
        M = fx.numel() // C_
        R = W1.shape[0]
        Cout = W2.shape[0]
        f32 = dict(device=fx.device, dtype=torch.float32)
        dout16, db2 = _take_grad16(dout)
        if db2 is None:
            db2 = colsum(dout, M, Cout)
        gdst = getattr(ctx, "grad_dst", (None, None))
        if hid.dtype == torch.bfloat16:   # tensor-core mode
            if dout16 is None:
                dout16 = cast_bf16(dout)
            W2t16, W1t16 = weight_bf16(W2, transpose=True), weight_bf16(W1, transpose=True)
            with _OnSide():
                dW2 = _claim_grad(gdst[1], (Cout, R), **f32)
                gemm_tc_wgrad(dout16, hid, 1, 1, M, Cout, R, C=dW2, tag="mlp_dW2")
            dpre16 = torch.empty(M, R, device=fx.device, dtype=torch.bfloat16)
            gemm_tc(dout16, W2t16, None, None, 1, 1, M, Cout, R, act=4, aux_in=pre, aux_bf16=1, C16=dpre16, tag="mlp_dpre")
            with _OnSide():
                db1 = colsum_bf16(dpre16, M, R)
                dW1 = _claim_grad(gdst[0], (R, C_), **f32)
                gemm_tc_wgrad(dpre16, x2, 1, 1, M, R, C_, C=dW1, tag="mlp_dW1")
            if _lib.load().tbns_layernorm_bwd_supported16(C_):
                dx2 = torch.empty(M, C_, device=fx.device, dtype=torch.bfloat16)   # feeds the LayerNorm backward only: bf16
                gemm_tc(dpre16, W1t16, None, None, 1, 1, M, R, C_, C16=dx2, tag="mlp_dx2")
                dfx, dfx16, dg, db, dfsum, _ln_keep = layernorm_bwd16(dx2, fx, mean, rstd, gamma, dres=dout)   # noqa: F841 (alive until the join)
            else:
                dx2 = torch.empty(M, C_, **f32)
                gemm_tc(dpre16, W1t16, dx2, None, 1, 1, M, R, C_, tag="mlp_dx2")
                dfx, dfx16, dg, db, dfsum = layernorm_bwd(dx2, fx, mean, rstd, gamma, dres=dout, want16=True, want_sum=True)
            _join_side()
            dfx = dfx.view_as(fx)
            _stash_grad16(dfx, dfx16, dfsum)   # the attention stage's backward consumes dfx next: bf16 copy + to_out bias gradient
            return dfx, dg, db, dW1, db1, dW2, db2, None, None, None, None, None
        dW2 = torch.empty(Cout, R, **f32)
        gemm(M=Cout, N=R, K=M, A=dout, lda=Cout, a_kind=1, B=hid, ldb=R, b_kind=1, C=dW2, ldc=R, precision=precision,
             split_k=_split_k(Cout, R, M))
        dpre = torch.empty(M, R, **f32)
        gemm(M=M, N=R, K=Cout, A=dout, lda=Cout, a_kind=0, B=W2, ldb=R, b_kind=1, C=dpre, ldc=R, act=2, aux_in=pre, ldaux=R,
             precision=precision)
        db1 = colsum(dpre, M, R)
        dW1 = torch.empty(R, C_, **f32)
        gemm(M=R, N=C_, K=M, A=dpre, lda=R, a_kind=1, B=x2, ldb=C_, b_kind=1, C=dW1, ldc=C_, precision=precision,
             split_k=_split_k(R, C_, M))
        dx2 = torch.empty(M, C_, **f32)
        gemm(M=M, N=C_, K=R, A=dpre, lda=R, a_kind=0, B=W1, ldb=C_, b_kind=1, C=dx2, ldc=C_, precision=precision)
        dfx, _, dg, db = layernorm_bwd(dx2, fx, mean, rstd, gamma, dres=dout)  # + residual branch
        dfx = dfx.view_as(fx)
        _stash_grad16(dfx)
        return dfx, dg, db, dW1, db1, dW2, db2, None, None, None, None, None


def mlp_tc_ok(K: int, R: int, Cout: int) -> bool:
    Kp = -(-K // 64) * 64
    return (tc_supported(Kp, R, 1) and tc_supported(R, Cout, 1) and tc_supported(R, Kp, 1) and tc_supported(Cout, R, 1)
            and wgrad_supported(Cout, R, 1) and wgrad_supported(R, Kp, 1))


def _mlp_tc_forward(in16, W1, b1, W2, b2, Kp, out_shape, ln_next=None):
    """Linear(Kp -> R) + GELU + Linear(R -> Cout) on the tensor cores from the packed bf16 input [M, Kp]"""
    M = in16.shape[0]
    R, Cout = W1.shape[0], W2.shape[0]
    dev = in16.device
    b1, b2 = b1.contiguous(), b2.contiguous()
    pre16 = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
    hid16 = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
    gemm_tc(in16, weight_bf16_padded(W1, Kp), None, b1, 1, 1, M, Kp, R, act=3, aux_out=pre16, aux_bf16=1, C16=hid16, tag="pre_fc1")
    out = torch.empty(*out_shape, Cout, device=dev, dtype=torch.float32)
    fuse = ln_next is not None and ln_next[0] is not None and ln_fusable(Cout)
    ln_out = gemm_tc(hid16, weight_bf16(W2), out, b2, 1, 1, M, R, Cout, tag="pre_fc2", ln=(ln_next if fuse else None))
    if fuse:
        _attach_ln(out, ln_out, *ln_next)   # the first block's ln_1
    return out, pre16, hid16


def _mlp_tc_backward(dout, in16, W1, W2, pre16, hid16, K, need_dx):
    """-> (dxp [M, Kp] fp32 | None, dW1 [R, K], db1, dW2, db2)"""
    M, Kp = in16.shape
    R, Cout = W1.shape[0], W2.shape[0]
    f32 = dict(device=dout.device, dtype=torch.float32)
    dout = dout.contiguous()
    dout16, db2 = _take_grad16(dout)
    if dout16 is None:
        dout16 = cast_bf16(dout)
    if db2 is None:
        db2 = colsum(dout, M, Cout)
    dW2 = torch.empty(Cout, R, **f32)
    gemm_tc_wgrad(dout16, hid16, 1, 1, M, Cout, R, C=dW2, tag="pre_dW2")
    dpre16 = torch.empty(M, R, device=dout.device, dtype=torch.bfloat16)
    gemm_tc(dout16, weight_bf16(W2, transpose=True), None, None, 1, 1, M, Cout, R, act=4, aux_in=pre16, aux_bf16=1, C16=dpre16,
            tag="pre_dpre")
    db1 = colsum_bf16(dpre16, M, R)
    dW1p = torch.empty(R, Kp, **f32)
    gemm_tc_wgrad(dpre16, in16, 1, 1, M, R, Kp, C=dW1p, tag="pre_dW1")
    dxp = None
    if need_dx:
        dxp = torch.empty(M, Kp, **f32)
        gemm_tc(dpre16, weight_bf16_padded(W1, Kp, transpose=True), dxp, None, 1, 1, M, R, Kp, tag="pre_dx")
    return dxp, dW1p[:, :K], db1, dW2, db2


class MlpFn(torch.autograd.Function):
    """`preprocess`: Linear(K -> R) + GELU + Linear(R -> Cout) on the tensor cores (bf16 mode)
    model/Transolver_Structured_Mesh_2D.py:13-38,165-166,206-207.  K (= fun_dim + 64 or + space_dim) is zero-padded to a
    multiple of 64 so the input rows are whole TMA boxes."""

    @staticmethod
    @_on_device
    def forward(ctx, inp, W1, b1, W2, b2):
        _begin_forward()
        K = inp.shape[-1]
        M = inp.numel() // K
        Kp = -(-K // 64) * 64
        in16 = torch.empty(M, Kp, device=inp.device, dtype=torch.bfloat16)
        src = inp.reshape(M, K).float().contiguous()
        check(_lib.load().tbns_pack_inputs(None, 0, None, 0, 0, _p(src), K, K, _p(in16), Kp, M, M, _stream()), "tbns_pack_inputs")
        _count(1)
        out, pre16, hid16 = _mlp_tc_forward(in16, W1, b1, W2, b2, Kp, inp.shape[:-1])
        ctx.save_for_backward(in16, W1, W2, pre16, hid16)
        ctx.K = K
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        in16, W1, W2, pre16, hid16 = ctx.saved_tensors
        K = ctx.K
        dxp, dW1, db1, dW2, db2 = _mlp_tc_backward(dout, in16, W1, W2, pre16, hid16, K, ctx.needs_input_grad[0])
        dinp = dxp[:, :K].reshape(*dout.shape[:-1], K) if dxp is not None else None
        return dinp, dW1, db1, dW2, db2


def _row_strided(t: torch.Tensor):
    """[.., F] fp32 tensor whose rows are uniformly strided (last dim contiguous) -> (tensor to keep alive, row stride)"""
    t = t if t.dtype == torch.float32 else t.float()
    F = t.shape[-1]
    ok = t.stride(-1) == 1 or F == 1
    ld = t.stride(-2) if t.dim() >= 2 else F
    for d in range(t.dim() - 2, 0, -1):     # leading dims must collapse onto the row stride
        ok = ok and t.stride(d - 1) == t.stride(d) * t.shape[d]
    if not ok or ld < F or (t.data_ptr() % 4):
        t = t.contiguous()
        ld = F
    return t, ld


class PackedMlpFn(torch.autograd.Function):
    """`preprocess(cat(x, fx))` (model/Transolver_Structured_Mesh_2D.py:203-207) with the concatenation fused into the load:
    the bf16 operand of the first Linear is written in ONE pass from a bf16 feature table [N, R] broadcast over the batch
    (the unified-position features; the reference materialises pos.repeat(B) in fp32 every call), an optional fp32 source
    (raw coordinates) and the input fields `fx` - which may be a strided window of a frame history, so the window shift of
    the rollout loops (SOL_Transolver_Structured_Mesh_2D.py:47-52) costs nothing."""

    @staticmethod
    @_on_device
    def forward(ctx, tab16, src1, fx, W1, b1, W2, b2, nln_w=None, nln_b=None, nln_eps=1e-5):
        _begin_forward()
        B, N = fx.shape[0], fx.shape[1]
        M = B * N
        R = tab16.shape[-1] if tab16 is not None else 0
        F1 = src1.shape[-1] if src1 is not None else 0
        F2 = fx.shape[-1]
        K = R + F1 + F2
        Kp = -(-K // 64) * 64
        s1, ld1 = _row_strided(src1) if src1 is not None else (None, 0)
        s2, ld2 = _row_strided(fx)
        in16 = torch.empty(M, Kp, device=fx.device, dtype=torch.bfloat16)
        check(_lib.load().tbns_pack_inputs(_p(tab16), R, _p(s1), ld1, F1, _p(s2), ld2, F2, _p(in16), Kp, M, N, _stream()),
              "tbns_pack_inputs")
        _count(1)
        out, pre16, hid16 = _mlp_tc_forward(in16, W1, b1, W2, b2, Kp, (B, N), (nln_w, nln_b, nln_eps))
        ctx.save_for_backward(in16, W1, W2, pre16, hid16)
        ctx.dims = (R, F1, F2, B, N)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        in16, W1, W2, pre16, hid16 = ctx.saved_tensors
        R, F1, F2, B, N = ctx.dims
        K = R + F1 + F2
        need1, need2 = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dxp, dW1, db1, dW2, db2 = _mlp_tc_backward(dout, in16, W1, W2, pre16, hid16, K, need1 or need2)
        d1 = dxp[:, R:R + F1].reshape(B, N, F1) if need1 else None
        d2 = dxp[:, R + F1:K].reshape(B, N, F2) if need2 else None
        return None, d1, d2, dW1, db1, dW2, db2, None, None, None


class LnLinearFn(torch.autograd.Function):
    """last layer: mlp2(ln_3(fx))   model/Transolver_Structured_Mesh_2D.py:72-73.
    out_dim == 1 (every reference script) runs as one fused pass per direction (tbns_ln_linear1_fwd/_bwd); other widths go
    through LayerNorm + the generic contraction."""

    @staticmethod
    @_on_device
    def forward(ctx, fx, gamma, beta, W, b, eps, precision):
        _begin_forward()
        fx = fx.contiguous()
        gamma, beta, W, b = (t.contiguous() for t in (gamma, beta, W, b))
        _chk(fx, gamma, beta, W, b)
        lib = _lib.load()
        C_ = fx.shape[-1]
        M = fx.numel() // C_
        Od = W.shape[0]
        ctx.precision = precision
        ctx.fused1 = bool(Od == 1 and lib.tbns_ln_linear1_supported(C_))
        out = torch.empty(*fx.shape[:-1], Od, device=fx.device, dtype=torch.float32)
        if ctx.fused1:
            mean = torch.empty(M, device=fx.device, dtype=torch.float32)
            rstd = torch.empty_like(mean)
            with _Timed("ln_linear1_fwd"):
                check(lib.tbns_ln_linear1_fwd(_p(fx), _p(gamma), _p(beta), _p(W), _p(b), _p(out), _p(mean), _p(rstd), M, C_,
                                              float(eps), _stream()), "tbns_ln_linear1_fwd")
            _count(1)
            ctx.save_for_backward(fx, gamma, beta, W, mean, rstd)
            return out
        x3, _, mean, rstd = layernorm_fwd(fx, gamma, beta, eps)
        gemm(M=M, N=Od, K=C_, A=x3, lda=C_, a_kind=0, B=W, ldb=C_, b_kind=0, C=out, ldc=Od, bias=b, precision=precision)
        ctx.save_for_backward(fx, gamma, W, x3, mean, rstd)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        precision = ctx.precision
        dout = dout.contiguous()
        if ctx.fused1:
            fx, gamma, beta, W, mean, rstd = ctx.saved_tensors
            lib = _lib.load()
            C_ = fx.shape[-1]
            M = fx.numel() // C_
            want16 = precision == TBNS_PREC_BF16
            dfx = torch.empty_like(fx)
            dfx16 = torch.empty(fx.shape, device=fx.device, dtype=torch.bfloat16) if want16 else None
            sums = torch.empty(3, C_, device=fx.device, dtype=torch.float32)
            ws = torch.empty(lib.tbns_layernorm_bwd_ws_floats(C_), device=fx.device, dtype=torch.float32)
            with _Timed("ln_linear1_bwd"):
                check(lib.tbns_ln_linear1_bwd(_p(dout), _p(W), _p(fx), _p(mean), _p(rstd), _p(gamma), None, _p(dfx), _p(dfx16),
                                              _p(sums), _p(ws), M, C_, _stream()), "tbns_ln_linear1_bwd")
            _count(2)
            S, D = sums[0], sums[1, :1]
            w = W[0]
            _stash_grad16(dfx, dfx16, sums[2])
            return dfx, w * S, w * D, (gamma * S + beta * D).view_as(W), D.clone(), None, None
        fx, gamma, W, x3, mean, rstd = ctx.saved_tensors
        C_ = fx.shape[-1]
        M = fx.numel() // C_
        Od = W.shape[0]
        f32 = dict(device=fx.device, dtype=torch.float32)
        db = colsum(dout, M, Od)
        dW = torch.empty(Od, C_, **f32)
        gemm(M=Od, N=C_, K=M, A=dout, lda=Od, a_kind=1, B=x3, ldb=C_, b_kind=1, C=dW, ldc=C_, precision=precision,
             split_k=_split_k(Od, C_, M))
        dx3 = torch.empty(M, C_, **f32)
        gemm(M=M, N=C_, K=Od, A=dout, lda=Od, a_kind=0, B=W, ldb=C_, b_kind=1, C=dx3, ldc=C_, precision=precision)
        dfx, dfx16, dg, dbeta, dfsum = layernorm_bwd(dx3, fx, mean, rstd, gamma, want16=precision == TBNS_PREC_BF16, want_sum=True)
        dfx = dfx.view_as(fx)
        _stash_grad16(dfx, dfx16, dfsum)
        return dfx, dg, dbeta, dW, db, None, None


def ln_linear_into(fx, gamma, beta, W, b, eps, precision, out):
    """inference-only form of LnLinearFn: mlp2(ln_3(fx)) written in place into `out`, a [.., out_dim] view whose rows may be
    strided (one column of a frame history: the rollout needs no concatenation afterwards).  Returns `out`."""
    lib = _lib.load()
    fx = fx.contiguous()
    C_ = fx.shape[-1]
    M = fx.numel() // C_
    Od = W.shape[0]
    if (Od == 1 and lib.tbns_ln_linear1_supported(C_) and out.dtype == torch.float32 and out.is_cuda and out.numel() == M
            and _row_strided(out)[0] is out):
        gamma, beta, W, b = (t.detach().contiguous() for t in (gamma, beta, W, b))
        _chk(fx, gamma, beta, W, b)
        mean = torch.empty(M, device=fx.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(fx.device), _Timed("ln_linear1_fwd"):
            check(lib.tbns_ln_linear1_fwd_strided(_p(fx), _p(gamma), _p(beta), _p(W), _p(b), _p(out), _row_strided(out)[1], _p(mean),
                                                  _p(rstd), M, C_, float(eps), _stream()), "tbns_ln_linear1_fwd_strided")
        _count(1)
        return out
    with torch.no_grad():
        out.copy_(LnLinearFn.apply(fx, gamma, beta, W, b, eps, precision))
    return out
