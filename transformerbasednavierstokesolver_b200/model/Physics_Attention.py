"""Drop-in replacements for the reference Physics-Attention modules.

Same constructor arguments, attributes (`heads`, `dim_head`, `H`, `W`, `temperature`, ...), parameter
names / shapes (so `state_dict()` round-trips with reference checkpoints) and `forward(x)` contract as
  Physics_Attention_Irregular_Mesh        reference model/Physics_Attention.py:6-57
  Physics_Attention_Structured_Mesh_2D    reference model/Physics_Attention.py:60-119
  Physics_Attention_Structured_Mesh_2D_Auto_Encoder   reference model/Physics_Attention.py:122-227 (encode / decode /
                                          reconstruct_fx on a cached, differentiable slice-weight tensor)
  Physics_Attention_Structured_Mesh_3D    reference model/Physics_Attention.py:232-288 (Conv3d projections as three passes
                                          of the 2D implicit-GEMM kernels over H-shifted planes)
but `forward` runs the sm_100a kernels of libtbns through `ops.PhysicsAttentionFn` (custom backward).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops


class _PhysicsAttentionBase(nn.Module):
    structured = False

    def _build(self, dim, heads, dim_head, dropout, slice_num, proj_factory):
        if dim_head not in (8, 16, 32, 64) or slice_num not in (4, 8, 16, 32, 64):
            raise NotImplementedError(f"the slice kernels are built for dim_head in {{8,16,32,64}} and slice_num in {{4,8,16,32,64}} "
                                      f"(got dim_head={dim_head}, slice_num={slice_num})")
        inner = dim_head * heads
        self.dim_head = dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.softmax = nn.Softmax(dim=-1)   # kept only so module listings match the reference
        self.dropout = nn.Dropout(dropout)
        self.temperature = nn.Parameter(torch.full((1, heads, 1, 1), 0.5))
        self.in_project_x = proj_factory(dim, inner)
        self.in_project_fx = proj_factory(dim, inner)
        self.in_project_slice = nn.Linear(dim_head, slice_num)
        nn.init.orthogonal_(self.in_project_slice.weight)
        self.to_q = nn.Linear(dim_head, dim_head, bias=False)
        self.to_k = nn.Linear(dim_head, dim_head, bias=False)
        self.to_v = nn.Linear(dim_head, dim_head, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.precision = None  # None -> config.get_default_precision()
        self._pack_key = None
        self._packed = None

    # packed projection weights (fprop / dgrad operand layouts), refreshed when the masters change (ops: cache validity rules)
    def _packed_weights(self, prec=None):
        px, pfx = self.in_project_x, self.in_project_fx
        lin = self.to_out[0]
        taps = 9 if px.weight.dim() == 4 else 1
        # bf16 mode on tensor-core shapes: only the bf16 operands are needed - one launch per refresh
        bf16_only = (prec == ops.TBNS_PREC_BF16 and px.weight.dim() in (2, 4) and
                     ops.pa_tc_shapes_ok(px.weight.shape[1], 2 * px.weight.shape[0], self.heads * self.in_project_slice.weight.shape[0],
                                         lin.weight.shape[0], taps))
        key = (px.weight.data_ptr(), px.weight._version, px.bias._version, pfx.weight.data_ptr(), pfx.weight._version,
               pfx.bias._version, px.weight.device, ops.cache_context(), bf16_only)
        if key != self._pack_key:
            with torch.no_grad(), torch.cuda.device(px.weight.device):
                self._packed = ops.pack_proj_weights(px.weight.detach().contiguous(), px.bias.detach().contiguous(),
                                                     pfx.weight.detach().contiguous(), pfx.bias.detach().contiguous(), bf16_only)
            self._pack_key = key
        return self._packed

    def _grid(self, N):
        return None

    def forward(self, x, residual=None):
        """x [B,N,C] -> [B,N,dim].  `residual` (optional, extension used by Transolver_block) is added in the
        output-projection epilogue."""
        if not x.is_cuda:
            raise RuntimeError("Physics-Attention (B200) has no CPU path: move the module and its input to a CUDA device")
        if self.training and self.dropout.p > 0.0:
            raise NotImplementedError("dropout > 0 in training mode is not supported by the fused kernels "
                                      "(every reference script uses dropout=0.0)")
        B, N, C = x.shape
        grid = self._grid((B, N, C))
        prec = ops.PRECISIONS[self.precision or config.get_default_precision()]
        packed = self._packed_weights(prec)
        lin = self.to_out[0]
        return ops.PhysicsAttentionFn.apply(
            x.float(), residual, self.temperature, self.in_project_x.weight, self.in_project_x.bias, self.in_project_fx.weight,
            self.in_project_fx.bias, self.in_project_slice.weight, self.in_project_slice.bias, self.to_q.weight, self.to_k.weight,
            self.to_v.weight, lin.weight, lin.bias, packed, self.heads, grid, prec)


def _forward_block(self, fx, ln, next_ln=None):
    """fx + self(ln(fx)) as ONE fused autograd stage (used by Transolver_block): LayerNorm, attention, residual.
    next_ln: the nn.LayerNorm that consumes the result (the block's ln_2) - computed in the output GEMM's epilogue."""
    if not fx.is_cuda:
        raise RuntimeError("Physics-Attention (B200) has no CPU path: move the module and its input to a CUDA device")
    if self.training and self.dropout.p > 0.0:
        raise NotImplementedError("dropout > 0 in training mode is not supported by the fused kernels")
    grid = self._grid(tuple(fx.shape))
    prec = ops.PRECISIONS[self.precision or config.get_default_precision()]
    packed = self._packed_weights(prec)
    lin = self.to_out[0]
    return ops.AttnBlockFn.apply(
        fx.float(), ln.weight, ln.bias, ln.eps, self.temperature, self.in_project_x.weight, self.in_project_x.bias,
        self.in_project_fx.weight, self.in_project_fx.bias, self.in_project_slice.weight, self.in_project_slice.bias,
        self.to_q.weight, self.to_k.weight, self.to_v.weight, lin.weight, lin.bias, packed, self.heads, grid, prec,
        *((next_ln.weight, next_ln.bias, next_ln.eps) if next_ln is not None else (None, None, 1e-5)))


_PhysicsAttentionBase.forward_block = _forward_block


class Physics_Attention_Irregular_Mesh(_PhysicsAttentionBase):
    """for irregular meshes in 1D, 2D or 3D space (temperature is NOT clamped, reference :40)."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., slice_num=64):
        super().__init__()
        self._build(dim, heads, dim_head, dropout, slice_num, lambda i, o: nn.Linear(i, o))


class Physics_Attention_Structured_Mesh_2D(_PhysicsAttentionBase):
    """for structured meshes in 2D space: 3x3 conv projections, temperature clamped to [0.1, 5] (reference :99)."""
    structured = True

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., slice_num=64, H=101, W=31, kernel=3):
        super().__init__()
        if kernel != 3:
            raise NotImplementedError("only kernel=3 is supported (the reference never passes another value)")
        self.H = H
        self.W = W
        self._build(dim, heads, dim_head, dropout, slice_num, lambda i, o: nn.Conv2d(i, o, kernel, 1, kernel // 2))

    def _grid(self, shape):
        B, N, C = shape
        if N != self.H * self.W:
            raise RuntimeError(f"shape '[{B}, {self.H}, {self.W}, {C}]' is invalid for input of size {B * N * C}")
        return (self.H, self.W)


class Physics_Attention_Structured_Mesh_2D_Auto_Encoder(Physics_Attention_Structured_Mesh_2D):
    """Auto-encoder variant (reference model/Physics_Attention.py:122-227): `forward` is the 2D module's; `encode` stops
    after the attention among slice tokens and returns them (the "code", [B, heads, slice_num, dim_head]), optionally
    caching the slice weights; `decode` deslices a code with the cached weights and applies `to_out`; `reconstruct_fx`
    first replaces the cache by `project_slice(cache)`.  The cache keeps its autograd history like the reference's.

    `slice_weights` reads / writes the cache in the reference layout [B, heads, N, slice_num]; internally it is the
    kernels' [B, N, heads*slice_num]."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., slice_num=64, H=101, W=31, kernel=3):
        super().__init__(dim, heads=heads, dim_head=dim_head, dropout=dropout, slice_num=slice_num, H=H, W=W, kernel=kernel)
        self.project_slice = nn.Linear(slice_num, slice_num)
        self._w = None

    @property
    def slice_weights(self):
        if self._w is None:
            return None
        B, N, HG = self._w.shape
        return self._w.view(B, N, self.heads, HG // self.heads).permute(0, 2, 1, 3)

    @slice_weights.setter
    def slice_weights(self, value):
        if value is None:
            self._w = None
            return
        B, H, N, G = value.shape
        self._w = value.permute(0, 2, 1, 3).reshape(B, N, H * G).contiguous().float()

    def _check(self, t):
        if not t.is_cuda:
            raise RuntimeError("Physics-Attention (B200) has no CPU path: move the module and its input to a CUDA device")
        if self.training and self.dropout.p > 0.0:
            raise NotImplementedError("dropout > 0 in training mode is not supported by the fused kernels")

    def encode(self, x, cache_slice=False):
        self._check(x)
        grid = self._grid(tuple(x.shape))
        prec = ops.PRECISIONS[self.precision or config.get_default_precision()]
        code, w = ops.PaEncodeFn.apply(
            x.float(), self.temperature, self.in_project_x.weight, self.in_project_x.bias, self.in_project_fx.weight,
            self.in_project_fx.bias, self.in_project_slice.weight, self.in_project_slice.bias, self.to_q.weight, self.to_k.weight,
            self.to_v.weight, self.to_out[0].weight, self._packed_weights(), self.heads, grid, prec)
        if cache_slice:
            self._w = w
        return code

    def decode(self, code):
        self._check(code)
        if self._w is None:
            raise RuntimeError("decode() needs cached slice weights: call encode(x, cache_slice=True) or set slice_weights first")
        prec = ops.PRECISIONS[self.precision or config.get_default_precision()]
        lin = self.to_out[0]
        return ops.PaDecodeFn.apply(code.float(), self._w, lin.weight, lin.bias, prec)

    def reconstruct_fx(self, code):
        if self._w is None:
            raise RuntimeError("reconstruct_fx() needs cached slice weights: call encode(x, cache_slice=True) first")
        self._w = ops.SliceLinearFn.apply(self._w, self.project_slice.weight, self.project_slice.bias)
        return self.decode(code)


class Physics_Attention_Structured_Mesh_3D(_PhysicsAttentionBase):
    """for structured meshes in 3D space (reference :232-288): Conv3d 3x3x3 projections on tokens n = (h*W + w)*D + d,
    temperature clamped to [0.1, 5] (:268).  `D` is the mesh depth as in the reference (the head width is `dim_head`)."""
    structured = True

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., slice_num=32, H=32, W=32, D=32, kernel=3):
        super().__init__()
        if kernel != 3:
            raise NotImplementedError("only kernel=3 is supported (the reference never passes another value)")
        self.H, self.W, self.D = H, W, D
        self._build(dim, heads, dim_head, dropout, slice_num, lambda i, o: nn.Conv3d(i, o, kernel, 1, kernel // 2))
        self._pack3_key = None
        self._pack3 = None

    def _packed3(self):
        px, pfx = self.in_project_x, self.in_project_fx
        key = (px.weight.data_ptr(), px.weight._version, px.bias._version, pfx.weight.data_ptr(), pfx.weight._version,
               pfx.bias._version, px.weight.device, ops.cache_context())
        if key != self._pack3_key:
            with torch.no_grad(), torch.cuda.device(px.weight.device):
                bx, bfx = px.bias.detach().contiguous(), pfx.bias.detach().contiguous()
                self._pack3 = [ops.pack_proj_weights(px.weight.detach()[:, :, kh].contiguous(), bx,
                                                     pfx.weight.detach()[:, :, kh].contiguous(), bfx) for kh in range(3)]
            self._pack3_key = key
        return self._pack3

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("Physics-Attention (B200) has no CPU path: move the module and its input to a CUDA device")
        if self.training and self.dropout.p > 0.0:
            raise NotImplementedError("dropout > 0 in training mode is not supported by the fused kernels")
        B, N, C = x.shape
        if N != self.H * self.W * self.D:
            raise RuntimeError(f"shape '[{B}, {self.H}, {self.W}, {self.D}, {C}]' is invalid for input of size {B * N * C}")
        prec = ops.PRECISIONS[self.precision or config.get_default_precision()]
        lin = self.to_out[0]
        XF = ops.Conv3dProjFn.apply(x.float(), self.in_project_x.weight, self.in_project_x.bias, self.in_project_fx.weight,
                                    self.in_project_fx.bias, self._packed3(), (self.H, self.W, self.D), prec)
        code, w = ops.XFEncodeFn.apply(XF, self.temperature, self.in_project_slice.weight, self.in_project_slice.bias,
                                       self.to_q.weight, self.to_k.weight, self.to_v.weight, lin.weight, self.heads, True)
        return ops.PaDecodeFn.apply(code, w, lin.weight, lin.bias, prec)
