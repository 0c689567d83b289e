"""Shared pieces of the Transolver models: the small MLP container (checkpoint-compatible names), the
fused Transolver block and weight init.  Reference: model/Transolver_Structured_Mesh_2D.py:13-75,
model/Transolver_Irregular_Mesh.py:12-71."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from .Physics_Attention import Physics_Attention_Irregular_Mesh, Physics_Attention_Structured_Mesh_2D

ACTIVATION = {'gelu': nn.GELU, 'tanh': nn.Tanh, 'sigmoid': nn.Sigmoid, 'relu': nn.ReLU, 'leaky_relu': nn.LeakyReLU(0.1),
              'softplus': nn.Softplus, 'ELU': nn.ELU, 'silu': nn.SiLU}


class MLP(nn.Module):
    """linear_pre -> [linears] -> linear_post; parameter names follow the reference so checkpoints load.
    `preprocess` runs through ops.MlpFn in bf16 mode (PyTorch modules in fp32 mode); inside a block it is only a parameter container — the block runs
    the fused LayerNorm+MLP kernels instead."""

    def __init__(self, n_input, n_hidden, n_output, n_layers=1, act='gelu', res=True):
        super().__init__()
        if act not in ACTIVATION:
            raise NotImplementedError
        self.act_name = act
        act_cls = ACTIVATION[act]
        self.n_input, self.n_hidden, self.n_output, self.n_layers, self.res = n_input, n_hidden, n_output, n_layers, res
        self.linear_pre = nn.Sequential(nn.Linear(n_input, n_hidden), act_cls())
        self.linear_post = nn.Linear(n_hidden, n_output)
        self.linears = nn.ModuleList([nn.Sequential(nn.Linear(n_hidden, n_hidden), act_cls()) for _ in range(n_layers)])

    def forward(self, x):
        pre, post = self.linear_pre[0], self.linear_post
        if (x.is_cuda and self.n_layers == 0 and self.act_name == 'gelu' and config.get_default_precision() == "bf16"
                and x.dtype == torch.float32 and ops.mlp_tc_ok(self.n_input, self.n_hidden, self.n_output)):
            # bf16 mode: both Linears on the tensor cores (libtbns), GELU fused into the first one's epilogue
            return ops.MlpFn.apply(x, pre.weight, pre.bias, post.weight, post.bias)
        x = self.linear_pre(x)
        for layer in self.linears:
            x = layer(x) + x if self.res else layer(x)
        return self.linear_post(x)


class Transolver_block(nn.Module):
    """fx = Attn(ln_1(fx)) + fx ; fx = mlp(ln_2(fx)) + fx ; last layer: mlp2(ln_3(fx)).
    LayerNorm, residual adds and the MLP run in libtbns kernels (ops.LayerNormFn / LnMlpFn / LnLinearFn)."""

    def __init__(self, num_heads: int, hidden_dim: int, dropout: float, act='gelu', mlp_ratio=4, last_layer=False, out_dim=1,
                 slice_num=32, H=None, W=None, structured=True):
        super().__init__()
        if act != 'gelu':
            raise NotImplementedError("the fused block implements act='gelu' (the only activation the reference scripts use)")
        self.last_layer = last_layer
        self.ln_1 = nn.LayerNorm(hidden_dim)
        if structured:
            self.Attn = Physics_Attention_Structured_Mesh_2D(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                                             dropout=dropout, slice_num=slice_num, H=H, W=W)
        else:
            self.Attn = Physics_Attention_Irregular_Mesh(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                                         dropout=dropout, slice_num=slice_num)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim, hidden_dim * mlp_ratio, hidden_dim, n_layers=0, res=False, act=act)
        if last_layer:
            self.ln_3 = nn.LayerNorm(hidden_dim)
            self.mlp2 = nn.Linear(hidden_dim, out_dim)
        self._next_ln = ()   # (ln_1 of the following block,): set by link_blocks; a tuple so that it is not a registered submodule

    def forward(self, fx, out=None):
        prec = ops.PRECISIONS[self.Attn.precision or config.get_default_precision()]
        # each LayerNorm is computed by the GEMM that produces its input (bf16 mode): ln_2 in the attention's output
        # projection, the next block's ln_1 in this block's second MLP Linear
        fx = self.Attn.forward_block(fx.contiguous(), self.ln_1, self.ln_2)
        pre, post = self.mlp.linear_pre[0], self.mlp.linear_post
        nxt = self._next_ln[0] if (self._next_ln and not self.last_layer) else None
        fx = ops.LnMlpFn.apply(fx, self.ln_2.weight, self.ln_2.bias, pre.weight, pre.bias, post.weight, post.bias, self.ln_2.eps, prec,
                               *((nxt.weight, nxt.bias, nxt.eps) if nxt is not None else (None, None, 1e-5)))
        if self.last_layer:
            if out is not None:
                if torch.is_grad_enabled() and fx.requires_grad:
                    raise RuntimeError("`out=` writes the prediction in place and is for inference (torch.no_grad()) only")
                return ops.ln_linear_into(fx, self.ln_3.weight, self.ln_3.bias, self.mlp2.weight, self.mlp2.bias, self.ln_3.eps, prec, out)
            return ops.LnLinearFn.apply(fx, self.ln_3.weight, self.ln_3.bias, self.mlp2.weight, self.mlp2.bias, self.ln_3.eps, prec)
        return fx


def link_blocks(blocks):
    """tell every block which LayerNorm consumes its output (the next block's ln_1), so that its last GEMM can compute it"""
    blocks = list(blocks)
    for a, b in zip(blocks[:-1], blocks[1:]):
        if isinstance(a, Transolver_block) and isinstance(b, Transolver_block):
            a._next_ln = (b.ln_1,)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """sinusoidal embedding of (possibly fractional) time indices - reference model/Embedding.py:67-85 (fp32, cos | sin)"""
    import math
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[..., :1])], dim=-1)
    return emb


def time_conditioning(time_fc: nn.Module, T: torch.Tensor, n_hidden: int) -> torch.Tensor:
    """Time_Input=True (exp_plas.py:148,187): time_fc(timestep_embedding(T)) for T [B,1] -> [B,1,n_hidden], broadcast over the
    mesh points by the caller (the reference repeats the embedding N times BEFORE the two Linears:
    model/Transolver_Structured_Mesh_2D.py:212-215; per-row Linears commute with the repeat).  A few [B, n_hidden] torch ops
    outside the attention path."""
    return time_fc(timestep_embedding(T, n_hidden))


def init_weights(module: nn.Module):
    """same distributions as the reference `_init_weights` (Transolver_Structured_Mesh_2D.py:174-181):
    Linear ~ trunc_normal(std 0.02), zero bias; LayerNorm affine = (1, 0); Conv2d keeps the torch default."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, (nn.LayerNorm, nn.BatchNorm1d)):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
