"""Transolver on a structured 2D mesh — drop-in for reference model/Transolver_Structured_Mesh_2D.py:122-220.
Same `Model(...)` keyword arguments, `forward(x, fx, T=None)` and state_dict keys; blocks run on libtbns."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import config, ops
from ._blocks import MLP, Transolver_block as _Block, init_weights, link_blocks, time_conditioning
from .Physics_Attention import Physics_Attention_Structured_Mesh_2D  # noqa: F401  (re-exported like the reference)


class Transolver_block(_Block):
    def __init__(self, num_heads, hidden_dim, dropout, act='gelu', mlp_ratio=4, last_layer=False, out_dim=1, slice_num=32,
                 H=85, W=85):
        super().__init__(num_heads, hidden_dim, dropout, act=act, mlp_ratio=mlp_ratio, last_layer=last_layer, out_dim=out_dim,
                         slice_num=slice_num, H=H, W=W, structured=True)


class Model(nn.Module):
    def __init__(self, space_dim=1, n_layers=5, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, act='gelu', mlp_ratio=1,
                 fun_dim=1, out_dim=1, slice_num=32, ref=8, unified_pos=False, H=85, W=85):
        super().__init__()
        self.__name__ = 'Transolver_2D'
        if Time_Input:
            self.time_fc = nn.Sequential(nn.Linear(n_hidden, n_hidden), nn.SiLU(), nn.Linear(n_hidden, n_hidden))
        self.H, self.W, self.ref, self.unified_pos = H, W, ref, unified_pos
        self.Time_Input, self.n_hidden, self.space_dim = Time_Input, n_hidden, space_dim
        in_dim = fun_dim + (ref * ref if unified_pos else space_dim)
        if unified_pos:
            self.pos = self.get_grid()
        self.preprocess = MLP(in_dim, n_hidden * 2, n_hidden, n_layers=0, res=False, act=act)
        self.blocks = nn.ModuleList([
            Transolver_block(num_heads=n_head, hidden_dim=n_hidden, dropout=dropout, act=act, mlp_ratio=mlp_ratio, out_dim=out_dim,
                             slice_num=slice_num, H=H, W=W, last_layer=(i == n_layers - 1)) for i in range(n_layers)])
        link_blocks(self.blocks)
        self.initialize_weights()
        self.placeholder = nn.Parameter((1 / n_hidden) * torch.rand(n_hidden, dtype=torch.float))

    def initialize_weights(self):
        init_weights(self)

    def get_grid(self, batchsize=1):
        """distance of every mesh point to a ref x ref lattice on [0,1]^2 -> [batch, H, W, ref*ref]
        (a plain tensor attribute like in the reference, which builds it on the GPU; here it follows the input's device)."""
        gx = torch.tensor(np.linspace(0, 1, self.H), dtype=torch.float)
        gy = torch.tensor(np.linspace(0, 1, self.W), dtype=torch.float)
        mesh = torch.stack(torch.meshgrid(gx, gy, indexing="ij"), -1)            # [H, W, 2]
        rx = torch.tensor(np.linspace(0, 1, self.ref), dtype=torch.float)
        lattice = torch.stack(torch.meshgrid(rx, rx, indexing="ij"), -1)         # [ref, ref, 2]
        d = torch.sqrt(((mesh[:, :, None, None, :] - lattice[None, None]) ** 2).sum(-1))
        return d.reshape(1, self.H, self.W, self.ref * self.ref).repeat(batchsize, 1, 1, 1).contiguous()

    def _pos16(self, device):
        """bf16 copy [N, ref*ref] of the unified-position table: the operand the packed preprocess reads (broadcast over the
        batch inside the pack kernel instead of `pos.repeat(B)`, reference :204)"""
        t = getattr(self, "_pos16_cache", None)
        if t is None or t.device != device:
            t = self.pos.to(device).reshape(self.H * self.W, self.ref * self.ref).to(torch.bfloat16).contiguous()
            self._pos16_cache = t
        return t

    def _packed_preprocess_ok(self, x, fx):
        pp = self.preprocess
        return (fx is not None and fx.is_cuda and fx.dtype == torch.float32 and config.get_default_precision() == "bf16"
                and pp.n_layers == 0 and pp.act_name == 'gelu' and ops.mlp_tc_ok(pp.n_input, pp.n_hidden, pp.n_output))

    def forward(self, x, fx, T=None, out=None):
        """reference signature forward(x, fx, T=None).  `out` (extension, inference only): a [B, N, out_dim] view - e.g. one
        column of a frame history - that receives the prediction in place (train.rollout)."""
        if self._packed_preprocess_ok(x, fx):
            # bf16 mode: position features / coordinates and the input fields are packed straight into the bf16 operand of
            # the first Linear (no pos.repeat, no cat, no padded copy; `fx` may be a strided window of a frame history)
            pre, post = self.preprocess.linear_pre[0], self.preprocess.linear_post
            tab16 = self._pos16(fx.device) if self.unified_pos else None
            src1 = None if self.unified_pos else x
            ln1 = self.blocks[0].ln_1 if (T is None and isinstance(self.blocks[0], _Block)) else None   # computed by the second Linear
            # all blocks' derived weights (packed projections, bf16 MLP operands) are refreshed on the side stream beside the
            # preprocess MLP; each block then finds them in the cache
            forked = ops.refresh_weights_ahead(self, self.blocks, ops.TBNS_PREC_BF16)
            fx = ops.PackedMlpFn.apply(tab16, src1, fx, pre.weight, pre.bias, post.weight, post.bias,
                                       *((ln1.weight, ln1.bias, ln1.eps) if ln1 is not None else (None, None, 1e-5)))
            if forked:
                ops._join_side()
        else:
            if self.unified_pos:
                if self.pos.device != x.device:
                    self.pos = self.pos.to(x.device)
                x = self.pos.repeat(x.shape[0], 1, 1, 1).reshape(x.shape[0], self.H * self.W, self.ref * self.ref)
            if fx is not None:
                fx = self.preprocess(torch.cat((x, fx), -1))
            else:
                fx = self.preprocess(x) + self.placeholder[None, None, :]
        if T is not None:
            fx = fx + time_conditioning(self.time_fc, T, self.n_hidden)
        for block in self.blocks[:-1]:
            fx = block(fx)
        return self.blocks[-1](fx) if out is None else self.blocks[-1](fx, out=out)
