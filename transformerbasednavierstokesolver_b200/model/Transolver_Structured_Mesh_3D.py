"""Transolver on a structured 3D mesh - drop-in for reference model/Transolver_Structured_Mesh_3D.py (`Transolver_block`
:40-75, `Model` :78-191): same keyword arguments, `forward(x, fx, T=None)` and state_dict keys."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import config, ops
from ._blocks import MLP, init_weights, time_conditioning
from .Physics_Attention import Physics_Attention_Structured_Mesh_3D


class Transolver_block(nn.Module):
    def __init__(self, num_heads, hidden_dim, dropout, act='gelu', mlp_ratio=4, last_layer=False, out_dim=1, slice_num=32,
                 H=32, W=32, D=32):
        super().__init__()
        if act != 'gelu':
            raise NotImplementedError("the fused block implements act='gelu' (the only activation the reference scripts use)")
        self.last_layer = last_layer
        self.ln_1 = nn.LayerNorm(hidden_dim)
        self.Attn = Physics_Attention_Structured_Mesh_3D(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                                         dropout=dropout, slice_num=slice_num, H=H, W=W, D=D)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim, hidden_dim * mlp_ratio, hidden_dim, n_layers=0, res=False, act=act)
        if last_layer:
            self.ln_3 = nn.LayerNorm(hidden_dim)
            self.mlp2 = nn.Linear(hidden_dim, out_dim)

    def forward(self, fx):
        prec = ops.PRECISIONS[self.Attn.precision or config.get_default_precision()]
        fx = fx.contiguous().float()
        fx = self.Attn(ops.LayerNormFn.apply(fx, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)) + fx
        pre, post = self.mlp.linear_pre[0], self.mlp.linear_post
        fx = ops.LnMlpFn.apply(fx, self.ln_2.weight, self.ln_2.bias, pre.weight, pre.bias, post.weight, post.bias, self.ln_2.eps, prec)
        if self.last_layer:
            return ops.LnLinearFn.apply(fx, self.ln_3.weight, self.ln_3.bias, self.mlp2.weight, self.mlp2.bias, self.ln_3.eps, prec)
        return fx


class Model(nn.Module):
    def __init__(self, space_dim=1, n_layers=5, n_hidden=256, dropout=0, n_head=8, Time_Input=False, act='gelu', mlp_ratio=1,
                 fun_dim=1, out_dim=1, slice_num=32, ref=8, unified_pos=False, H=32, W=32, D=32):
        super().__init__()
        self.__name__ = 'Transolver_3D'
        self.use_checkpoint = False
        self.H, self.W, self.D, self.ref, self.unified_pos = H, W, D, ref, unified_pos
        self.Time_Input, self.n_hidden, self.space_dim = Time_Input, n_hidden, space_dim
        in_dim = fun_dim + (ref ** 3 if unified_pos else space_dim)
        if unified_pos:
            self.pos = self.get_grid()
        self.preprocess = MLP(in_dim, n_hidden * 2, n_hidden, n_layers=0, res=False, act=act)
        if Time_Input:
            self.time_fc = nn.Sequential(nn.Linear(n_hidden, n_hidden), nn.SiLU(), nn.Linear(n_hidden, n_hidden))
        self.blocks = nn.ModuleList([
            Transolver_block(num_heads=n_head, hidden_dim=n_hidden, dropout=dropout, act=act, mlp_ratio=mlp_ratio, out_dim=out_dim,
                             slice_num=slice_num, H=H, W=W, D=D, last_layer=(i == n_layers - 1)) for i in range(n_layers)])
        init_weights(self)
        self.placeholder = nn.Parameter((1 / n_hidden) * torch.rand(n_hidden, dtype=torch.float))

    def get_grid(self, batchsize=1):
        """distance of every mesh point to a ref^3 lattice on [0,1]^3 -> [batch, H, W, D, ref^3] (reference :147-168)"""
        axes = [torch.tensor(np.linspace(0, 1, n), dtype=torch.float) for n in (self.H, self.W, self.D)]
        mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), -1)                       # [H, W, D, 3]
        r = torch.tensor(np.linspace(0, 1, self.ref), dtype=torch.float)
        lattice = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1)                  # [ref, ref, ref, 3]
        d = torch.sqrt(((mesh[:, :, :, None, None, None, :] - lattice[None, None, None]) ** 2).sum(-1))
        return d.reshape(1, self.H, self.W, self.D, self.ref ** 3).repeat(batchsize, 1, 1, 1, 1).contiguous()

    def forward(self, x, fx, T=None):
        if self.unified_pos:
            if self.pos.device != x.device:
                self.pos = self.pos.to(x.device)
            x = self.pos.repeat(x.shape[0], 1, 1, 1, 1).reshape(x.shape[0], self.H * self.W * self.D, self.ref ** 3)
        if fx is not None:
            fx = self.preprocess(torch.cat((x, fx), -1))
        else:
            fx = self.preprocess(x) + self.placeholder[None, None, :]
        if T is not None:
            fx = fx + time_conditioning(self.time_fc, T, self.n_hidden)
        for block in self.blocks:
            if self.use_checkpoint and torch.is_grad_enabled():
                # activation checkpointing as in the reference (:186-187): the block's kernels re-run in backward
                import torch.utils.checkpoint as checkpoint
                fx = checkpoint.checkpoint(block, fx, use_reentrant=False)
            else:
                fx = block(fx)
        return fx
