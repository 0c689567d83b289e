"""Transolver on irregular meshes / point clouds — drop-in for reference model/Transolver_Irregular_Mesh.py:74-158."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ._blocks import MLP, Transolver_block as _Block, init_weights, link_blocks, time_conditioning
from .Physics_Attention import Physics_Attention_Irregular_Mesh  # noqa: F401


class Transolver_block(_Block):
    def __init__(self, num_heads, hidden_dim, dropout, act='gelu', mlp_ratio=4, last_layer=False, out_dim=1, slice_num=32):
        super().__init__(num_heads, hidden_dim, dropout, act=act, mlp_ratio=mlp_ratio, last_layer=last_layer, out_dim=out_dim,
                         slice_num=slice_num, structured=False)


class Model(nn.Module):
    def __init__(self, space_dim=1, n_layers=5, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, act='gelu', mlp_ratio=1,
                 fun_dim=1, out_dim=1, slice_num=32, ref=8, unified_pos=False):
        super().__init__()
        self.__name__ = 'Transolver_1D'
        if Time_Input:
            self.time_fc = nn.Sequential(nn.Linear(n_hidden, n_hidden), nn.SiLU(), nn.Linear(n_hidden, n_hidden))
        self.ref, self.unified_pos = ref, unified_pos
        self.Time_Input, self.n_hidden, self.space_dim = Time_Input, n_hidden, space_dim
        in_dim = fun_dim + (ref * ref if unified_pos else space_dim)
        self.preprocess = MLP(in_dim, n_hidden * 2, n_hidden, n_layers=0, res=False, act=act)
        self.blocks = nn.ModuleList([
            Transolver_block(num_heads=n_head, hidden_dim=n_hidden, dropout=dropout, act=act, mlp_ratio=mlp_ratio, out_dim=out_dim,
                             slice_num=slice_num, last_layer=(i == n_layers - 1)) for i in range(n_layers)])
        link_blocks(self.blocks)
        self.initialize_weights()
        self.placeholder = nn.Parameter((1 / n_hidden) * torch.rand(n_hidden, dtype=torch.float))

    def initialize_weights(self):
        init_weights(self)

    def get_grid(self, x, batchsize=1):
        """x [B,N,2] -> distances to a ref x ref lattice on [0,1]^2, [B,N,ref*ref]"""
        rx = torch.tensor(np.linspace(0, 1, self.ref), dtype=torch.float, device=x.device)
        lattice = torch.stack(torch.meshgrid(rx, rx, indexing="ij"), -1).reshape(1, 1, self.ref * self.ref, 2)
        return torch.sqrt(((x[:, :, None, :] - lattice) ** 2).sum(-1)).contiguous()

    def forward(self, x, fx, T=None):
        if self.unified_pos:
            x = self.get_grid(x, x.shape[0])
        fx = self.preprocess(torch.cat((x, fx), -1) if fx is not None else x)
        fx = fx + self.placeholder[None, None, :]   # irregular model adds it unconditionally (reference :148)
        if T is not None:
            fx = fx + time_conditioning(self.time_fc, T, self.n_hidden)
        for block in self.blocks:
            fx = block(fx)
        return fx
