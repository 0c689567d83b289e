"""Transolver auto-encoder on a structured 2D mesh - drop-in for reference model/Transolver_Structured_Mesh2D_Encoder.py:
`Transolver_Encoder_block` (:41-96) and `Model` (:99-226) with `forward`, `encode`, `decode` and the slice-weight accessors.
The last block's attention is cut after the slice-token stage (encode) and restarted at the deslice (decode)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config, ops
from ._blocks import MLP, init_weights
from .Physics_Attention import Physics_Attention_Structured_Mesh_2D_Auto_Encoder
from .Transolver_Structured_Mesh_2D import Model as _Model2D


class Transolver_Encoder_block(nn.Module):
    def __init__(self, num_heads, hidden_dim, dropout, act='gelu', mlp_ratio=4, last_layer=False, out_dim=1, slice_num=32, H=85, W=85):
        super().__init__()
        if act != 'gelu':
            raise NotImplementedError("the fused block implements act='gelu' (the only activation the reference scripts use)")
        self.last_layer = last_layer
        self.ln_1 = nn.LayerNorm(hidden_dim)
        self.Attn = Physics_Attention_Structured_Mesh_2D_Auto_Encoder(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                                                      dropout=dropout, slice_num=slice_num, H=H, W=W)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim, hidden_dim * mlp_ratio, hidden_dim, n_layers=0, res=False, act=act)
        if last_layer:
            self.ln_3 = nn.LayerNorm(hidden_dim)
            self.mlp2 = nn.Linear(hidden_dim, out_dim)

    def _prec(self):
        return ops.PRECISIONS[self.Attn.precision or config.get_default_precision()]

    def _mlp(self, fx):
        pre, post = self.mlp.linear_pre[0], self.mlp.linear_post
        return ops.LnMlpFn.apply(fx, self.ln_2.weight, self.ln_2.bias, pre.weight, pre.bias, post.weight, post.bias, self.ln_2.eps,
                                 self._prec())

    def forward(self, fx):
        if self.last_layer:                       # reference :72-76
            return self.decode(self.encode(fx))
        return self._mlp(self.Attn.forward_block(fx.contiguous(), self.ln_1))

    def encode(self, fx):
        if self.last_layer:                       # reference :80-82
            x1 = ops.LayerNormFn.apply(fx.contiguous().float(), self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
            return self.Attn.encode(x1, cache_slice=True)
        return self._mlp(self.Attn.forward_block(fx.contiguous(), self.ln_1))

    def decode(self, code):
        if not self.last_layer:                   # reference :87-90 prints and returns None
            print("the model has to be the last layer")
            return None
        fx = self.Attn.reconstruct_fx(code)       # replaces the cached slice weights by project_slice(weights)  (:92)
        fx = self.Attn.decode(code) + fx          # (:93)
        fx = self._mlp(fx)
        return ops.LnLinearFn.apply(fx, self.ln_3.weight, self.ln_3.bias, self.mlp2.weight, self.mlp2.bias, self.ln_3.eps, self._prec())


class Model(_Model2D):
    def __init__(self, space_dim=1, n_layers=5, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, act='gelu', mlp_ratio=1,
                 fun_dim=1, out_dim=1, slice_num=32, ref=8, unified_pos=False, H=85, W=85):
        super().__init__(space_dim=space_dim, n_layers=0, n_hidden=n_hidden, dropout=dropout, n_head=n_head, Time_Input=Time_Input,
                         act=act, mlp_ratio=mlp_ratio, fun_dim=fun_dim, out_dim=out_dim, slice_num=slice_num, ref=ref,
                         unified_pos=unified_pos, H=H, W=W)
        self.blocks = nn.ModuleList([
            Transolver_Encoder_block(num_heads=n_head, hidden_dim=n_hidden, dropout=dropout, act=act, mlp_ratio=mlp_ratio,
                                     out_dim=out_dim, slice_num=slice_num, H=H, W=W, last_layer=(i == n_layers - 1))
            for i in range(n_layers)])
        init_weights(self.blocks)

    def _embed(self, x, fx):
        if self.unified_pos:
            if self.pos.device != x.device:
                self.pos = self.pos.to(x.device)
            x = self.pos.repeat(x.shape[0], 1, 1, 1).reshape(x.shape[0], self.H * self.W, self.ref * self.ref)
        if fx is not None:
            return self.preprocess(torch.cat((x, fx), -1))
        return self.preprocess(x) + self.placeholder[None, None, :]

    def encode(self, x, fx):                      # reference :200-213
        h = self._embed(x, fx)
        for block in self.blocks:
            h = block.encode(h)
        return h

    def decode(self, code):                       # reference :215-216
        return self.blocks[-1].decode(code)

    def get_attention_slice(self):
        return self.blocks[-1].Attn.slice_weights

    def set_attention_slice(self, slice):
        self.blocks[-1].Attn.slice_weights = slice
