"""n-step differentiable autoregressive unroll around the 2D Transolver — drop-in for reference
model/SOL_Transolver_Structured_Mesh_2D.py:6-52 (attributes `.transolver_model`, `.n`, `.step`)."""
import torch
import torch.nn as nn

from .Transolver_Structured_Mesh_2D import Model as transolver_model


class SOL_Transolver_Structured_Mesh_2D(nn.Module):
    def __init__(self, space_dim=1, n_layers=5, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, act='gelu', mlp_ratio=1,
                 fun_dim=1, out_dim=1, slice_num=32, ref=8, unified_pos=False, H=85, W=85, step=1, look_ahead=5):
        super().__init__()
        self.transolver_model = transolver_model(space_dim=space_dim, n_layers=n_layers, n_hidden=n_hidden, dropout=dropout,
                                                 n_head=n_head, Time_Input=Time_Input, act=act, mlp_ratio=mlp_ratio, fun_dim=fun_dim,
                                                 out_dim=out_dim, slice_num=slice_num, ref=ref, unified_pos=unified_pos, H=H, W=W)
        self.n = look_ahead   # number of chained model calls per forward
        self.step = step      # scalar fields produced per call

    def forward(self, x, fx):
        u = None
        for _ in range(self.n):
            u = self.transolver_model(x, fx=fx)
            fx = torch.cat((fx[..., self.step:], u), dim=-1)   # feed the prediction back in
        return u
