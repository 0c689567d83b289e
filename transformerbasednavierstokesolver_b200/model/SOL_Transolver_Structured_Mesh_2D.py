"""n-step differentiable autoregressive unroll around the 2D Transolver - drop-in for reference
model/SOL_Transolver_Structured_Mesh_2D.py:6-52: attributes `.transolver_model`, `.n` (= look_ahead), `.step`; every other
constructor argument is the 2D `Model`'s and is passed through unchanged."""
import torch
import torch.nn as nn

from . import Transolver_Structured_Mesh_2D as _m2d


class SOL_Transolver_Structured_Mesh_2D(nn.Module):
    def __init__(self, *model_args, step=1, look_ahead=5, **model_kwargs):
        super().__init__()
        self.transolver_model = _m2d.Model(*model_args, **model_kwargs)
        self.n, self.step = look_ahead, step   # chained model calls per forward; scalar fields produced per call

    def forward(self, x, fx):
        """look_ahead chained calls; the window of input fields slides by `step` channels and takes the prediction in"""
        window, pred = fx, None
        for _ in range(self.n):
            pred = self.transolver_model(x, fx=window)
            window = torch.cat((window[..., self.step:], pred), dim=-1)
        return pred
