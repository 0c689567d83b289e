"""CPU oracle of the full Transolver model + training step around the hot path.  TEST INFRASTRUCTURE ONLY
(also the `cpu_baseline` / `--impl reference` leg of bench.py, kind "port": the reference is Python and cannot
travel to the GPU box, so its algorithm is timed through this restatement on the box's host cores).

Restates  model/Transolver_Structured_Mesh_2D.py:183-220 (Model.get_grid / forward),
          model/Transolver_Irregular_Mesh.py:140-157, exp_ns.py:191-218 (teacher-forced training step),
          exp_ns.py:225-241 (closed-loop rollout)  on top of oracle/physics_attention.py.
All functions are differentiable torch code, so torch autograd provides the CPU backward for timing."""
from __future__ import annotations

import numpy as np
import torch

from . import physics_attention as O


def unified_pos_table(Hg: int, Wg: int, ref: int, dtype=torch.float32):
    gx = torch.tensor(np.linspace(0, 1, Hg), dtype=dtype)
    gy = torch.tensor(np.linspace(0, 1, Wg), dtype=dtype)
    mesh = torch.stack(torch.meshgrid(gx, gy, indexing="ij"), -1)
    r = torch.tensor(np.linspace(0, 1, ref), dtype=dtype)
    lat = torch.stack(torch.meshgrid(r, r, indexing="ij"), -1)
    return torch.sqrt(((mesh[:, :, None, None] - lat[None, None]) ** 2).sum(-1)).reshape(1, Hg * Wg, ref * ref)


def mlp_fwd(x, sd, prefix):
    h = O.gelu(x @ sd[prefix + "linear_pre.0.weight"].t() + sd[prefix + "linear_pre.0.bias"])
    return h @ sd[prefix + "linear_post.weight"].t() + sd[prefix + "linear_post.bias"]


def timestep_embedding(timesteps, dim: int, max_period: float = 10000.0):
    """sinusoidal embedding of (possibly fractional) time indices, model/Embedding.py:67-85: [..., dim] = cos | sin"""
    half = dim // 2
    freqs = torch.exp(-np.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.dtype)
    args = timesteps[:, None] * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], -1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[..., :1])], -1)
    return emb


def model_forward(x, fx, sd: dict, n_layers: int, heads: int, grid=None, unified_pos=False, ref=8, irregular=False, T=None):
    if unified_pos and grid is not None:
        x = unified_pos_table(grid[0], grid[1], ref, x.dtype).repeat(x.shape[0], 1, 1)
    if fx is not None:
        h = mlp_fwd(torch.cat((x, fx), -1), sd, "preprocess.")
        if irregular:
            h = h + sd["placeholder"][None, None, :]
    else:
        h = mlp_fwd(x, sd, "preprocess.") + sd["placeholder"][None, None, :]
    if T is not None:   # Time_Input=True (exp_plas.py:148,187): model/Transolver_Structured_Mesh_2D.py:212-215
        # the reference computes the embedding and time_fc in fp32 whatever the model dtype (Embedding.py:81 `.float()`)
        e = timestep_embedding(T.float(), h.shape[-1])             # T [B,1] -> [B,1,C]; the reference repeats it over N first
        e = torch.nn.functional.silu(e @ sd["time_fc.0.weight"].float().t() + sd["time_fc.0.bias"].float())
        h = h + (e @ sd["time_fc.2.weight"].float().t() + sd["time_fc.2.bias"].float()).to(h.dtype)
    for i in range(n_layers):
        pre = f"blocks.{i}."
        bsd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        h, _ = O.block_forward(h, bsd, heads, grid)
    return h


def teacher_forced_loss(x, fx, yy, fwd, T: int, step: int = 1):
    """exp_ns.py:191-208: sum over T/step calls of the summed relative L2, ground truth fed back."""
    loss = 0.0
    bsz = x.shape[0]
    for t in range(0, T, step):
        y = yy[..., t:t + step]
        im = fwd(x, fx)
        loss = loss + O.rel_l2_sum(im.reshape(bsz, -1), y.reshape(bsz, -1))
        fx = torch.cat((fx[..., step:], y), dim=-1)
    return loss


def rollout(x, fx, fwd, T: int, step: int = 1):
    """exp_ns.py:225-241 / ns_vorticity_unrolling.py:264-286: predictions fed back."""
    preds = []
    for _ in range(0, T, step):
        im = fwd(x, fx)
        preds.append(im)
        fx = torch.cat((fx[..., step:], im), dim=-1)
    return torch.cat(preds, -1)
