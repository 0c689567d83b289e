"""CPU oracle for the Physics-Attention hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch, stage-by-stage restatement of the reference algorithm
(`/root/reference/model/Physics_Attention.py:6-57` irregular, `:60-119` structured 2D) and of the
Transolver block around it (`model/Transolver_Structured_Mesh_2D.py:13-75`), cut into exactly the
stages the CUDA kernels implement, with an explicit hand-derived backward for every stage
(SURVEY.md §8 a-bwd).  It runs in whatever dtype its inputs have (tests use fp64 and fp32).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it; the product package never does (it raises when the CUDA library is missing).

Parity pin: the reference ships no tests / golden vectors.  The pins are
(1) `tests/golden/*.pt`, produced by `oracle/make_golden.py` from the *live* reference modules
imported from /root/reference in the build container, and (2) the live-reference cases of `tests/test_oracle.py`,
which re-import the reference whenever /root/reference (or the byte-compiled `oracle/_ref`) is present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

# Timing switch for bench.py's cpu_baseline / --impl reference leg only: evaluate the 3x3 projections with the same
# library op the reference calls (nn.Conv2d -> F.conv2d, Physics_Attention.py:94-96) instead of the nine shifted
# matmuls, so the CPU baseline is not handicapped by the restatement.  Parity tests keep it False.
USE_LIBRARY_CONV = False

EPS_NORM = 1e-5  # `slice_norm + 1e-5`, Physics_Attention.py:43 / :102
LN_EPS = 1e-5  # nn.LayerNorm default, Transolver_Structured_Mesh_2D.py:59


# ------------------------------------------------------------------------------------------------
# stage 0: LayerNorm (Transolver_Structured_Mesh_2D.py:59,63,66 ; forward :70-73)
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps: float = LN_EPS):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (x - mean) * rstd
    return xhat * gamma + beta, mean.squeeze(-1), rstd.squeeze(-1)


def layernorm_bwd(dy, x, mean, rstd, gamma):
    xhat = (x - mean.unsqueeze(-1)) * rstd.unsqueeze(-1)
    dgamma = (dy * xhat).reshape(-1, x.shape[-1]).sum(0)
    dbeta = dy.reshape(-1, x.shape[-1]).sum(0)
    g = dy * gamma
    c1 = g.mean(-1, keepdim=True)
    c2 = (g * xhat).mean(-1, keepdim=True)
    dx = (g - c1 - xhat * c2) * rstd.unsqueeze(-1)
    return dx, dgamma, dbeta


# ------------------------------------------------------------------------------------------------
# stage 1a: projections x -> XF = [x_mid | fx_mid]   (Physics_Attention.py:36-39 Linear, :94-97 Conv2d)
# The 3x3 / pad 1 convolution is restated as nine shifted matmuls (the implicit-GEMM form the CUDA
# kernel uses): out[b,i,j,:] = sum_{ky,kx} x[b,i+ky-1,j+kx-1,:] @ W[:,:,ky,kx]^T, zero outside.
# ------------------------------------------------------------------------------------------------
def _shift2d(x4, dy: int, dx: int):
    """x4 [B,Hg,Wg,C] -> y with y[b,i,j] = x4[b,i+dy,j+dx] (zero outside)."""
    B, Hg, Wg, C = x4.shape
    y = torch.zeros_like(x4)
    i0, i1 = max(0, -dy), min(Hg, Hg - dy)
    j0, j1 = max(0, -dx), min(Wg, Wg - dx)
    if i1 > i0 and j1 > j0:
        y[:, i0:i1, j0:j1] = x4[:, i0 + dy:i1 + dy, j0 + dx:j1 + dx]
    return y


def proj_fwd(x, Wx, bx, Wfx, bfx, grid: Optional[Tuple[int, int]] = None):
    """x [B,N,C] -> XF [B,N,2I].  grid=(Hg,Wg) selects the structured (conv) variant."""
    B, N, C = x.shape
    Wcat = torch.cat([Wx, Wfx], 0)  # [2I, C] or [2I, C, 3, 3]
    bcat = torch.cat([bx, bfx], 0)
    if grid is None:
        return x @ Wcat.t() + bcat
    Hg, Wg = grid
    if Hg * Wg != N:
        raise RuntimeError(f"shape '[{B}, {Hg}, {Wg}, {C}]' is invalid for input of size {x.numel()}")
    x4 = x.reshape(B, Hg, Wg, C)
    if USE_LIBRARY_CONV:
        y = torch.nn.functional.conv2d(x4.permute(0, 3, 1, 2).contiguous(), Wcat, bcat, stride=1, padding=1)
        return y.permute(0, 2, 3, 1).contiguous().reshape(B, N, -1)
    out = torch.zeros(B, Hg, Wg, Wcat.shape[0], dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            out = out + _shift2d(x4, ky - 1, kx - 1) @ Wcat[:, :, ky, kx].t()
    return (out + bcat).reshape(B, N, -1)


def proj_bwd(dXF, x, Wx, Wfx, grid: Optional[Tuple[int, int]] = None):
    """returns dx, dWx, dbx, dWfx, dbfx (SURVEY §8 a-bwd last line)."""
    B, N, C = x.shape
    I = Wx.shape[0]
    Wcat = torch.cat([Wx, Wfx], 0)
    dbcat = dXF.reshape(-1, 2 * I).sum(0)
    if grid is None:
        dx = dXF @ Wcat
        dWcat = dXF.reshape(-1, 2 * I).t() @ x.reshape(-1, C)
    else:
        Hg, Wg = grid
        x4 = x.reshape(B, Hg, Wg, C)
        d4 = dXF.reshape(B, Hg, Wg, 2 * I)
        dx4 = torch.zeros_like(x4)
        dWcat = torch.zeros_like(Wcat)
        for ky in range(3):
            for kx in range(3):
                # out[i,j] += x[i+ky-1, j+kx-1] W_tap^T  =>  dx[p,q] += dOut[p-(ky-1), q-(kx-1)] W_tap
                dx4 = dx4 + _shift2d(d4, -(ky - 1), -(kx - 1)) @ Wcat[:, :, ky, kx]
                xs = _shift2d(x4, ky - 1, kx - 1).reshape(-1, C)
                dWcat[:, :, ky, kx] = d4.reshape(-1, 2 * I).t() @ xs
        dx = dx4.reshape(B, N, C)
    return dx, dWcat[:I], dbcat[:I], dWcat[I:], dbcat[I:]


# ------------------------------------------------------------------------------------------------
# stage 1b/1c: slice weights, slice norm, un-normalised slice tokens
#   (Physics_Attention.py:40-42 irregular, :98-101 structured)
# ------------------------------------------------------------------------------------------------
def clamp_temperature(temperature, clamp: bool):
    t = temperature.reshape(-1)
    return torch.clamp(t, min=0.1, max=5.0) if clamp else t


def slice_fwd(XF, Ws, bs, temperature, heads: int, clamp: bool):
    """XF [B,N,2I] -> w [B,N,H,G], s [B,H,G], Tt [B,H,G,D]."""
    B, N, I2 = XF.shape
    I = I2 // 2
    D = I // heads
    X = XF[..., :I].reshape(B, N, heads, D)
    F = XF[..., I:].reshape(B, N, heads, D)
    tau = clamp_temperature(temperature, clamp)  # [H]
    logits = (X @ Ws.t() + bs) / tau[None, None, :, None]
    w = torch.softmax(logits, -1)
    s = w.sum(1)  # [B,H,G]
    Tt = torch.einsum("bnhg,bnhd->bhgd", w, F)
    return w, s, Tt


def slice_bwd(dw_flat, dTt, ds, XF, Ws, bs, temperature, heads: int, clamp: bool):
    """dw_flat [B,N,H,G] is the deslice gradient, dTt [B,H,G,D], ds [B,H,G].
    returns dXF, dWs, dbs, dtemperature (shape of `temperature`)."""
    B, N, I2 = XF.shape
    I = I2 // 2
    D = I // heads
    X = XF[..., :I].reshape(B, N, heads, D)
    F = XF[..., I:].reshape(B, N, heads, D)
    tau = clamp_temperature(temperature, clamp)
    L = X @ Ws.t() + bs  # pre-temperature logits [B,N,H,G]
    w = torch.softmax(L / tau[None, None, :, None], -1)
    dw = dw_flat + torch.einsum("bnhd,bhgd->bnhg", F, dTt) + ds[:, None]
    dF = torch.einsum("bnhg,bhgd->bnhd", w, dTt)
    dLp = w * (dw - (dw * w).sum(-1, keepdim=True))  # grad wrt L/tau
    dL = dLp / tau[None, None, :, None]
    dX = dL @ Ws  # [B,N,H,D]
    dWs = torch.einsum("bnhg,bnhd->gd", dL, X)
    dbs = dL.sum((0, 1, 2))
    dtau = -(dLp * L).sum((0, 1, 3)) / tau ** 2
    if clamp:
        t = temperature.reshape(-1)
        dtau = dtau * ((t >= 0.1) & (t <= 5.0)).to(dtau.dtype)
    dXF = torch.cat([dX.reshape(B, N, I), dF.reshape(B, N, I)], -1)
    return dXF, dWs, dbs, dtau.reshape(temperature.shape)


# ------------------------------------------------------------------------------------------------
# stage 2: token normalisation + attention among slice tokens (+ fold of to_out into P)
#   (Physics_Attention.py:43-52 / :102-111 ; fold: SURVEY §7 "useful algebra")
# ------------------------------------------------------------------------------------------------
def token_attn_fwd(s, Tt, Wq, Wk, Wv, Wo):
    """s [B,H,G], Tt [B,H,G,D], Wo [C_out, I] -> dict(tok,q,k,v,A,O,P) with P [B,H*G,C_out]."""
    B, H, G, D = Tt.shape
    tok = Tt / (s + EPS_NORM)[..., None]
    q, k, v = tok @ Wq.t(), tok @ Wk.t(), tok @ Wv.t()
    A = torch.softmax((q @ k.transpose(-1, -2)) * (D ** -0.5), -1)
    O = A @ v  # [B,H,G,D]
    Wo_h = Wo.reshape(Wo.shape[0], H, D)  # [C,H,D]
    P = torch.einsum("bhgd,chd->bhgc", O, Wo_h).reshape(B, H * G, Wo.shape[0])
    return dict(tok=tok, q=q, k=k, v=v, A=A, O=O, P=P)


def token_attn_bwd(dP, s, Tt, st, Wq, Wk, Wv, Wo):
    """dP [B,H*G,C] -> dTt, ds, dWq, dWk, dWv, dWo."""
    B, H, G, D = Tt.shape
    C = Wo.shape[0]
    tok, q, k, v, A, O = st["tok"], st["q"], st["k"], st["v"], st["A"], st["O"]
    dP4 = dP.reshape(B, H, G, C)
    Wo_h = Wo.reshape(C, H, D)
    dO = torch.einsum("bhgc,chd->bhgd", dP4, Wo_h)
    dWo = torch.einsum("bhgc,bhgd->chd", dP4, O).reshape(C, H * D)
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
    dTt = dtok * inv[..., None]
    ds = -(dtok * tok).sum(-1) * inv
    return dTt, ds, dWq, dWk, dWv, dWo


# ------------------------------------------------------------------------------------------------
# stage 3: deslice merged with to_out (Physics_Attention.py:55-57 / :116-119)
# ------------------------------------------------------------------------------------------------
def deslice_out_fwd(w, P, bo, residual=None):
    B, N, H, G = w.shape
    out = w.reshape(B, N, H * G) @ P + bo
    return out if residual is None else out + residual


def deslice_out_bwd(dout, w, P):
    B, N, H, G = w.shape
    wf = w.reshape(B, N, H * G)
    dP = wf.transpose(1, 2) @ dout  # [B,HG,C]
    dw = (dout @ P.transpose(1, 2)).reshape(B, N, H, G)
    dbo = dout.reshape(-1, dout.shape[-1]).sum(0)
    return dw, dP, dbo


# ------------------------------------------------------------------------------------------------
# full attention module
# ------------------------------------------------------------------------------------------------
PA_KEYS = ("temperature", "in_project_x.weight", "in_project_x.bias", "in_project_fx.weight",
           "in_project_fx.bias", "in_project_slice.weight", "in_project_slice.bias", "to_q.weight",
           "to_k.weight", "to_v.weight", "to_out.0.weight", "to_out.0.bias")


def pa_forward(x, p: dict, heads: int, grid=None, residual=None):
    """p: state_dict of a reference attention module (keys PA_KEYS). Returns (out, saved)."""
    clamp = grid is not None  # structured clamps tau (:99), irregular does not (:40)
    XF = proj_fwd(x, p["in_project_x.weight"], p["in_project_x.bias"], p["in_project_fx.weight"],
                  p["in_project_fx.bias"], grid)
    w, s, Tt = slice_fwd(XF, p["in_project_slice.weight"], p["in_project_slice.bias"], p["temperature"], heads, clamp)
    st = token_attn_fwd(s, Tt, p["to_q.weight"], p["to_k.weight"], p["to_v.weight"], p["to_out.0.weight"])
    out = deslice_out_fwd(w, st["P"], p["to_out.0.bias"], residual)
    return out, dict(x=x, XF=XF, w=w, s=s, Tt=Tt, st=st, grid=grid, heads=heads, clamp=clamp)


# ------------------------------------------------------------------------------------------------
# auto-encoder variant: the pipeline cut after the token stage / restarted at the deslice
# (Physics_Attention_Structured_Mesh_2D_Auto_Encoder, model/Physics_Attention.py:122-227).
# Plain differentiable torch: parity tests take gradients of these functions with autograd in fp64.
# Slice weights are kept in the [B,N,H,G] layout of this file (the reference caches [B,H,N,G]).
# ------------------------------------------------------------------------------------------------
def ae_encode(x, p: dict, heads: int, grid):
    """encode() :185-212 -> (code = out_slice_token [B,H,G,D], slice weights [B,N,H,G])"""
    XF = proj_fwd(x, p["in_project_x.weight"], p["in_project_x.bias"], p["in_project_fx.weight"], p["in_project_fx.bias"], grid)
    w, s, Tt = slice_fwd(XF, p["in_project_slice.weight"], p["in_project_slice.bias"], p["temperature"], heads, True)
    st = token_attn_fwd(s, Tt, p["to_q.weight"], p["to_k.weight"], p["to_v.weight"], p["to_out.0.weight"])
    return st["O"], w


def ae_project_slice(w, p: dict):
    """reconstruct_fx() :215 - nn.Linear(slice_num, slice_num) over the slice index of the cached weights"""
    return w @ p["project_slice.weight"].t() + p["project_slice.bias"]


def ae_decode(code, w, p: dict):
    """decode() :221-227 (and the tail of reconstruct_fx :217-219): deslice `code` with the given weights, then to_out"""
    B, H, G, D = code.shape
    Wo = p["to_out.0.weight"]
    P = torch.einsum("bhgd,chd->bhgc", code, Wo.reshape(Wo.shape[0], H, D)).reshape(B, H * G, Wo.shape[0])
    return deslice_out_fwd(w, P, p["to_out.0.bias"])


# ------------------------------------------------------------------------------------------------
# structured 3D variant (Physics_Attention_Structured_Mesh_3D, model/Physics_Attention.py:232-288): Conv3d 3x3x3 / pad 1
# projections on tokens ordered n = (h*W + w)*D + d; everything after the projections is the 2D algebra (temperature
# clamped).  Differentiable torch (library conv3d): parity tests take its gradients with autograd in fp64.
# ------------------------------------------------------------------------------------------------
def pa3d_forward(x, p: dict, heads: int, grid3):
    Hg, Wg, Dg = grid3
    B, N, C = x.shape
    x5 = x.reshape(B, Hg, Wg, Dg, C).permute(0, 4, 1, 2, 3)
    conv = torch.nn.functional.conv3d
    xm = conv(x5, p["in_project_x.weight"], p["in_project_x.bias"], padding=1).permute(0, 2, 3, 4, 1).reshape(B, N, -1)
    fm = conv(x5, p["in_project_fx.weight"], p["in_project_fx.bias"], padding=1).permute(0, 2, 3, 4, 1).reshape(B, N, -1)
    XF = torch.cat([xm, fm], -1)
    w, s, Tt = slice_fwd(XF, p["in_project_slice.weight"], p["in_project_slice.bias"], p["temperature"], heads, True)
    st = token_attn_fwd(s, Tt, p["to_q.weight"], p["to_k.weight"], p["to_v.weight"], p["to_out.0.weight"])
    return deslice_out_fwd(w, st["P"], p["to_out.0.bias"])


def pa_backward(dout, p: dict, sv: dict):
    """returns (dx, grads dict keyed like PA_KEYS)."""
    dw, dP, dbo = deslice_out_bwd(dout, sv["w"], sv["st"]["P"])
    dTt, ds, dWq, dWk, dWv, dWo = token_attn_bwd(dP, sv["s"], sv["Tt"], sv["st"], p["to_q.weight"], p["to_k.weight"],
                                                 p["to_v.weight"], p["to_out.0.weight"])
    dXF, dWs, dbs, dtemp = slice_bwd(dw, dTt, ds, sv["XF"], p["in_project_slice.weight"], p["in_project_slice.bias"],
                                     p["temperature"], sv["heads"], sv["clamp"])
    dx, dWx, dbx, dWfx, dbfx = proj_bwd(dXF, sv["x"], p["in_project_x.weight"], p["in_project_fx.weight"], sv["grid"])
    g = {"temperature": dtemp, "in_project_x.weight": dWx, "in_project_x.bias": dbx, "in_project_fx.weight": dWfx,
         "in_project_fx.bias": dbfx, "in_project_slice.weight": dWs, "in_project_slice.bias": dbs, "to_q.weight": dWq,
         "to_k.weight": dWk, "to_v.weight": dWv, "to_out.0.weight": dWo, "to_out.0.bias": dbo}
    return dx, g


# ------------------------------------------------------------------------------------------------
# LN + MLP epilogue (Transolver_Structured_Mesh_2D.py:13-38 with n_layers=0, act=gelu(erf); :71-73)
# ------------------------------------------------------------------------------------------------
def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def ln_mlp_fwd(fx, g, b, W1, b1, W2, b2, residual: bool = True):
    x2, mean, rstd = layernorm_fwd(fx, g, b)
    pre = x2 @ W1.t() + b1
    hid = gelu(pre)
    out = hid @ W2.t() + b2
    if residual:
        out = out + fx
    return out, dict(fx=fx, x2=x2, mean=mean, rstd=rstd, pre=pre, hid=hid, residual=residual)


def ln_mlp_bwd(dout, g, W1, W2, sv):
    C = dout.shape[-1]
    dW2 = dout.reshape(-1, C).t() @ sv["hid"].reshape(-1, sv["hid"].shape[-1])
    db2 = dout.reshape(-1, C).sum(0)
    dpre = (dout @ W2) * gelu_grad(sv["pre"])
    dW1 = dpre.reshape(-1, dpre.shape[-1]).t() @ sv["x2"].reshape(-1, C)
    db1 = dpre.reshape(-1, dpre.shape[-1]).sum(0)
    dx2 = dpre @ W1
    dfx, dg, db = layernorm_bwd(dx2, sv["fx"], sv["mean"], sv["rstd"], g)
    if sv["residual"]:
        dfx = dfx + dout
    return dfx, dict(ln_w=dg, ln_b=db, W1=dW1, b1=db1, W2=dW2, b2=db2)


def ln_linear_fwd(fx, g, b, W, bias):
    """last layer: mlp2(ln_3(fx)) (Transolver_Structured_Mesh_2D.py:72-73)."""
    x3, mean, rstd = layernorm_fwd(fx, g, b)
    return x3 @ W.t() + bias, dict(fx=fx, x3=x3, mean=mean, rstd=rstd)


def ln_linear_bwd(dout, g, W, sv):
    C = sv["fx"].shape[-1]
    dW = dout.reshape(-1, dout.shape[-1]).t() @ sv["x3"].reshape(-1, C)
    dbias = dout.reshape(-1, dout.shape[-1]).sum(0)
    dx3 = dout @ W
    dfx, dg, db = layernorm_bwd(dx3, sv["fx"], sv["mean"], sv["rstd"], g)
    return dfx, dict(ln_w=dg, ln_b=db, W=dW, b=dbias)


# ------------------------------------------------------------------------------------------------
# Transolver block (Transolver_Structured_Mesh_2D.py:69-75)
# ------------------------------------------------------------------------------------------------
def split_block_state(sd: dict):
    """block state_dict -> (attn dict, others)"""
    attn = {k[len("Attn."):]: v for k, v in sd.items() if k.startswith("Attn.")}
    return attn, sd


def block_forward(fx, sd: dict, heads: int, grid=None):
    attn, _ = split_block_state(sd)
    x1, m1, r1 = layernorm_fwd(fx, sd["ln_1.weight"], sd["ln_1.bias"])
    fx2, sv_a = pa_forward(x1, attn, heads, grid, residual=fx)
    fx3, sv_m = ln_mlp_fwd(fx2, sd["ln_2.weight"], sd["ln_2.bias"], sd["mlp.linear_pre.0.weight"],
                           sd["mlp.linear_pre.0.bias"], sd["mlp.linear_post.weight"], sd["mlp.linear_post.bias"])
    sv = dict(fx=fx, m1=m1, r1=r1, attn=sv_a, mlp=sv_m)
    if "ln_3.weight" in sd:
        out, sv_l = ln_linear_fwd(fx3, sd["ln_3.weight"], sd["ln_3.bias"], sd["mlp2.weight"], sd["mlp2.bias"])
        sv["last"] = sv_l
        return out, sv
    return fx3, sv


def block_backward(dout, sd: dict, sv: dict):
    attn, _ = split_block_state(sd)
    g = {}
    if "last" in sv:
        dout, gl = ln_linear_bwd(dout, sd["ln_3.weight"], sd["mlp2.weight"], sv["last"])
        g.update({"ln_3.weight": gl["ln_w"], "ln_3.bias": gl["ln_b"], "mlp2.weight": gl["W"], "mlp2.bias": gl["b"]})
    dfx2, gm = ln_mlp_bwd(dout, sd["ln_2.weight"], sd["mlp.linear_pre.0.weight"], sd["mlp.linear_post.weight"], sv["mlp"])
    g.update({"ln_2.weight": gm["ln_w"], "ln_2.bias": gm["ln_b"], "mlp.linear_pre.0.weight": gm["W1"],
              "mlp.linear_pre.0.bias": gm["b1"], "mlp.linear_post.weight": gm["W2"], "mlp.linear_post.bias": gm["b2"]})
    dx1, ga = pa_backward(dfx2, attn, sv["attn"])
    g.update({"Attn." + k: v for k, v in ga.items()})
    dfx, dg1, db1 = layernorm_bwd(dx1, sv["fx"], sv["m1"], sv["r1"], sd["ln_1.weight"])
    g.update({"ln_1.weight": dg1, "ln_1.bias": db1})
    return dfx + dfx2, g


# ------------------------------------------------------------------------------------------------
# loss (utils/testloss.py:31-42 with size_average=False)
# ------------------------------------------------------------------------------------------------
def rel_l2_sum(x, y):
    n = x.shape[0]
    d = torch.linalg.vector_norm(x.reshape(n, -1) - y.reshape(n, -1), dim=1)
    return (d / torch.linalg.vector_norm(y.reshape(n, -1), dim=1)).sum()


def rel_l2(a, b) -> float:
    """global relative L2 between two tensors (the parity metric of BASELINE.json north_star)."""
    a = a.detach().double().reshape(-1)
    b = b.detach().double().reshape(-1)
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-300))
