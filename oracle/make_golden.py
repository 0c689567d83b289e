"""Generate tests/golden/*.pt from the LIVE reference modules (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden
Every fixture holds: constructor kwargs, the reference module's state_dict, seeded inputs, the
reference output, a seeded upstream gradient and every gradient autograd produced — all computed by
the unmodified reference code in float64 (so fp32 kernels can be judged against a clean target).
"""
from __future__ import annotations

import os

import torch

from . import reference_shim as ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _run(module, inputs, seed):
    module = module.double()
    inputs = [None if t is None else t.double().requires_grad_(True) for t in inputs]
    out = module(*inputs)
    g = torch.Generator().manual_seed(seed + 1)
    dout = torch.randn(out.shape, generator=g, dtype=torch.float64)
    out.backward(dout)
    return dict(
        state={k: v.detach().clone() for k, v in module.state_dict().items()},
        inputs=[None if t is None else t.detach().clone() for t in inputs],
        out=out.detach().clone(), dout=dout,
        dinputs=[None if t is None else t.grad.clone() for t in inputs],
        grads={k: (p.grad.clone() if p.grad is not None else None) for k, p in module.named_parameters()},
    )


def pa_structured(name, seed, dim, heads, dim_head, G, Hg, Wg, B, taus=None, slice_gain=1.0):
    PA = ref.physics_attention()
    torch.manual_seed(seed)
    m = PA.Physics_Attention_Structured_Mesh_2D(dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg)
    with torch.no_grad():
        if taus is not None:
            m.temperature.copy_(torch.tensor(taus).reshape(1, heads, 1, 1))
        m.in_project_slice.weight.mul_(slice_gain)
        m.in_project_slice.bias.normal_(0, 0.3)
    x = torch.randn(B, Hg * Wg, dim)
    fx = _run(m, [x], seed)
    fx["kwargs"] = dict(dim=dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg)
    fx["kind"] = "Physics_Attention_Structured_Mesh_2D"
    torch.save(fx, os.path.join(OUT, name))


def pa_irregular(name, seed, dim, heads, dim_head, G, N, B, taus=None, slice_gain=1.0):
    PA = ref.physics_attention()
    torch.manual_seed(seed)
    m = PA.Physics_Attention_Irregular_Mesh(dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G)
    with torch.no_grad():
        if taus is not None:
            m.temperature.copy_(torch.tensor(taus).reshape(1, heads, 1, 1))
        m.in_project_slice.weight.mul_(slice_gain)
        m.in_project_slice.bias.normal_(0, 0.3)
    x = torch.randn(B, N, dim)
    fx = _run(m, [x], seed)
    fx["kwargs"] = dict(dim=dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G)
    fx["kind"] = "Physics_Attention_Irregular_Mesh"
    torch.save(fx, os.path.join(OUT, name))


def pa_autoencoder(name, seed, dim, heads, dim_head, G, Hg, Wg, B):
    """Physics_Attention_Structured_Mesh_2D_Auto_Encoder (model/Physics_Attention.py:122-227) in the call order of
    Transolver_Encoder_block.decode (model/Transolver_Structured_Mesh2D_Encoder.py:86-94): encode(cache) -> reconstruct_fx
    (replaces the cached slice weights by project_slice(weights)) -> decode (uses the REPLACED cache)."""
    PA = ref.physics_attention()
    torch.manual_seed(seed)
    m = PA.Physics_Attention_Structured_Mesh_2D_Auto_Encoder(dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg)
    with torch.no_grad():
        m.in_project_slice.weight.mul_(3.0)
        m.in_project_slice.bias.normal_(0, 0.3)
        m.temperature.copy_(torch.linspace(0.3, 1.2, heads).reshape(1, heads, 1, 1))
    m = m.double()
    x = torch.randn(B, Hg * Wg, dim, dtype=torch.float64, requires_grad=True)
    g = torch.Generator().manual_seed(seed + 1)
    fwd = m(x).detach().clone()                       # plain forward == the 2D module's
    code = m.encode(x, cache_slice=True)
    w_enc = m.slice_weights.detach().clone()          # [B,H,N,G]
    rec = m.reconstruct_fx(code)
    w_proj = m.slice_weights.detach().clone()
    dec = m.decode(code)
    r1 = torch.randn(rec.shape, generator=g, dtype=torch.float64)
    r2 = torch.randn(dec.shape, generator=g, dtype=torch.float64)
    ((rec * r1).sum() + (dec * r2).sum()).backward()
    torch.save(dict(kind="Physics_Attention_Structured_Mesh_2D_Auto_Encoder",
                    kwargs=dict(dim=dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg),
                    state={k: v.detach().clone() for k, v in m.state_dict().items()}, x=x.detach().clone(), fwd=fwd,
                    code=code.detach().clone(), w_enc=w_enc, w_proj=w_proj, rec=rec.detach().clone(), dec=dec.detach().clone(),
                    r1=r1, r2=r2, dx=x.grad.clone(),
                    grads={k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}),
               os.path.join(OUT, name))


def pa_structured3d(name, seed, dim, heads, dim_head, G, Hg, Wg, Dg, B):
    PA = ref.physics_attention()
    torch.manual_seed(seed)
    m = PA.Physics_Attention_Structured_Mesh_3D(dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg, D=Dg)
    with torch.no_grad():
        m.in_project_slice.weight.mul_(3.0)
        m.in_project_slice.bias.normal_(0, 0.3)
        m.temperature.copy_(torch.linspace(0.08, 1.2, heads).reshape(1, heads, 1, 1))   # first head below the clamp bound
    x = torch.randn(B, Hg * Wg * Dg, dim)
    fx = _run(m, [x], seed)
    fx["kwargs"] = dict(dim=dim, heads=heads, dim_head=dim_head, dropout=0.0, slice_num=G, H=Hg, W=Wg, D=Dg)
    fx["kind"] = "Physics_Attention_Structured_Mesh_3D"
    torch.save(fx, os.path.join(OUT, name))


def block(name, seed, structured, hidden, heads, G, last, B, Hg=None, Wg=None, N=None, mlp_ratio=1, out_dim=1):
    torch.manual_seed(seed)
    if structured:
        T = ref.transolver_2d()
        m = T.Transolver_block(num_heads=heads, hidden_dim=hidden, dropout=0.0, act="gelu", mlp_ratio=mlp_ratio,
                               last_layer=last, out_dim=out_dim, slice_num=G, H=Hg, W=Wg)
        N = Hg * Wg
    else:
        T = ref.transolver_irregular()
        m = T.Transolver_block(num_heads=heads, hidden_dim=hidden, dropout=0.0, act="gelu", mlp_ratio=mlp_ratio,
                               last_layer=last, out_dim=out_dim, slice_num=G)
    with torch.no_grad():  # non-trivial LN affine + biases so every gradient path is exercised
        for k, p in m.named_parameters():
            if k.startswith("ln_") and k.endswith("weight"):
                p.add_(0.2 * torch.randn_like(p))
            if k.startswith("ln_") and k.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
    fx = torch.randn(B, N, hidden)
    d = _run(m, [fx], seed)
    d["kwargs"] = dict(num_heads=heads, hidden_dim=hidden, dropout=0.0, act="gelu", mlp_ratio=mlp_ratio,
                       last_layer=last, out_dim=out_dim, slice_num=G)
    if structured:
        d["kwargs"].update(H=Hg, W=Wg)
    d["kind"] = "block_structured" if structured else "block_irregular"
    torch.save(d, os.path.join(OUT, name))


def ckpt_block(name, seed, layer=3, Hg=12, Wg=10, B=1):
    """Attention module of one trained block of checkpoints/ep400_sim100.pt (C=64,H=8,D=8,G=32) on a small grid
    (the module is translation invariant, only reshape depends on H,W)."""
    sd = torch.load(os.path.join(ref.REF_ROOT, "checkpoints", "ep400_sim100.pt"), map_location="cpu", weights_only=True)
    pref = f"blocks.{layer}.Attn."
    asd = {k[len(pref):]: v for k, v in sd.items() if k.startswith(pref)}
    PA = ref.physics_attention()
    m = PA.Physics_Attention_Structured_Mesh_2D(64, heads=8, dim_head=8, dropout=0.0, slice_num=32, H=Hg, W=Wg)
    m.load_state_dict(asd, strict=True)
    torch.manual_seed(seed)
    x = torch.nn.functional.layer_norm(torch.randn(B, Hg * Wg, 64), (64,))
    d = _run(m, [x], seed)
    d["kwargs"] = dict(dim=64, heads=8, dim_head=8, dropout=0.0, slice_num=32, H=Hg, W=Wg)
    d["kind"] = "Physics_Attention_Structured_Mesh_2D"
    d["source"] = f"checkpoints/ep400_sim100.pt blocks.{layer}.Attn"
    torch.save(d, os.path.join(OUT, name))


def _sharpen(m):
    """trunc_normal(0.02) init leaves the slice softmax and the token attention numerically uniform, which makes
    their gradients pure cancellation noise; give the fixture models trained-like magnitudes instead."""
    with torch.no_grad():
        for blk in m.blocks:
            a = blk.Attn
            a.in_project_slice.weight.mul_(40.0)
            a.in_project_slice.bias.normal_(0, 0.3)
            a.to_q.weight.mul_(20.0)
            a.to_k.weight.mul_(20.0)
            a.to_v.weight.mul_(5.0)
            a.temperature.copy_(torch.linspace(0.3, 1.2, a.heads).reshape(1, -1, 1, 1))
            a.in_project_x.weight.mul_(3.0)


def model_2d(name, seed, unified_pos, rollout_steps=3):
    T = ref.transolver_2d()
    torch.manual_seed(seed)
    kw = dict(space_dim=2, n_layers=2, n_hidden=32, dropout=0.0, n_head=4, Time_Input=False, mlp_ratio=1, fun_dim=3,
              out_dim=1, slice_num=8, ref=4, unified_pos=unified_pos, H=8, W=7)
    with ref.cpu_cuda_identity():
        m = T.Model(**kw).double()
    _sharpen(m)
    if unified_pos:
        m.pos = m.pos.double()
    x = torch.rand(2, 56, 2).double()
    fx = torch.randn(2, 56, 3).double()
    y = torch.randn(2, 56, 1).double()
    loss_fn = ref.testloss().TestLoss(size_average=False)
    out = m(x, fx)
    loss = loss_fn(out.reshape(2, -1), y.reshape(2, -1))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    # closed-loop rollout (exp_ns.py:225-241 / ns_vorticity_unrolling.py:264-286), step = 1
    with torch.no_grad():
        f = fx.clone()
        preds = []
        for _ in range(rollout_steps):
            im = m(x, fx=f)
            preds.append(im)
            f = torch.cat((f[..., 1:], im), -1)
        roll = torch.cat(preds, -1)
    torch.save(dict(kind="model_2d", kwargs=kw, state={k: v.detach().clone() for k, v in m.state_dict().items()},
                    x=x, fx=fx, y=y, out=out.detach(), loss=loss.detach(), grads=grads, rollout=roll),
               os.path.join(OUT, name))


def model_2d_time(name, seed):
    """Time_Input=True (exp_plas.py:148,187): forward(x, fx, T) with the sinusoidal timestep embedding + time_fc"""
    T = ref.transolver_2d()
    torch.manual_seed(seed)
    kw = dict(space_dim=2, n_layers=2, n_hidden=32, dropout=0.0, n_head=4, Time_Input=True, mlp_ratio=1, fun_dim=3,
              out_dim=1, slice_num=8, ref=4, unified_pos=0, H=6, W=7)
    m = T.Model(**kw).double()
    m.time_fc.float()     # the reference builds the embedding with `.float()` (model/Embedding.py:81): time_fc can only run in fp32
    _sharpen(m)
    with torch.no_grad():
        for p in m.time_fc.parameters():
            p.normal_(0, 0.3)
    x = torch.rand(2, 42, 2).double()
    fx = torch.randn(2, 42, 3).double()
    y = torch.randn(2, 42, 1).double()
    tt = torch.tensor([[3.0], [7.5]], dtype=torch.float64)
    loss_fn = ref.testloss().TestLoss(size_average=False)
    out = m(x, fx, T=tt)
    loss = loss_fn(out.reshape(2, -1), y.reshape(2, -1))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    torch.save(dict(kind="model_2d", kwargs=kw, state={k: v.detach().clone() for k, v in m.state_dict().items()},
                    x=x, fx=fx, y=y, T=tt, out=out.detach(), loss=loss.detach(), grads=grads), os.path.join(OUT, name))


def model_irregular(name, seed):
    T = ref.transolver_irregular()
    torch.manual_seed(seed)
    kw = dict(space_dim=2, n_layers=2, n_hidden=32, dropout=0.0, n_head=4, Time_Input=False, mlp_ratio=2, fun_dim=0,
              out_dim=1, slice_num=8, ref=8, unified_pos=0)
    m = T.Model(**kw).double()
    _sharpen(m)
    x = torch.rand(1, 45, 2).double()
    y = torch.randn(1, 45, 1).double()
    loss_fn = ref.testloss().TestLoss(size_average=False)
    out = m(x, None)
    loss = loss_fn(out.reshape(1, -1), y.reshape(1, -1))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    torch.save(dict(kind="model_irregular", kwargs=kw, state={k: v.detach().clone() for k, v in m.state_dict().items()},
                    x=x, y=y, out=out.detach(), loss=loss.detach(), grads=grads), os.path.join(OUT, name))


def main():
    assert ref.available(), "reference not mounted"
    os.makedirs(OUT, exist_ok=True)
    # taus straddle the clamp bounds [0.1, 5]: below, at, inside, above (grad 0 outside; SURVEY §8 a-bwd)
    pa_structured("pa_structured_small.pt", 11, dim=32, heads=4, dim_head=8, G=4, Hg=6, Wg=5, B=2,
                  taus=[0.05, 0.1, 0.7, 7.0], slice_gain=3.0)
    pa_structured("pa_structured_g64.pt", 12, dim=32, heads=2, dim_head=16, G=64, Hg=5, Wg=9, B=1, slice_gain=5.0)
    pa_irregular("pa_irregular_small.pt", 13, dim=32, heads=4, dim_head=8, G=8, N=37, B=2,
                 taus=[0.05, 0.3, 1.0, 7.0], slice_gain=3.0)
    pa_irregular("pa_irregular_inner_ne_dim.pt", 14, dim=24, heads=2, dim_head=16, G=16, N=50, B=1)
    block("block_structured_mid.pt", 21, True, hidden=32, heads=4, G=8, last=False, B=2, Hg=4, Wg=7)
    block("block_structured_last.pt", 22, True, hidden=32, heads=4, G=8, last=True, B=1, Hg=5, Wg=5, out_dim=2)
    block("block_irregular_last.pt", 23, False, hidden=32, heads=2, G=16, last=True, B=1, N=41, mlp_ratio=2)
    pa_autoencoder("pa_autoencoder_small.pt", 15, dim=32, heads=4, dim_head=8, G=8, Hg=6, Wg=5, B=2)
    pa_autoencoder("pa_autoencoder_head1.pt", 16, dim=32, heads=1, dim_head=32, G=16, Hg=7, Wg=4, B=1)   # shipped encoder checkpoints' shape
    pa_structured3d("pa_structured3d_small.pt", 17, dim=16, heads=2, dim_head=8, G=8, Hg=4, Wg=3, Dg=5, B=2)
    ckpt_block("pa_ckpt_ep400_block3.pt", 31)
    model_2d("model_2d_unified.pt", 41, unified_pos=1)
    model_2d("model_2d_plainpos.pt", 42, unified_pos=0)
    model_irregular("model_irregular.pt", 43)
    model_2d_time("model_2d_time.pt", 44)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
