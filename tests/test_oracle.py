"""CPU tests: the oracle restatement (forward and hand-derived backward) against the golden vectors
that oracle/make_golden.py produced from the live reference modules, and against the reference itself
when /root/reference is mounted (build container)."""
import pytest
import torch

from oracle import physics_attention as O
from oracle import reference_shim as ref

TOL = 1e-10  # float64 restatement vs float64 reference


def _grid(fx):
    kw = fx["kwargs"]
    return (kw["H"], kw["W"]) if "H" in kw else None


@pytest.mark.parametrize("name", ["pa_structured_small.pt", "pa_structured_g64.pt", "pa_irregular_small.pt",
                                  "pa_irregular_inner_ne_dim.pt", "pa_ckpt_ep400_block3.pt"])
def test_pa_oracle_matches_golden(golden, name):
    fx = golden(name)
    out, sv = O.pa_forward(fx["inputs"][0], fx["state"], fx["kwargs"]["heads"], _grid(fx))
    assert O.rel_l2(out, fx["out"]) < TOL
    dx, g = O.pa_backward(fx["dout"], fx["state"], sv)
    assert O.rel_l2(dx, fx["dinputs"][0]) < TOL
    for k in O.PA_KEYS:
        ref_g = fx["grads"][k]
        if float(ref_g.abs().max()) == 0.0:
            assert float(g[k].abs().max()) == 0.0, k
        else:
            assert O.rel_l2(g[k], ref_g) < 1e-9, k


def test_temperature_clamp_gradient_is_zero_outside(golden):
    fx = golden("pa_structured_small.pt")  # taus = 0.05, 0.1, 0.7, 7.0
    g = fx["grads"]["temperature"].reshape(-1)
    assert g[0] == 0 and g[3] == 0 and g[1] != 0 and g[2] != 0
    _, sv = O.pa_forward(fx["inputs"][0], fx["state"], 4, _grid(fx))
    _, og = O.pa_backward(fx["dout"], fx["state"], sv)
    og = og["temperature"].reshape(-1)
    assert og[0] == 0 and og[3] == 0


@pytest.mark.parametrize("name", ["block_structured_mid.pt", "block_structured_last.pt", "block_irregular_last.pt"])
def test_block_oracle_matches_golden(golden, name):
    fx = golden(name)
    out, sv = O.block_forward(fx["inputs"][0], fx["state"], fx["kwargs"]["num_heads"], _grid(fx))
    assert O.rel_l2(out, fx["out"]) < TOL
    dfx, g = O.block_backward(fx["dout"], fx["state"], sv)
    assert O.rel_l2(dfx, fx["dinputs"][0]) < TOL
    for k, ref_g in fx["grads"].items():
        assert O.rel_l2(g[k], ref_g) < 1e-9, k


def test_structured_rejects_wrong_grid(golden):
    fx = golden("pa_structured_small.pt")
    with pytest.raises(RuntimeError):
        O.pa_forward(fx["inputs"][0][:, :-1], fx["state"], 4, _grid(fx))


@pytest.mark.skipif(not ref.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("structured", [True, False])
def test_oracle_vs_live_reference_fp32(structured):
    PA = ref.physics_attention()
    torch.manual_seed(5)
    if structured:
        m = PA.Physics_Attention_Structured_Mesh_2D(64, heads=8, dim_head=8, dropout=0.0, slice_num=32, H=9, W=11)
        x = torch.randn(2, 99, 64)
        grid = (9, 11)
    else:
        m = PA.Physics_Attention_Irregular_Mesh(64, heads=4, dim_head=16, dropout=0.0, slice_num=16)
        x = torch.randn(2, 77, 64)
        grid = None
    x.requires_grad_(True)
    y = m(x)
    dout = torch.randn_like(y)
    y.backward(dout)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    out, sv = O.pa_forward(x.detach(), sd, m.heads, grid)
    assert O.rel_l2(out, y) < 2e-6
    dx, g = O.pa_backward(dout, sd, sv)
    assert O.rel_l2(dx, x.grad) < 2e-5
    for k, p in m.named_parameters():
        assert O.rel_l2(g[k], p.grad) < 5e-5, k


def test_library_conv_switch_is_equivalent(golden):
    fx = golden("pa_structured_small.pt")
    a, _ = O.pa_forward(fx["inputs"][0], fx["state"], 4, _grid(fx))
    O.USE_LIBRARY_CONV = True
    try:
        b, _ = O.pa_forward(fx["inputs"][0], fx["state"], 4, _grid(fx))
    finally:
        O.USE_LIBRARY_CONV = False
    assert O.rel_l2(b, a) < 1e-12


@pytest.mark.parametrize("name", ["pa_autoencoder_small.pt", "pa_autoencoder_head1.pt"])
def test_autoencoder_oracle_matches_golden(golden, name):
    """Physics_Attention_Structured_Mesh_2D_Auto_Encoder (model/Physics_Attention.py:122-227): encode / reconstruct_fx /
    decode in the order Transolver_Encoder_block.decode uses them; gradients of the oracle by autograd in fp64."""
    fx = golden(name)
    kw = fx["kwargs"]
    p = {k: v.clone().requires_grad_(True) for k, v in fx["state"].items()}
    x = fx["x"].clone().requires_grad_(True)
    out, _ = O.pa_forward(x, p, kw["heads"], (kw["H"], kw["W"]))
    assert O.rel_l2(out.detach(), fx["fwd"]) < TOL
    code, w = O.ae_encode(x, p, kw["heads"], (kw["H"], kw["W"]))
    assert O.rel_l2(code.detach(), fx["code"]) < TOL
    assert O.rel_l2(w.detach().permute(0, 2, 1, 3), fx["w_enc"]) < TOL          # reference caches [B,H,N,G]
    wp = O.ae_project_slice(w, p)
    assert O.rel_l2(wp.detach().permute(0, 2, 1, 3), fx["w_proj"]) < TOL
    rec, dec = O.ae_decode(code, wp, p), O.ae_decode(code, wp, p)                # decode runs on the REPLACED cache
    assert O.rel_l2(rec.detach(), fx["rec"]) < TOL and O.rel_l2(dec.detach(), fx["dec"]) < TOL
    ((rec * fx["r1"]).sum() + (dec * fx["r2"]).sum()).backward()
    assert O.rel_l2(x.grad, fx["dx"]) < 1e-9
    for k, g in fx["grads"].items():
        assert O.rel_l2(p[k].grad, g) < 1e-9, k


def test_time_conditioned_model_oracle_matches_golden(golden):
    """Time_Input=True (exp_plas.py:148,187): forward(x, fx, T) of the reference 2D model; the embedding and time_fc are fp32
    in the reference whatever the model dtype (model/Embedding.py:81), hence the looser bound on time_fc gradients."""
    from oracle import model as OM
    fx = golden("model_2d_time.pt")
    kw = fx["kwargs"]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state"].items()}
    out = OM.model_forward(fx["x"], fx["fx"], sd, kw["n_layers"], kw["n_head"], grid=(kw["H"], kw["W"]), T=fx["T"])
    assert O.rel_l2(out.detach(), fx["out"]) < TOL
    loss = O.rel_l2_sum(out.reshape(2, -1), fx["y"].reshape(2, -1))
    assert abs(float(loss) - float(fx["loss"])) < 1e-12
    loss.backward()
    for k, g in fx["grads"].items():
        assert O.rel_l2(sd[k].grad.double(), g.double()) < (1e-5 if k.startswith("time_fc") else 1e-8), k


def test_3d_oracle_matches_golden(golden):
    """Physics_Attention_Structured_Mesh_3D (model/Physics_Attention.py:232-288): oracle forward and its autograd gradients"""
    fx = golden("pa_structured3d_small.pt")
    kw = fx["kwargs"]
    p = {k: v.clone().requires_grad_(True) for k, v in fx["state"].items()}
    x = fx["inputs"][0].clone().requires_grad_(True)
    out = O.pa3d_forward(x, p, kw["heads"], (kw["H"], kw["W"], kw["D"]))
    assert O.rel_l2(out.detach(), fx["out"]) < TOL
    out.backward(fx["dout"])
    assert O.rel_l2(x.grad, fx["dinputs"][0]) < 1e-9
    for k, g in fx["grads"].items():
        if float(g.abs().max()) == 0.0:
            assert float(p[k].grad.abs().max()) == 0.0, k
        else:
            assert O.rel_l2(p[k].grad, g) < 1e-9, k
