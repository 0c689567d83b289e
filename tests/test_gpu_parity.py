"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test drives the CUDA kernels through the
C ABI (ctypes -> libtbns.so) and compares with the CPU oracle / the golden vectors produced by the live
reference.  Tolerances are the ones BASELINE.json's north_star states:
    fp32 mode: per-layer relative L2 <= 1e-5      bf16 mode: <= 2e-3
Gradients are held to 1e-4 (fp32) — they pass through long fp32 reductions over all tokens."""
import math

import pytest
import torch

from oracle import physics_attention as O

pytestmark = pytest.mark.gpu

FP32_OUT_TOL = 1e-5
FP32_GRAD_TOL = 1e-4
BF16_OUT_TOL = 2e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from transformerbasednavierstokesolver_b200 import _lib
    assert _lib.load().tbns_device_ok() == 1, _lib.load().tbns_last_error()
    return torch.device("cuda:0")


def _ops():
    from transformerbasednavierstokesolver_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------
# generic GEMM engine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (200, 130, 52), (37, 5, 7), (1, 300, 33), (513, 64, 1), (300, 257, 421), (129, 33, 32)])
@pytest.mark.parametrize("a_kind,b_kind", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_layouts(dev, M, N, K, a_kind, b_kind):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + a_kind * 2 + b_kind)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(K, N, generator=g)
    Ad = (A if a_kind == 0 else A.t()).contiguous().to(dev)
    Bd = (B.t() if b_kind == 0 else B).contiguous().to(dev)
    C = torch.empty(M, N, device=dev)
    ops.gemm(M=M, N=N, K=K, A=Ad, lda=Ad.shape[1], a_kind=a_kind, B=Bd, ldb=Bd.shape[1], b_kind=b_kind, C=C, ldc=N)
    assert O.rel_l2(C.cpu(), A.double() @ B.double()) < 2e-6


def test_gemm_batched_splitk_epilogues(dev):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    Bt, M, N, K = 3, 70, 36, 1000
    A = torch.randn(Bt, M, K, generator=g)
    B = torch.randn(Bt, K, N, generator=g)
    bias = torch.randn(N, generator=g)
    res = torch.randn(Bt, M, N, generator=g)
    C = torch.empty(Bt, M, N, device=dev)
    ops.gemm(M=M, N=N, K=K, A=A.to(dev), lda=K, a_kind=0, B=B.to(dev), ldb=N, b_kind=1, C=C, ldc=N, batch=Bt, sA=M * K, sB=K * N,
             sC=M * N, sR=M * N, bias=bias.to(dev), residual=res.to(dev), ldr=N, split_k=5)
    ref = A.double() @ B.double() + bias.double() + res.double()
    assert O.rel_l2(C.cpu(), ref) < 2e-6
    # GELU epilogue with pre-activation side output, then GELU' multiply
    pre = torch.empty(Bt, M, N, device=dev)
    ops.gemm(M=M, N=N, K=K, A=A.to(dev), lda=K, a_kind=0, B=B.to(dev), ldb=N, b_kind=1, C=C, ldc=N, batch=Bt, sA=M * K, sB=K * N,
             sC=M * N, sAux=M * N, bias=bias.to(dev), act=1, aux_out=pre, ldaux=N)
    pre_ref = (A.double() @ B.double() + bias.double()) / 1.0
    assert O.rel_l2(pre.cpu(), pre_ref) < 2e-6
    assert O.rel_l2(C.cpu(), O.gelu(pre_ref)) < 2e-6
    small = (pre_ref / 30.0).float()
    ops.gemm(M=M, N=N, K=K, A=A.to(dev), lda=K, a_kind=0, B=B.to(dev), ldb=N, b_kind=1, C=C, ldc=N, batch=Bt, sA=M * K, sB=K * N,
             sC=M * N, sAux=M * N, act=2, aux_in=small.to(dev), ldaux=N)
    assert O.rel_l2(C.cpu(), (A.double() @ B.double()) * O.gelu_grad(small.double())) < 2e-6


@pytest.mark.parametrize("Bt,Hg,Wg,C,I2", [(2, 5, 7, 8, 12), (1, 9, 4, 12, 20), (1, 1, 1, 4, 4), (2, 9, 12, 32, 64), (3, 16, 16, 20, 72)])
def test_gemm_conv_modes(dev, Bt, Hg, Wg, C, I2):
    """conv fprop / dgrad / wgrad gathers against the oracle's shifted-matmul restatement"""
    ops = _ops()
    g = torch.Generator().manual_seed(Bt + Hg + Wg)
    I = I2 // 2
    N = Hg * Wg
    x = torch.randn(Bt, N, C, generator=g)
    Wx = torch.randn(I, C, 3, 3, generator=g)
    Wfx = torch.randn(I, C, 3, 3, generator=g)
    bx, bfx = torch.randn(I, generator=g), torch.randn(I, generator=g)
    Wf, Wd, bcat, _, _ = ops.pack_proj_weights(Wx.to(dev), bx.to(dev), Wfx.to(dev), bfx.to(dev))
    XF = torch.empty(Bt * N, I2, device=dev)
    ops.gemm(M=Bt * N, N=I2, K=9 * C, A=x.to(dev), lda=C, a_kind=0, B=Wf, ldb=9 * C, b_kind=0, C=XF, ldc=I2, conv_mode=1, Hg=Hg,
             Wg=Wg, Cin=C, bias=bcat)
    ref = O.proj_fwd(x.double(), Wx.double(), bx.double(), Wfx.double(), bfx.double(), (Hg, Wg))
    assert O.rel_l2(XF.cpu().reshape(Bt, N, I2), ref) < 2e-6
    dXF = torch.randn(Bt, N, I2, generator=g)
    rdx, rdWx, rdbx, rdWfx, rdbfx = O.proj_bwd(dXF.double(), x.double(), Wx.double(), Wfx.double(), (Hg, Wg))
    dx = torch.empty(Bt, N, C, device=dev)
    ops.gemm(M=Bt * N, N=C, K=9 * I2, A=dXF.to(dev), lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=dx, ldc=C, conv_mode=1, Hg=Hg,
             Wg=Wg, Cin=I2, flip=1)
    assert O.rel_l2(dx.cpu(), rdx) < 2e-6
    dWx = torch.empty(I, C, 3, 3, device=dev)
    dWfx = torch.empty(I, C, 3, 3, device=dev)
    for sk in (1, 3):
        ops.gemm(M=9 * C, N=I2, K=Bt * N, A=x.to(dev), lda=C, a_kind=1, B=dXF.to(dev), ldb=I2, b_kind=1, conv_mode=2, Hg=Hg, Wg=Wg,
                 Cin=C, split_k=sk, scatter=(dWx, dWfx), I=I, taps=9)
        assert O.rel_l2(dWx.cpu(), rdWx) < 2e-6
        assert O.rel_l2(dWfx.cpu(), rdWfx) < 2e-6
    db = ops.colsum(dXF.to(dev).reshape(Bt * N, I2), Bt * N, I2)
    assert O.rel_l2(db.cpu(), torch.cat([rdbx, rdbfx])) < 2e-6


@pytest.mark.parametrize("a_kind,b_kind", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_fp32_mode_is_3xtf32_on_tensor_cores(dev, a_kind, b_kind):
    """fp32 mode (TBNS_PREC_FP32) runs the 3xTF32 tcgen05 engine (csrc/gemm_x3.cu): long K (ring wrap, several phases of every
    stage barrier), split-K, both operand orientations.  Its error against fp64 must be fp32-class (a plain TF32 product would
    be ~5e-4), and it must not be bit-identical to the FMA engine (TBNS_PREC_FP32_EXACT) - that would mean it never ran."""
    ops = _ops()
    from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_FP32, TBNS_PREC_FP32_EXACT
    g = torch.Generator().manual_seed(11 + 2 * a_kind + b_kind)
    M, N, K = 384, 256, 1000
    # wide dynamic range: the lo parts matter
    A = torch.randn(M, K, generator=g) * torch.exp(2 * torch.randn(M, K, generator=g))
    B = torch.randn(K, N, generator=g) * torch.exp(2 * torch.randn(K, N, generator=g))
    Ad = (A if a_kind == 0 else A.t()).contiguous().to(dev)
    Bd = (B.t() if b_kind == 0 else B).contiguous().to(dev)
    ref = A.double() @ B.double()
    outs = {}
    for prec in (TBNS_PREC_FP32, TBNS_PREC_FP32_EXACT):
        for sk in (1, 4):
            C = torch.empty(M, N, device=dev)
            ops.gemm(M=M, N=N, K=K, A=Ad, lda=Ad.shape[1], a_kind=a_kind, B=Bd, ldb=Bd.shape[1], b_kind=b_kind, C=C, ldc=N,
                     precision=prec, split_k=sk)
            outs[(prec, sk)] = C.cpu()
            assert O.rel_l2(outs[(prec, sk)], ref) < 1e-6, (prec, sk)
    assert not torch.equal(outs[(TBNS_PREC_FP32, 1)], outs[(TBNS_PREC_FP32_EXACT, 1)])
    # element-wise: no entry is off by more than fp32-class error relative to the magnitude of its dot product
    scale = (A.double().abs() @ B.double().abs())
    assert ((outs[(TBNS_PREC_FP32, 1)].double() - ref).abs() / scale).max() < 2e-6


def test_gemm_bf16_mode_rounds_operands(dev):
    ops = _ops()
    from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_BF16
    g = torch.Generator().manual_seed(9)
    A = torch.randn(64, 96, generator=g)
    B = torch.randn(96, 48, generator=g)
    C = torch.empty(64, 48, device=dev)
    ops.gemm(M=64, N=48, K=96, A=A.to(dev), lda=96, a_kind=0, B=B.to(dev), ldb=48, b_kind=1, C=C, ldc=48, precision=TBNS_PREC_BF16)
    ref = A.bfloat16().double() @ B.bfloat16().double()
    assert O.rel_l2(C.cpu(), ref) < 2e-6


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C", [(1000, 256), (37, 64), (5, 30), (4096, 128), (1001, 256), (77, 512), (1, 128)])
def test_layernorm(dev, rows, C):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    gam, bet = torch.randn(C, generator=g), torch.randn(C, generator=g)
    dy = torch.randn(rows, C, generator=g)
    dres = torch.randn(rows, C, generator=g)
    y, y16, mean, rstd = ops.layernorm_fwd(x.to(dev), gam.to(dev), bet.to(dev), want16=True)
    assert torch.equal(y16, y.bfloat16())
    ry, rmean, rrstd = O.layernorm_fwd(x.double(), gam.double(), bet.double())
    assert O.rel_l2(y.cpu(), ry) < 2e-6
    dx, dx16, dg, db = ops.layernorm_bwd(dy.to(dev), x.to(dev), mean, rstd, gam.to(dev), dres.to(dev), want16=True)
    assert torch.equal(dx16, dx.bfloat16())
    rdx, rdg, rdb = O.layernorm_bwd(dy.double(), x.double(), rmean, rrstd, gam.double())
    assert O.rel_l2(dx.cpu(), rdx + dres.double()) < 5e-6
    assert O.rel_l2(dg.cpu(), rdg) < 5e-6
    assert O.rel_l2(db.cpu(), rdb) < 5e-6


@pytest.mark.parametrize("rows,C,Od", [(5000, 256, 1), (777, 128, 1), (300, 512, 1), (1000, 256, 3), (50, 96, 1)])
def test_last_layer_ln_linear(dev, rows, C, Od):
    """mlp2(ln_3(fx)): the fused out_dim=1 kernels (C in {128,256,512}) and the generic route, against the oracle
    (model/Transolver_Structured_Mesh_2D.py:72-73)."""
    ops = _ops()
    from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_FP32
    g = torch.Generator().manual_seed(rows + C + Od)
    fx = (torch.randn(2, rows // 2, C, generator=g) * 1.5 + 0.3)
    gam, bet = torch.randn(C, generator=g), torch.randn(C, generator=g)
    W, b = torch.randn(Od, C, generator=g) / C ** 0.5, torch.randn(Od, generator=g)
    dout = torch.randn(2, rows // 2, Od, generator=g)
    t = [v.to(dev).requires_grad_(True) for v in (fx, gam, bet, W, b)]
    out = ops.LnLinearFn.apply(*t, 1e-5, TBNS_PREC_FP32)
    out.backward(dout.to(dev))
    rout, sv = O.ln_linear_fwd(fx.double(), gam.double(), bet.double(), W.double(), b.double())
    rdfx, rg = O.ln_linear_bwd(dout.double(), gam.double(), W.double(), sv)
    assert O.rel_l2(out.detach().cpu(), rout) < 1e-5
    for got, want in zip([v.grad for v in t], [rdfx, rg["ln_w"], rg["ln_b"], rg["W"], rg["b"]]):
        assert got.shape == want.shape
        assert O.rel_l2(got.cpu(), want) < 1e-5


# ------------------------------------------------------------------------------------------------
# attention module / block / model against the golden vectors of the live reference
# ------------------------------------------------------------------------------------------------
def _build_pa(fx, dev, precision):
    from transformerbasednavierstokesolver_b200.model import Physics_Attention as PA
    cls = getattr(PA, fx["kind"])
    m = cls(**fx["kwargs"])
    m.load_state_dict({k: v.float() for k, v in fx["state"].items()}, strict=True)
    m.precision = precision
    return m.to(dev)


PA_FIXTURES = ["pa_structured_small.pt", "pa_structured_g64.pt", "pa_irregular_small.pt", "pa_irregular_inner_ne_dim.pt",
               "pa_ckpt_ep400_block3.pt"]


@pytest.mark.parametrize("name", PA_FIXTURES)
def test_pa_module_fp32_matches_reference_golden(dev, golden, name):
    fx = golden(name)
    m = _build_pa(fx, dev, "fp32")
    x = fx["inputs"][0].float().to(dev).requires_grad_(True)
    out = m(x)
    out.backward(fx["dout"].float().to(dev))
    assert O.rel_l2(out.cpu(), fx["out"]) < FP32_OUT_TOL
    assert O.rel_l2(x.grad.cpu(), fx["dinputs"][0]) < FP32_GRAD_TOL
    for k, p in m.named_parameters():
        ref = fx["grads"][k]
        if float(ref.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, k   # clamped temperature heads
        else:
            assert O.rel_l2(p.grad.cpu(), ref) < FP32_GRAD_TOL, k


@pytest.mark.parametrize("name", PA_FIXTURES)
def test_pa_module_bf16_mode(dev, golden, name):
    """bf16 mode on shared bf16-representable weights and inputs (SURVEY.md §7 hard part 1)"""
    fx = golden(name)
    state = {k: v.float().bfloat16().double() for k, v in fx["state"].items()}
    x = fx["inputs"][0].float().bfloat16().double()
    kw = fx["kwargs"]
    ref, _ = O.pa_forward(x, state, kw["heads"], (kw["H"], kw["W"]) if "H" in kw else None)
    fx2 = dict(fx, state=state)
    m = _build_pa(fx2, dev, "bf16")
    out = m(x.float().to(dev))
    assert O.rel_l2(out.cpu(), ref) < BF16_OUT_TOL


@pytest.mark.parametrize("name", ["block_structured_mid.pt", "block_structured_last.pt", "block_irregular_last.pt"])
def test_block_fp32_matches_reference_golden(dev, golden, name):
    fx = golden(name)
    if fx["kind"] == "block_structured":
        from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Transolver_block
    else:
        from transformerbasednavierstokesolver_b200.model.Transolver_Irregular_Mesh import Transolver_block
    m = Transolver_block(**fx["kwargs"])
    m.load_state_dict({k: v.float() for k, v in fx["state"].items()}, strict=True)
    m.Attn.precision = "fp32"
    m = m.to(dev)
    x = fx["inputs"][0].float().to(dev).requires_grad_(True)
    out = m(x)
    out.backward(fx["dout"].float().to(dev))
    assert O.rel_l2(out.cpu(), fx["out"]) < FP32_OUT_TOL
    assert O.rel_l2(x.grad.cpu(), fx["dinputs"][0]) < FP32_GRAD_TOL
    for k, p in m.named_parameters():
        assert O.rel_l2(p.grad.cpu(), fx["grads"][k]) < FP32_GRAD_TOL, k


@pytest.mark.parametrize("name", ["model_2d_unified.pt", "model_2d_plainpos.pt", "model_irregular.pt", "model_2d_time.pt"])
def test_model_fp32_matches_reference_golden(dev, golden, name):
    import transformerbasednavierstokesolver_b200 as pkg
    fx = golden(name)
    pkg.set_default_precision("fp32")
    try:
        if fx["kind"] == "model_2d":
            from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
        else:
            from transformerbasednavierstokesolver_b200.model.Transolver_Irregular_Mesh import Model
        m = Model(**fx["kwargs"])
        m.load_state_dict({k: v.float() for k, v in fx["state"].items()}, strict=True)
        m = m.to(dev)
        x = fx["x"].float().to(dev)
        f = fx["fx"].float().to(dev) if "fx" in fx else None
        y = fx["y"].float().to(dev)
        out = m(x, f, T=fx["T"].float().to(dev)) if "T" in fx else m(x, f)   # Time_Input=True fixture: forward(x, fx, T)
        assert O.rel_l2(out.cpu(), fx["out"]) < FP32_OUT_TOL
        n = out.shape[0]
        loss = (torch.linalg.vector_norm(out.reshape(n, -1) - y.reshape(n, -1), dim=1) / torch.linalg.vector_norm(y.reshape(n, -1), dim=1)).sum()
        assert abs(float(loss.detach()) - float(fx["loss"])) < 1e-5 * abs(float(fx["loss"]))
        loss.backward()
        # trunc_normal(0.02) init makes the slice softmax near-uniform: the q/k/slice/temperature gradients are ~1e-10
        # differences of O(1) terms, i.e. fp32 cancellation noise (the tight per-module gradient checks are above, on
        # sharpened slices) -> hold the well-conditioned ones to 1e-3 and the cancellation-dominated ones to 5e-2
        for k, p in m.named_parameters():
            if k in fx["grads"]:
                loose = k.endswith(("temperature", "to_q.weight", "to_k.weight", "in_project_slice.weight", "in_project_slice.bias"))
                assert O.rel_l2(p.grad.cpu(), fx["grads"][k]) < (5e-2 if loose else 1e-3), k
        if "rollout" in fx:  # closed-loop autoregressive rollout (exp_ns.py:225-241)
            with torch.no_grad():
                ff = f.clone()
                preds = []
                for _ in range(fx["rollout"].shape[-1]):
                    im = m(x, fx=ff)
                    preds.append(im)
                    ff = torch.cat((ff[..., 1:], im), -1)
            roll = torch.cat(preds, -1)
            assert O.rel_l2(roll.cpu(), fx["rollout"]) < 5e-5
            e_ref = float(O.rel_l2_sum(fx["rollout"], fx["y"].expand_as(fx["rollout"])))
            e_new = float(O.rel_l2_sum(roll.cpu().double(), fx["y"].expand_as(fx["rollout"])))
            assert abs(e_new - e_ref) < 1e-5 * abs(e_ref)   # "unchanged rollout error" (fp32 mode)
    finally:
        pkg.set_default_precision("bf16")


# ------------------------------------------------------------------------------------------------
# BASELINE configuration sizes against the oracle (CPU, seconds)
# ------------------------------------------------------------------------------------------------
def _rand_state(cls, kwargs, seed):
    torch.manual_seed(seed)
    m = cls(**kwargs)
    with torch.no_grad():
        m.in_project_slice.weight.mul_(4.0)   # sharpen the slice softmax (random init is near-uniform)
        m.temperature.copy_(torch.linspace(0.2, 1.5, kwargs["heads"]).reshape(1, -1, 1, 1))
    return m


@pytest.mark.parametrize("cfg", ["cfg1_ns64", "cfg3_darcy85", "cfg4_elas972"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pa_baseline_config_sizes(dev, cfg, precision):
    from transformerbasednavierstokesolver_b200.model import Physics_Attention as PA
    if cfg == "cfg1_ns64":
        cls, kw, B, N = PA.Physics_Attention_Structured_Mesh_2D, dict(dim=256, heads=8, dim_head=32, slice_num=32, H=64, W=64), 2, 4096
    elif cfg == "cfg3_darcy85":
        cls, kw, B, N = PA.Physics_Attention_Structured_Mesh_2D, dict(dim=128, heads=8, dim_head=16, slice_num=64, H=85, W=85), 4, 7225   # SURVEY §8d cfg 3: batch 4
    else:
        cls, kw, B, N = PA.Physics_Attention_Irregular_Mesh, dict(dim=128, heads=8, dim_head=16, slice_num=64), 2, 972
    m = _rand_state(cls, kw, 7)
    x = torch.nn.functional.layer_norm(torch.randn(B, N, kw["dim"]), (kw["dim"],))
    if precision == "bf16":
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(p.bfloat16().float())
        x = x.bfloat16().float()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    grid = (kw["H"], kw["W"]) if "H" in kw else None
    ref, sv = O.pa_forward(x, sd, kw["heads"], grid)          # fp32 oracle on CPU
    m.precision = precision
    m = m.to(dev)
    xd = x.to(dev).requires_grad_(True)
    out = m(xd)
    tol = FP32_OUT_TOL if precision == "fp32" else BF16_OUT_TOL
    assert O.rel_l2(out.cpu(), ref) < tol
    # size-independent properties: slice weights are a partition of unity; sum_g s_g == N per (b,h)
    dout = torch.randn_like(ref)
    out.backward(dout.to(dev))
    rdx, rg = O.pa_backward(dout, sd, sv)
    # ALL 13 gradients (SURVEY §8 a-bwd), both modes.  bf16 mode: GEMM weights / biases see two to three bf16 operand roundings
    # of long sums (gate 1e-2); everything downstream of the slice-softmax backward is a sum of large cancelling terms
    # (gate 5e-2; tests/test_gpu_bench_path.py states the measured values); fp32 mode: 2e-4 / 5e-3.
    cancel = ("temperature", "in_project_slice.weight", "in_project_slice.bias", "to_q.weight", "to_k.weight",
              "in_project_x.weight", "in_project_x.bias")   # everything downstream of the slice-softmax backward
    assert O.rel_l2(xd.grad.cpu(), rdx) < (2e-4 if precision == "fp32" else 1e-2)
    got = dict(m.named_parameters())
    assert set(got) == set(rg)
    errs = {k: O.rel_l2(got[k].grad.cpu(), rg[k]) for k in rg}
    print(cfg, precision, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        if precision == "fp32":
            assert e < (5e-3 if k in cancel else 2e-4), (k, e)   # oracle itself runs in fp32 here
        else:
            assert e < (5e-2 if k in cancel else 1e-2), (k, e)


def test_slice_weights_partition_of_unity_full_size(dev):
    """property test at cfg-5 width (N = 65 536 tokens): rows of w sum to 1, sum_g s = N, Tt = w^T F"""
    ops = _ops()
    from transformerbasednavierstokesolver_b200 import _lib
    lib = _lib.load()
    B, N, H, D, G = 1, 65536, 8, 32, 64
    g = torch.Generator(device="cpu").manual_seed(1)
    XF = torch.randn(B * N, 2 * H * D, generator=g).to(dev)
    Ws = (torch.randn(G, D, generator=g) * 0.5).to(dev)
    bs = torch.randn(G, generator=g).to(dev)
    tau = torch.linspace(0.05, 6.0, H).to(dev)
    nchunk = lib.tbns_slice_groups(B, N, H)
    w = torch.empty(B, N, H * G, device=dev)
    w16 = torch.empty(B, N, H * G, device=dev, dtype=torch.bfloat16)
    part = torch.empty(B * H * nchunk * G * (D + 1), device=dev)
    _lib.check(lib.tbns_pa_slice_fwd(XF.data_ptr(), Ws.data_ptr(), bs.data_ptr(), tau.data_ptr(), w.data_ptr(), w16.data_ptr(),
                                     part.data_ptr(), B, N, H, D, G, 1, torch.cuda.current_stream().cuda_stream), "slice_fwd")
    assert torch.equal(w16, w.bfloat16())
    w4 = w.view(B, N, H, G)
    assert float((w4.sum(-1) - 1).abs().max()) < 1e-5
    p = part.view(B, H, nchunk, G, D + 1).sum(2)
    assert float((p[..., D].sum(-1) - N).abs().max()) < 0.05
    F = XF[:, H * D:].view(B, N, H, D)
    Tt = torch.einsum("bnhg,bnhd->bhgd", w4.double(), F.double())
    assert O.rel_l2(p[..., :D].cpu(), Tt.cpu()) < 1e-5
    # clamp: heads with tau outside [0.1, 5] behave like tau at the bound
    L = (XF[:, :H * D].view(B, N, H, D).double() @ Ws.double().t() + bs.double()) / tau.double().clamp(0.1, 5.0)[None, None, :, None]
    assert O.rel_l2(w4.cpu(), torch.softmax(L, -1).cpu()) < 1e-5


def test_cfg5_rollout_shape_block_forward_bf16(dev):
    """BASELINE cfg 5 shape (256x256 grid, C=256, 8 heads, slice_num 64): one Transolver block forward in bf16 mode against
    the fp32 oracle on bf16-representable weights/inputs, plus the closed-loop rollout plumbing of train.rollout."""
    from transformerbasednavierstokesolver_b200 import train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model, Transolver_block
    torch.manual_seed(11)
    blk = Transolver_block(num_heads=8, hidden_dim=256, dropout=0.0, mlp_ratio=1, last_layer=False, slice_num=64, H=256, W=256)
    with torch.no_grad():
        blk.Attn.in_project_slice.weight.mul_(4.0)
        for p in blk.parameters():
            p.copy_(p.bfloat16().float())
    fx = torch.randn(1, 65536, 256).bfloat16().float()
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    ref, _ = O.block_forward(fx, sd, 8, (256, 256))
    blk.Attn.precision = "bf16"
    blk = blk.to(dev)
    with torch.no_grad():
        out = blk(fx.to(dev))
    assert O.rel_l2(out.cpu(), ref) < BF16_OUT_TOL
    # rollout plumbing on a small model at the same slice_num / head shape: window shift feeds predictions back
    torch.manual_seed(12)
    m = Model(space_dim=2, n_layers=2, n_hidden=256, n_head=8, fun_dim=10, out_dim=1, slice_num=64, ref=8, unified_pos=1, H=32, W=32).to(dev)
    x, f, _ = train.synthetic_ns_batch(2, 32, 10, 10, seed=5, device=dev)
    roll = train.rollout(m, x, f, T=4, step=1)
    assert roll.shape == (2, 1024, 4) and bool(torch.isfinite(roll).all())
    with torch.no_grad():
        first = m(x, fx=f)
    assert torch.allclose(first[..., 0], roll[..., 0], atol=1e-5)


def test_cfg1_ten_step_rollout_bf16_error_unchanged(dev):
    """BASELINE cfg 1 model (NS 64x64, 8 layers, n_hidden 256, 8 heads, slice_num 32, unified_pos) in bf16 mode: the 10-step
    closed-loop rollout (exp_ns.py:225-241) against the oracle run in fp32 on the SAME random-init weights, and the
    accumulated rollout error against a synthetic target ("unchanged 10-step autoregressive rollout error", north_star)."""
    import transformerbasednavierstokesolver_b200 as pkg
    from oracle import model as OM
    from transformerbasednavierstokesolver_b200 import train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
    torch.manual_seed(21)
    pkg.set_default_precision("bf16")
    m = Model(space_dim=2, n_layers=8, n_hidden=256, n_head=8, fun_dim=10, out_dim=1, slice_num=32, ref=8, unified_pos=1, H=64, W=64,
              mlp_ratio=1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x, f, yy = train.synthetic_ns_batch(1, 64, 10, 10, seed=9)
    with torch.no_grad():
        ref = OM.rollout(x, f, lambda a, b: OM.model_forward(a, b, sd, 8, 8, grid=(64, 64), unified_pos=True, ref=8), T=10)
    m = m.to(dev)
    roll = train.rollout(m, x.to(dev), f.to(dev), T=10, step=1).cpu()
    assert roll.shape == ref.shape == (1, 4096, 10)
    # weights here are NOT bf16-representable: every tensor-core operand rounding (per-layer budget 2e-3) counts against
    # the fp32 reference, over 8 layers per call and 10 chained calls
    r_all, r_first = O.rel_l2(roll, ref), O.rel_l2(roll[..., :1], ref[..., :1])
    e_ref = float(O.rel_l2_sum(ref.reshape(1, -1), yy.reshape(1, -1)))
    e_new = float(O.rel_l2_sum(roll.reshape(1, -1), yy.reshape(1, -1)))
    print(f"cfg1 bf16 rollout: rel-L2 first call {r_first:.2e}, ten steps {r_all:.2e}; rollout error {e_new:.5f} vs {e_ref:.5f}")
    assert r_first < 8 * 2e-3 / 2      # 8 layers, errors add incoherently: well inside 8 x the per-layer gate
    assert r_all < 2e-2
    assert abs(e_new - e_ref) < 5e-3 * abs(e_ref)   # "unchanged rollout error"


@pytest.mark.parametrize("name", ["pa_autoencoder_small.pt", "pa_autoencoder_head1.pt"])
def test_autoencoder_attention_matches_reference_golden(dev, golden, name):
    """Physics_Attention_Structured_Mesh_2D_Auto_Encoder (model/Physics_Attention.py:122-227) in fp32 mode: forward, encode
    (code + cached slice weights), reconstruct_fx (cache replaced by project_slice(cache)), decode on the replaced cache -
    the call order of Transolver_Encoder_block.decode - and every gradient of sum(rec*r1) + sum(dec*r2)."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200.model import Physics_Attention as PA
    fx = golden(name)
    m = PA.Physics_Attention_Structured_Mesh_2D_Auto_Encoder(**fx["kwargs"])
    m.load_state_dict({k: v.float() for k, v in fx["state"].items()})
    m.precision = "fp32"
    m = m.to(dev)
    x = fx["x"].float().to(dev).requires_grad_(True)
    assert O.rel_l2(m(x).detach().cpu(), fx["fwd"]) < FP32_OUT_TOL
    code = m.encode(x, cache_slice=True)
    assert O.rel_l2(code.detach().cpu(), fx["code"]) < FP32_OUT_TOL
    assert O.rel_l2(m.slice_weights.detach().cpu(), fx["w_enc"]) < FP32_OUT_TOL
    rec = m.reconstruct_fx(code)
    assert O.rel_l2(m.slice_weights.detach().cpu(), fx["w_proj"]) < FP32_OUT_TOL
    dec = m.decode(code)
    assert O.rel_l2(rec.detach().cpu(), fx["rec"]) < FP32_OUT_TOL and O.rel_l2(dec.detach().cpu(), fx["dec"]) < FP32_OUT_TOL
    ((rec * fx["r1"].float().to(dev)).sum() + (dec * fx["r2"].float().to(dev)).sum()).backward()
    assert O.rel_l2(x.grad.cpu(), fx["dx"]) < 1e-4
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert O.rel_l2(p.grad.cpu(), fx["grads"][k]) < 1e-4, k


def test_autoencoder_model_encode_decode_bf16_vs_oracle(dev):
    """Transolver_Structured_Mesh2D_Encoder.Model at a tensor-core shape (n_hidden 128, 4 heads of 32, slice_num 32): decode(encode)
    == forward, and the last block's encode / decode in bf16 mode against the fp32 oracle on bf16-representable weights."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh2D_Encoder as E
    torch.manual_seed(31)
    pkg.set_default_precision("bf16")
    model = E.Model(space_dim=2, n_layers=2, n_hidden=128, n_head=4, fun_dim=1, out_dim=1, slice_num=32, ref=4, unified_pos=1, H=16, W=16)
    with torch.no_grad():
        for p_ in model.parameters():
            p_.copy_(p_.bfloat16().float())
    blk = model.blocks[-1]
    sd = {k: v.detach().clone().double() for k, v in blk.state_dict().items()}
    asd = {k[len("Attn."):]: v for k, v in sd.items() if k.startswith("Attn.")}
    h = torch.randn(2, 256, 128).bfloat16().float()
    x1, _, _ = O.layernorm_fwd(h.double(), sd["ln_1.weight"], sd["ln_1.bias"])
    code_ref, w_ref = O.ae_encode(x1, asd, 4, (16, 16))
    wp = O.ae_project_slice(w_ref, asd)
    y = O.ae_decode(code_ref, wp, asd) + O.ae_decode(code_ref, wp, asd)
    model = model.to(dev)
    code = blk.encode(h.to(dev))
    assert O.rel_l2(code.detach().cpu(), code_ref) < 3e-3
    out = blk.decode(code)
    y2, _ = O.ln_mlp_fwd(y, sd["ln_2.weight"], sd["ln_2.bias"], sd["mlp.linear_pre.0.weight"], sd["mlp.linear_pre.0.bias"],
                         sd["mlp.linear_post.weight"], sd["mlp.linear_post.bias"])
    ref_out, _ = O.ln_linear_fwd(y2, sd["ln_3.weight"], sd["ln_3.bias"], sd["mlp2.weight"], sd["mlp2.bias"])
    assert O.rel_l2(out.detach().cpu(), ref_out) < 5e-3
    from transformerbasednavierstokesolver_b200 import train
    x, f, _ = train.synthetic_ns_batch(2, 16, 1, 1, seed=4, device=dev)
    with torch.no_grad():
        a = model(x, f)
        b = model.decode(model.encode(x, f))
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert tuple(model.get_attention_slice().shape) == (2, 4, 256, 32)
    out.sum().backward()   # gradients flow through decode, the projected cache and encode
    assert all(p_.grad is not None and bool(torch.isfinite(p_.grad).all()) for p_ in blk.parameters())


def test_3d_attention_matches_reference_golden(dev, golden):
    """Physics_Attention_Structured_Mesh_3D (model/Physics_Attention.py:232-288) in fp32 mode against the live-reference golden:
    output, input gradient and every parameter gradient (Conv3d weights come back in [out, in, 3, 3, 3] layout)."""
    from transformerbasednavierstokesolver_b200.model import Physics_Attention as PA
    fx = golden("pa_structured3d_small.pt")
    m = PA.Physics_Attention_Structured_Mesh_3D(**fx["kwargs"])
    m.load_state_dict({k: v.float() for k, v in fx["state"].items()})
    m.precision = "fp32"
    m = m.to(dev)
    x = fx["inputs"][0].float().to(dev).requires_grad_(True)
    out = m(x)
    assert O.rel_l2(out.detach().cpu(), fx["out"]) < FP32_OUT_TOL
    out.backward(fx["dout"].float().to(dev))
    assert O.rel_l2(x.grad.cpu(), fx["dinputs"][0]) < 1e-4
    for k, p in m.named_parameters():
        g = fx["grads"][k]
        if float(g.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert O.rel_l2(p.grad.cpu(), g) < 1e-4, k


def test_3d_model_bf16_forward_backward(dev):
    """Transolver_Structured_Mesh_3D.Model at a tensor-core shape (n_hidden 128, 4 heads of 32, 8x8x8 mesh): bf16 mode against the
    fp32 mode of the same weights, gradients finite."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh_3D as M3
    torch.manual_seed(41)
    try:
        model = M3.Model(space_dim=3, n_layers=2, n_hidden=128, n_head=4, fun_dim=2, out_dim=1, slice_num=32, ref=2, unified_pos=1,
                         H=8, W=8, D=8).to(dev)
        x = torch.rand(2, 512, 3, device=dev)
        f = torch.randn(2, 512, 2, device=dev)
        pkg.set_default_precision("fp32")
        with torch.no_grad():
            ref = model(x, f)
        pkg.set_default_precision("bf16")
        out = model(x, f)
        assert O.rel_l2(out.detach().cpu(), ref.cpu()) < 1e-2
        out.square().sum().backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for n, p in model.named_parameters() if n != "placeholder")
    finally:
        pkg.set_default_precision("bf16")
