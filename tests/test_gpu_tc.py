"""GPU tests of the tcgen05/TMA/TMEM implicit-GEMM kernel (tbns_gemm_tc) against the already-verified fp32 SIMT
engine run in bf16-operand mode (identical operand rounding, so only the fp32 summation order differs) and against
the CPU oracle on bf16-rounded inputs."""
import pytest
import torch

from oracle import physics_attention as O

pytestmark = pytest.mark.gpu


def _bf16(t):
    from transformerbasednavierstokesolver_b200 import _lib
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tbns_cast_bf16(t.data_ptr(), out.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream), "cast")
    return out


def _tc(A16, W16, C, bias, Bimg, Hg, Wg, Cin, N, taps, flip, **kw):
    from transformerbasednavierstokesolver_b200 import ops
    ops.gemm_tc(A16, W16, C, bias, Bimg, Hg, Wg, Cin, N, taps, flip, **kw)


def test_cast_bf16_matches_torch():
    dev = torch.device("cuda:0")
    x = torch.randn(1000003, device=dev)[:1000000 - 8].contiguous()  # odd length exercises the tail
    x = torch.randn(999997 + 3, device=dev)
    y = _bf16(x)
    assert torch.equal(y, x.bfloat16())


@pytest.mark.parametrize("Bimg,Hg,Wg,C,I2", [(2, 64, 64, 256, 512), (1, 85, 85, 128, 256), (1, 12, 10, 64, 128), (3, 7, 33, 64, 64)])
def test_tc_conv_fprop_and_dgrad(Bimg, Hg, Wg, C, I2):
    from transformerbasednavierstokesolver_b200 import ops
    from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_BF16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Hg * Wg + C)
    I = I2 // 2
    N = Hg * Wg
    x = torch.randn(Bimg, N, C, generator=g).to(dev)
    Wx = (torch.randn(I, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(dev)
    Wfx = (torch.randn(I, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(dev)
    bx, bfx = torch.randn(I, generator=g).to(dev), torch.randn(I, generator=g).to(dev)
    Wf, Wd, bcat, _, _ = ops.pack_proj_weights(Wx, bx, Wfx, bfx)
    # fprop
    ref = torch.empty(Bimg * N, I2, device=dev)
    ops.gemm(M=Bimg * N, N=I2, K=9 * C, A=x, lda=C, a_kind=0, B=Wf, ldb=9 * C, b_kind=0, C=ref, ldc=I2, conv_mode=1, Hg=Hg, Wg=Wg,
             Cin=C, bias=bcat, precision=TBNS_PREC_BF16)
    out = torch.full((Bimg * N, I2), float("nan"), device=dev)
    _tc(_bf16(x), _bf16(Wf), out, bcat, Bimg, Hg, Wg, C, I2, 9, 0)
    torch.cuda.synchronize()
    assert O.rel_l2(out.cpu(), ref.cpu()) < 1e-5
    cpu = O.proj_fwd(x.cpu().bfloat16().double(), Wx.cpu().bfloat16().double(), bx.cpu().double(), Wfx.cpu().bfloat16().double(),
                     bfx.cpu().double(), (Hg, Wg))
    assert O.rel_l2(out.cpu().reshape(Bimg, N, I2), cpu) < 1e-5
    # dgrad (transposed convolution): A = dXF [.., 2I], W = Wd [C, 9*2I]
    if C % 64 == 0 and I2 % 64 == 0:
        dXF = torch.randn(Bimg, N, I2, generator=g).to(dev)
        ref2 = torch.empty(Bimg, N, C, device=dev)
        ops.gemm(M=Bimg * N, N=C, K=9 * I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=ref2, ldc=C, conv_mode=1, Hg=Hg,
                 Wg=Wg, Cin=I2, flip=1, precision=TBNS_PREC_BF16)
        out2 = torch.full((Bimg, N, C), float("nan"), device=dev)
        _tc(_bf16(dXF), _bf16(Wd), out2, None, Bimg, Hg, Wg, I2, C, 9, 1)
        torch.cuda.synchronize()
        assert O.rel_l2(out2.cpu(), ref2.cpu()) < 1e-5


@pytest.mark.parametrize("M,K,N", [(972, 128, 256), (128, 64, 64), (5000, 256, 512), (81920, 256, 256)])
def test_tc_linear(M, K, N):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.full((M, N), float("nan"), device=dev)
    _tc(_bf16(A), _bf16(W), out, bias, 1, 1, M, K, N, 1, 0)
    torch.cuda.synchronize()
    ref = A.bfloat16().double() @ W.bfloat16().double().t() + bias.double()
    assert O.rel_l2(out.cpu(), ref.cpu()) < 1e-5


def test_tc_epilogues_and_batched_weights():
    """bias + GELU (+pre side output) / GELU' multiply / residual / bf16 output / per-image weight matrices"""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    Bimg, Ntok, K, N = 3, 200, 128, 192
    A = torch.randn(Bimg, Ntok, K, generator=g).to(dev)
    W = (torch.randn(Bimg, N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(Bimg, Ntok, N, generator=g).to(dev)
    A64, W64 = A.bfloat16().double(), W.bfloat16().double()
    base = torch.einsum("bmk,bnk->bmn", A64, W64)
    # batched weights + bias + residual, fp32 and bf16 outputs
    out = torch.full((Bimg, Ntok, N), float("nan"), device=dev)
    out16 = torch.empty(Bimg, Ntok, N, device=dev, dtype=torch.bfloat16)
    _tc(_bf16(A), _bf16(W), out, bias, Bimg, 1, Ntok, K, N, 1, 0, C16=out16, w_batched=1, residual=res)
    ref = base + bias.double() + res.double()
    assert O.rel_l2(out.cpu(), ref.cpu()) < 1e-5
    assert torch.equal(out16, out.bfloat16())
    # GELU with pre-activation side output (shared weights = image 0's matrix)
    pre = torch.empty(Bimg, Ntok, N, device=dev)
    _tc(_bf16(A), _bf16(W[0].contiguous()), out, bias, Bimg, 1, Ntok, K, N, 1, 0, act=1, aux_out=pre)
    pre_ref = torch.einsum("bmk,nk->bmn", A64, W64[0]) + bias.double()
    assert O.rel_l2(pre.cpu(), pre_ref.cpu()) < 1e-5
    assert O.rel_l2(out.cpu(), O.gelu(pre_ref).cpu()) < 1e-5
    # GELU' multiply
    aux = (pre_ref / 3).float()
    _tc(_bf16(A), _bf16(W[0].contiguous()), out, None, Bimg, 1, Ntok, K, N, 1, 0, act=2, aux_in=aux)
    assert O.rel_l2(out.cpu(), (torch.einsum("bmk,nk->bmn", A64, W64[0]) * O.gelu_grad(aux.double())).cpu()) < 1e-5
    # act 3: GELU with the DERIVATIVE GELU'(pre) as side output (bf16); act 4: multiply by that stored derivative
    gp16 = torch.empty(Bimg, Ntok, N, device=dev, dtype=torch.bfloat16)
    h16 = torch.empty(Bimg, Ntok, N, device=dev, dtype=torch.bfloat16)
    _tc(_bf16(A), _bf16(W[0].contiguous()), None, bias, Bimg, 1, Ntok, K, N, 1, 0, act=3, aux_out=gp16, aux_bf16=1, C16=h16)
    assert O.rel_l2(gp16.double().cpu(), O.gelu_grad(pre_ref).cpu()) < 3e-3      # bf16 rounding of the stored values
    assert O.rel_l2(h16.double().cpu(), O.gelu(pre_ref).cpu()) < 3e-3
    _tc(_bf16(A), _bf16(W[0].contiguous()), out, None, Bimg, 1, Ntok, K, N, 1, 0, act=4, aux_in=gp16, aux_bf16=1)
    assert O.rel_l2(out.cpu(), (torch.einsum("bmk,nk->bmn", A64, W64[0]) * gp16.double()).cpu()) < 1e-5


@pytest.mark.parametrize("Bimg,Hg,Wg,C,I2", [(2, 64, 64, 256, 512), (1, 85, 85, 128, 256), (3, 7, 33, 128, 64)])
def test_tc_conv_wgrad(Bimg, Hg, Wg, C, I2):
    """conv weight gradient on tensor cores (MN-major operands) vs the SIMT engine in bf16-operand mode"""
    from transformerbasednavierstokesolver_b200 import ops
    from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_BF16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Hg + Wg + C)
    I = I2 // 2
    N = Hg * Wg
    x = torch.randn(Bimg, N, C, generator=g).to(dev)
    dXF = torch.randn(Bimg, N, I2, generator=g).to(dev)
    rWx, rWfx = torch.empty(I, C, 3, 3, device=dev), torch.empty(I, C, 3, 3, device=dev)
    ops.gemm(M=9 * C, N=I2, K=Bimg * N, A=x, lda=C, a_kind=1, B=dXF, ldb=I2, b_kind=1, conv_mode=2, Hg=Hg, Wg=Wg, Cin=C,
             precision=TBNS_PREC_BF16, split_k=4, scatter=(rWx, rWfx), I=I, taps=9)
    dWx = torch.full((I, C, 3, 3), float("nan"), device=dev)
    dWfx = torch.full((I, C, 3, 3), float("nan"), device=dev)
    ops.gemm_tc_wgrad(_bf16(x), _bf16(dXF), Bimg, Hg, Wg, C, I2, taps=9, scatter=(dWx, dWfx), I=I)
    torch.cuda.synchronize()
    assert O.rel_l2(dWx.cpu(), rWx.cpu()) < 1e-5
    assert O.rel_l2(dWfx.cpu(), rWfx.cpu()) < 1e-5


@pytest.mark.parametrize("Bimg,Ntok,Ma,Nb,batched", [(1, 5000, 256, 256, 0), (4, 972, 512, 128, 1), (2, 4096, 256, 256, 1), (20, 4096, 128, 64, 0)])
def test_tc_plain_wgrad(Bimg, Ntok, Ma, Nb, batched):
    """D = A^T B over tokens (MLP weight gradients; batched = dP = w^T dOut per image)"""
    from transformerbasednavierstokesolver_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Ntok + Ma)
    A = torch.randn(Bimg, Ntok, Ma, generator=g).to(dev)
    Bm = torch.randn(Bimg, Ntok, Nb, generator=g).to(dev)
    batch = Bimg if batched else 1
    out = torch.full((batch, Ma, Nb), float("nan"), device=dev)
    ops.gemm_tc_wgrad(_bf16(A), _bf16(Bm), Bimg, 1, Ntok, Ma, Nb, taps=1, batched=batched, C=out)
    torch.cuda.synchronize()
    A64, B64 = A.bfloat16().double(), Bm.bfloat16().double()
    ref = torch.einsum("bta,btn->ban", A64, B64)
    if not batched:
        ref = ref.sum(0, keepdim=True)
    assert O.rel_l2(out.cpu(), ref.cpu()) < 1e-5


@pytest.mark.parametrize("B,N,H,G", [(2, 4096, 8, 32), (1, 1000, 4, 64), (3, 130, 2, 32), (20, 4096, 8, 32)])
def test_tc_slice_fwd_matches_oracle(B, N, H, G):
    """tensor-core slice forward (bf16 XF through TMA, tcgen05 kind::f16) vs the fp64 oracle stage
    (oracle.slice_fwd <- model/Physics_Attention.py:98-101) on the same bf16-representable inputs: slice weights, their
    sums and the un-normalised slice tokens"""
    from transformerbasednavierstokesolver_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    D = 32
    g = torch.Generator().manual_seed(B * N + G)
    XF16 = torch.randn(B * N, 2 * H * D, generator=g).bfloat16()
    Ws = (torch.randn(G, D, generator=g) * 0.4).bfloat16().float()
    bs = torch.randn(G, generator=g)
    tau = torch.linspace(0.05, 1.5, H)          # the first head sits below the clamp (0.1)
    w_ref, s_ref, Tt_ref = O.slice_fwd(XF16.double().view(B, N, -1), Ws.double(), bs.double(), tau.double(), H, True)
    groups = lib.tbns_slice_groups(B, N, H)
    st = torch.cuda.current_stream().cuda_stream
    XF16, Ws, bs, tau = XF16.to(dev), Ws.to(dev), bs.to(dev), tau.to(dev)
    w16 = torch.full((B, N, H * G), float("nan"), device=dev, dtype=torch.bfloat16)
    part = torch.full((B * H * groups * G * (D + 1),), float("nan"), device=dev)
    _lib.check(lib.tbns_pa_slice_fwd_tc(XF16.data_ptr(), Ws.data_ptr(), bs.data_ptr(), tau.data_ptr(), w16.data_ptr(), part.data_ptr(),
                                        B, N, H, D, G, 1, st), "slice_fwd_tc")
    torch.cuda.synchronize()
    assert O.rel_l2(w16.double().cpu().view(B, N, H, G), w_ref) < 4e-3          # bf16 storage of the weights (2^-9 per value)
    p = part.view(B, H, groups, G, D + 1).sum(2).double().cpu()
    assert O.rel_l2(p[..., :D], Tt_ref) < 2e-3                                   # sums of bf16 weights x bf16 features
    assert O.rel_l2(p[..., D], s_ref) < 1e-3                                     # sum_n w of the bf16 weights


@pytest.mark.parametrize("B,N,H,G", [(2, 4096, 8, 32), (1, 1000, 4, 64), (3, 130, 2, 32), (20, 4096, 8, 32)])
def test_tc_slice_bwd_matches_oracle(B, N, H, G):
    """tensor-core slice backward vs the fp64 oracle stage (oracle.slice_bwd) on bf16-representable inputs, plus the
    projection-bias gradients derived from its partials (tbns_pa_proj_bias_grad)"""
    from transformerbasednavierstokesolver_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    D = 32
    I = H * D
    g = torch.Generator().manual_seed(B * N + G + 1)
    XF16 = torch.randn(B * N, 2 * I, generator=g).bfloat16()
    Ws = (torch.randn(G, D, generator=g) * 0.4).bfloat16().float()
    bs = torch.randn(G, generator=g)
    tau = torch.linspace(0.3, 1.5, H)
    dw16 = torch.randn(B, N, H * G, generator=g).bfloat16()
    dTt = torch.randn(B, H, G, D, generator=g).bfloat16().float()
    ds = torch.randn(B, H, G, generator=g)
    XF64 = XF16.double().view(B, N, 2 * I)
    dXF_ref, dWs_ref, dbs_ref, dtau_ref = O.slice_bwd(dw16.double().view(B, N, H, G), dTt.double(), ds.double(), XF64, Ws.double(),
                                                      bs.double(), tau.double(), H, True)
    w_ref, s_ref, _ = O.slice_fwd(XF64, Ws.double(), bs.double(), tau.double(), H, True)
    # oracle bias gradients of the projections: db_x = sum_t dX, db_fx = sum_t dF
    dbx_ref, dbfx_ref = dXF_ref[..., :I].sum((0, 1)), dXF_ref[..., I:].sum((0, 1))
    groups = lib.tbns_slice_groups(B, N, H)
    st = torch.cuda.current_stream().cuda_stream
    XF16, Ws, bs, tau, dw16, dTt, ds = (t.to(dev) for t in (XF16, Ws, bs, tau, dw16, dTt, ds))
    dXF16 = torch.full((B * N, 2 * I), float("nan"), device=dev, dtype=torch.bfloat16)
    dWs_p = torch.full((B * H * groups, G * (D + 1)), float("nan"), device=dev)
    dtau_p = torch.full((B * H * groups,), float("nan"), device=dev)
    _lib.check(lib.tbns_pa_slice_bwd_tc(XF16.data_ptr(), Ws.data_ptr(), bs.data_ptr(), tau.data_ptr(), dw16.data_ptr(), dTt.data_ptr(),
                                        ds.data_ptr(), dXF16.data_ptr(), dWs_p.data_ptr(), dtau_p.data_ptr(), B, N, H, D, G, 1, st),
               "slice_bwd_tc")
    s_dev = s_ref.float().to(dev).contiguous()
    dbx = torch.full((I,), float("nan"), device=dev)
    dbfx = torch.full((I,), float("nan"), device=dev)
    _lib.check(lib.tbns_pa_proj_bias_grad(dWs_p.data_ptr(), Ws.data_ptr(), s_dev.data_ptr(), dTt.data_ptr(), dbx.data_ptr(), dbfx.data_ptr(),
                                          B, H, D, G, groups, st), "proj_bias_grad")
    torch.cuda.synchronize()
    assert O.rel_l2(dXF16.double().cpu().view(B, N, 2 * I), dXF_ref) < 6e-3          # bf16 storage + bf16 w / dL operands
    a = dWs_p.view(B, H, groups, G, D + 1).sum((0, 1, 2)).double().cpu()
    assert O.rel_l2(a[:, :D], dWs_ref) < 5e-3
    assert O.rel_l2(a[:, D], dbs_ref) < 1e-2       # sum_t dL: signed terms that largely cancel, each rounded to bf16
    ta = dtau_p.view(B, H, groups).sum((0, 2)).double().cpu()
    assert O.rel_l2(ta, dtau_ref) < 1e-2           # sum of signed dL'*L terms: cancellation amplifies operand rounding
    assert O.rel_l2(dbx.double().cpu(), dbx_ref) < 1e-2
    assert O.rel_l2(dbfx.double().cpu(), dbfx_ref) < 1e-4   # exact fp32 arithmetic on the oracle's s: only summation order differs


@pytest.mark.parametrize("M,K,R,Cout,need_dx", [(8192, 74, 512, 256, False), (972, 2, 256, 128, True), (4096, 65, 256, 128, True)])
def test_preprocess_mlp_tc(M, K, R, Cout, need_dx):
    """ops.MlpFn (preprocess on the tensor cores, K zero-padded to 64) against fp64 torch on bf16-representable operands:
    model/Transolver_Structured_Mesh_2D.py:13-38,206-207.  bf16-mode tolerance 2e-3 on the output (hidden activation is rounded to bf16), 5e-3 on gradients."""
    from transformerbasednavierstokesolver_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + K)
    r16 = lambda t: t.bfloat16().float()
    inp = r16(torch.randn(2, M // 2, K, generator=g)).to(dev).requires_grad_(need_dx)
    W1 = r16(torch.randn(R, K, generator=g) / K ** 0.5).to(dev).requires_grad_(True)
    b1 = (0.1 * torch.randn(R, generator=g)).to(dev).requires_grad_(True)
    W2 = r16(torch.randn(Cout, R, generator=g) / R ** 0.5).to(dev).requires_grad_(True)
    b2 = (0.1 * torch.randn(Cout, generator=g)).to(dev).requires_grad_(True)
    dout = torch.randn(2, M // 2, Cout, generator=g).to(dev)
    assert ops.mlp_tc_ok(K, R, Cout)
    out = ops.MlpFn.apply(inp, W1, b1, W2, b2)
    out.backward(dout)
    got = [out] + [t.grad for t in ((inp,) if need_dx else ()) + (W1, b1, W2, b2)]
    ref_in = [t.detach().cpu().double().requires_grad_(True) for t in (inp, W1, b1, W2, b2)]
    ri, rW1, rb1, rW2, rb2 = ref_in
    rout = torch.nn.functional.gelu(ri @ rW1.t() + rb1) @ rW2.t() + rb2
    rout.backward(dout.cpu().double())
    want = [rout.detach()] + [t.grad for t in ((ri,) if need_dx else ()) + (rW1, rb1, rW2, rb2)]
    for i, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape
        # output: one bf16 rounding (hidden activation); gradients: two more (incoming gradient, GELU' product)
        assert O.rel_l2(a.detach().cpu().double(), b) < (2e-3 if i == 0 else 5e-3)


@pytest.mark.parametrize("B,H,G,Cout,nchunk", [(3, 2, 32, 128, 5), (2, 4, 64, 256, 19), (20, 8, 32, 256, 11)])
def test_token_stage_mma_kernels_match_oracle(B, H, G, Cout, nchunk):
    """token stage on warp-level MMAs (dim_head 32; 3xTF32): forward outputs (s, tok, q, k, v, A, O, P, the bf16 operand
    copies) and every backward output against the fp64 oracle (model/Physics_Attention.py:102-111, SURVEY §8 a-bwd) at fp32
    accuracy."""
    from transformerbasednavierstokesolver_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    D = 32
    I = H * D
    g = torch.Generator().manual_seed(B * 100 + G + Cout)
    part = torch.rand(B, H, nchunk, G, D + 1, generator=g) * 2 - 0.7
    part[..., D] = torch.rand(B, H, nchunk, G, generator=g) * 30 + 0.5          # slice norms are positive
    Wq, Wk, Wv = (torch.randn(D, D, generator=g) * 0.3 for _ in range(3))
    Wo = torch.randn(Cout, I, generator=g) * 0.1
    dP = torch.randn(B, H * G, Cout, generator=g)
    p64 = part.double().sum(2)
    s_ref, Tt_ref = p64[..., D], p64[..., :D]
    st = O.token_attn_fwd(s_ref, Tt_ref, Wq.double(), Wk.double(), Wv.double(), Wo.double())
    rdTt, rds, rdWq, rdWk, rdWv, rdWo = O.token_attn_bwd(dP.double(), s_ref, Tt_ref, st, Wq.double(), Wk.double(), Wv.double(), Wo.double())

    f32 = dict(device=dev, dtype=torch.float32)
    s = torch.empty(B, H, G, **f32)
    Tt, tok, q, k, v, Oo = (torch.full((B, H, G, D), float("nan"), **f32) for _ in range(6))
    A = torch.full((B, H, G, G), float("nan"), **f32)
    P = torch.full((B, H * G, Cout), float("nan"), **f32)
    P16 = torch.empty(B, H * G, Cout, device=dev, dtype=torch.bfloat16)
    PT16 = torch.empty(B, Cout, H * G, device=dev, dtype=torch.bfloat16)
    dd = lambda t: t.to(dev).contiguous()
    partd, Wqd, Wkd, Wvd, Wod, dPd = dd(part), dd(Wq), dd(Wk), dd(Wv), dd(Wo), dd(dP)
    stream = torch.cuda.current_stream().cuda_stream
    p_ = lambda t: t.data_ptr()
    _lib.check(lib.tbns_pa_token_attn_fwd(p_(partd), nchunk, p_(Wqd), p_(Wkd), p_(Wvd), p_(Wod), p_(s), p_(Tt), p_(tok), p_(q), p_(k), p_(v),
                                          p_(A), p_(Oo), p_(P), p_(P16), p_(PT16), B, H, D, G, Cout, stream), "token fwd")
    torch.cuda.synchronize()
    tol = 2e-6
    assert O.rel_l2(s.cpu(), s_ref) < tol and O.rel_l2(Tt.cpu(), Tt_ref) < tol
    for name, got in (("tok", tok), ("q", q), ("k", k), ("v", v), ("A", A), ("O", Oo)):
        assert O.rel_l2(got.cpu(), st[name]) < tol, name
    assert O.rel_l2(P.cpu(), st["P"]) < tol
    assert O.rel_l2(PT16.float().cpu(), st["P"].transpose(1, 2)) < 4e-3                     # bf16 storage
    Pc = st["P"].reshape(B, H, G, Cout)
    Pc = (Pc - Pc.mean(2, keepdim=True)).reshape(B, H * G, Cout)                          # centred over the slices of a head
    assert O.rel_l2(P16.float().cpu(), Pc) < 8e-3   # bf16 storage of fp32 differences (cancellation noise of the centring itself)
    dTt = torch.full((B, H, G, D), float("nan"), **f32)
    ds = torch.full((B, H, G), float("nan"), **f32)
    dWqkv = torch.full((B * H, 3 * D * D), float("nan"), **f32)
    dWo = torch.full((B, Cout * I), float("nan"), **f32)
    _lib.check(lib.tbns_pa_token_attn_bwd(p_(dPd), p_(Wqd), p_(Wkd), p_(Wvd), p_(Wod), p_(s), p_(tok), p_(q), p_(k), p_(v), p_(A), p_(Oo),
                                          p_(dTt), p_(ds), p_(dWqkv), p_(dWo), B, H, D, G, Cout, stream), "token bwd")
    torch.cuda.synchronize()
    gtol = 1e-5
    assert O.rel_l2(dTt.cpu(), rdTt) < gtol and O.rel_l2(ds.cpu(), rds) < gtol
    got = dWqkv.double().sum(0).view(3, D, D).cpu()
    assert O.rel_l2(got[0], rdWq) < gtol and O.rel_l2(got[1], rdWk) < gtol and O.rel_l2(got[2], rdWv) < gtol
    assert O.rel_l2(dWo.double().sum(0).view(Cout, I).cpu(), rdWo) < gtol
