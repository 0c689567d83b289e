"""GPU tests of the training / rollout harness around the CUDA path: SOL unrolling against the reference golden rollout,
batched vs literal teacher forcing, CUDA-graph replay vs eager steps."""
import copy

import pytest
import torch

from oracle import physics_attention as O

pytestmark = pytest.mark.gpu


def _small_model(dev, precision="fp32"):
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
    pkg.set_default_precision(precision)
    torch.manual_seed(3)
    return Model(space_dim=2, n_layers=2, n_hidden=64, n_head=4, fun_dim=4, out_dim=1, slice_num=8, ref=4, unified_pos=1, H=8, W=8).to(dev)


def test_sol_unrolled_forward_matches_reference_rollout(golden):
    """SOL_Transolver_Structured_Mesh_2D(look_ahead=n).forward == n-th frame of the reference's closed-loop rollout"""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200.model.SOL_Transolver_Structured_Mesh_2D import SOL_Transolver_Structured_Mesh_2D
    dev = torch.device("cuda:0")
    fx = golden("model_2d_unified.pt")
    pkg.set_default_precision("fp32")
    try:
        n = fx["rollout"].shape[-1]
        sol = SOL_Transolver_Structured_Mesh_2D(**fx["kwargs"], step=1, look_ahead=n)
        sol.transolver_model.load_state_dict({k: v.float() for k, v in fx["state"].items()}, strict=True)
        sol = sol.to(dev)
        with torch.no_grad():
            u = sol(fx["x"].float().to(dev), fx["fx"].float().to(dev))
        assert O.rel_l2(u[..., 0].cpu(), fx["rollout"][..., n - 1]) < 5e-5
    finally:
        pkg.set_default_precision("bf16")


def test_batched_teacher_forcing_equals_loop_on_gpu():
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import train
    dev = torch.device("cuda:0")
    try:
        m = _small_model(dev, "fp32")
        x, fx, yy = train.synthetic_ns_batch(2, 8, 4, 3, seed=1, device=dev)
        opt = torch.optim.SGD(m.parameters(), lr=0.0)
        la = train.train_step(m, opt, None, None, x, fx, yy, T=3, step=1, batched=True)
        ga = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        lb = train.train_step(m, opt, None, None, x, fx, yy, T=3, step=1, batched=False)
        assert abs(float(la) - float(lb)) < 1e-5 * abs(float(lb))
        for k, p in m.named_parameters():
            if p.grad is not None and float(ga[k].abs().max()) > 1e-8:
                assert O.rel_l2(p.grad, ga[k]) < 2e-3, k
    finally:
        pkg.set_default_precision("bf16")


@pytest.mark.parametrize("buckets", [1, 2])
def test_graphed_step_matches_eager_steps(buckets):
    """CUDA-graph replay of the optimizer step (one backward graph, or backward cut into `buckets` stage graphs whose
    gradient ranges are all-reduced separately under data parallelism) against the eager loop."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import train
    dev = torch.device("cuda:0")
    try:
        m1 = _small_model(dev, "fp32")
        m2 = copy.deepcopy(m1)
        batches = [train.synthetic_ns_batch(2, 8, 4, 3, seed=10 + i, device=dev) for i in range(3)]
        g1 = train.FlatGradients(m1.parameters())
        o1 = torch.optim.AdamW(m1.parameters(), lr=1e-3, weight_decay=1e-5, fused=True)
        for b in batches:
            train.train_step(m1, o1, None, g1, *b, T=3, step=1, batched=True)
        g2 = train.FlatGradients(m2.parameters())
        o2 = torch.optim.AdamW(m2.parameters(), lr=torch.tensor(1e-3, device=dev), weight_decay=1e-5, fused=True, capturable=True)
        state0 = copy.deepcopy(m2.state_dict())
        gs = train.GraphedTrainStep(m2, o2, None, g2, batches[0], T=3, step=1, batched=True, warmup=2, buckets=buckets)
        assert gs.nb == buckets
        # warm-up and capture advanced the weights / optimizer state: restart both from the initial point
        m2.load_state_dict(state0)
        for st in o2.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
        for b in batches:
            gs(b)
        torch.cuda.synchronize()
        for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
            assert torch.allclose(p1, p2, rtol=1e-4, atol=1e-6), k
    finally:
        pkg.set_default_precision("bf16")


def test_flat_adamw_matches_torch_adamw_with_onecycle():
    """train.FlatAdamW (one libtbns kernel over flat parameter / gradient / moment buffers) against torch.optim.AdamW with
    OneCycleLR driving lr AND beta1 (cycle_momentum) as in exp_ns.py:172-176:
    (a) the optimizers alone on identical gradient streams (update rule, bias corrections, decoupled weight decay);
    (b) a model trained eagerly with FlatAdamW vs the same steps replayed from CUDA graphs (GraphedTrainStep hands the changing
        hyper-parameters to the device before each replay);
    (c) the loss trajectory against the stock optimizer (the parameters live at different addresses, so gradients may differ in
        the last bit and Adam turns that into O(lr) differences on near-zero-gradient entries: trajectories, not bits)."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import train
    dev = torch.device("cuda:0")
    # (a)
    torch.manual_seed(0)
    a1 = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.Linear(64, 3), torch.nn.Linear(3, 1)).to(dev)   # odd sizes: padded layout
    a2 = copy.deepcopy(a1)
    oa1 = torch.optim.AdamW(a1.parameters(), lr=1e-3, weight_decay=1e-2)
    sa1 = torch.optim.lr_scheduler.OneCycleLR(oa1, max_lr=5e-3, total_steps=16)
    ga2 = train.FlatGradients(a2.parameters())
    oa2 = train.FlatAdamW(a2.parameters(), ga2, lr=1e-3, weight_decay=1e-2)
    sa2 = torch.optim.lr_scheduler.OneCycleLR(oa2, max_lr=5e-3, total_steps=16)
    assert all(p.data_ptr() % 16 == 0 for p in a2.parameters())              # views of the flat buffer stay 16-byte aligned
    for _ in range(8):
        gs = [torch.randn_like(p) for p in a1.parameters()]
        for p, g in zip(a1.parameters(), gs):
            p.grad = g.clone()
        for p, g in zip(a2.parameters(), gs):
            p.grad.copy_(g)
        oa1.step(); sa1.step(); oa2.step(); sa2.step()
    torch.cuda.synchronize()
    assert oa2.param_groups[0]["betas"][0] != 0.9                             # the scheduler really cycled beta1
    for p1, p2 in zip(a1.parameters(), a2.parameters()):
        assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-7)
    try:
        m1 = _small_model(dev, "fp32")
        m2 = copy.deepcopy(m1)
        m3 = copy.deepcopy(m1)
        batches = [train.synthetic_ns_batch(2, 8, 4, 3, seed=30 + i, device=dev) for i in range(5)]
        g1 = train.FlatGradients(m1.parameters())
        o1 = torch.optim.AdamW(m1.parameters(), lr=1e-3, weight_decay=1e-2)
        s1 = torch.optim.lr_scheduler.OneCycleLR(o1, max_lr=5e-3, total_steps=16)
        g2 = train.FlatGradients(m2.parameters())
        o2 = train.FlatAdamW(m2.parameters(), g2, lr=1e-3, weight_decay=1e-2)
        s2 = torch.optim.lr_scheduler.OneCycleLR(o2, max_lr=5e-3, total_steps=16)
        l1, l2 = [], []
        for b in batches:
            l1.append(float(train.train_step(m1, o1, s1, g1, *b, T=3, step=1, batched=True)))
            l2.append(float(train.train_step(m2, o2, s2, g2, *b, T=3, step=1, batched=True)))
        for x1, x2 in zip(l1, l2):                                             # (c)
            assert abs(x1 - x2) <= 2e-3 * abs(x1), (l1, l2)
        # (b) flat optimizer through CUDA graphs: restart from the initial point after warm-up / capture
        g3 = train.FlatGradients(m3.parameters())
        o3 = train.FlatAdamW(m3.parameters(), g3, lr=1e-3, weight_decay=1e-2)
        state0 = copy.deepcopy(m3.state_dict())
        gs = train.GraphedTrainStep(m3, o3, None, g3, batches[0], T=3, step=1, batched=True, warmup=2)
        m3.load_state_dict(state0)
        o3.exp_avg.zero_()
        o3.exp_avg_sq.zero_()
        o3.t = 0
        o3.param_groups[0]["lr"], o3.param_groups[0]["betas"] = 1e-3, (0.9, 0.999)
        gs.sched = torch.optim.lr_scheduler.OneCycleLR(o3, max_lr=5e-3, total_steps=16)
        for b in batches:
            gs(b)
        torch.cuda.synchronize()
        for (k, p2), (_, p3) in zip(m2.named_parameters(), m3.named_parameters()):
            assert torch.allclose(p2, p3, rtol=1e-4, atol=1e-6), k
        # optimizer checkpoint round trip (moments live in flat buffers)
        sd = o3.state_dict()
        o3.exp_avg.zero_()
        o3.t = 0
        o3.load_state_dict(sd)
        assert o3.t == len(batches) and torch.equal(o3.exp_avg, sd["flat"]["exp_avg"]) and float(o3.exp_avg.abs().sum()) > 0
        # checkpoints still load into the relocated parameters
        m3.load_state_dict(m1.state_dict())
        assert torch.equal(next(m3.parameters()), next(m1.parameters())) and next(m3.parameters()).data_ptr() == o3.flat_p.data_ptr()
    finally:
        pkg.set_default_precision("bf16")


def test_direct_gradient_destinations_equal_gathered_gradients():
    """FlatGradients.begin() arms every parameter with its view of the flat buffer; the bf16-route backward stages write the
    large weight gradients (conv projections, MLP weights) straight into those views (ops._claim_grad).  The flat buffer must
    equal, bit for bit, the gradients of a plain backward pass - for a parameter used once (teacher-forced step) and for
    parameters shared by several calls (unrolled training: the first gradient lands in place, autograd accumulates the
    rest into it)."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import ops, train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
    from transformerbasednavierstokesolver_b200.model.SOL_Transolver_Structured_Mesh_2D import SOL_Transolver_Structured_Mesh_2D
    dev = torch.device("cuda:0")
    pkg.set_default_precision("bf16")
    torch.manual_seed(3)
    cfg = dict(space_dim=2, n_layers=2, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
               slice_num=32, ref=8, unified_pos=1, H=64, W=64)
    sol = SOL_Transolver_Structured_Mesh_2D(step=1, look_ahead=5, **cfg).to(dev)
    model = sol.transolver_model
    assert isinstance(model, Model)
    x, fx, yy = train.synthetic_ns_batch(2, 64, 10, 10, seed=5, device=dev)
    cases = {"teacher_forced": lambda: train.step_loss(model, x, fx, yy, 10, 1, True),
             "unrolled_la5": lambda: train.unrolled_step_loss(sol, x, fx, yy, 10, 1, True)}   # every weight is used by 5 chained calls
    for name, loss_fn in cases.items():
        for p in model.parameters():
            p.grad = None
        loss_fn().backward()
        want = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        g = train.FlatGradients(model.parameters())
        g.flat.fill_(float("nan"))
        g.begin()
        loss_fn().backward()
        conv = model.blocks[0].Attn.in_project_x.weight
        view = g.views[[id(p) for p in g.params].index(id(conv))]
        if name == "teacher_forced":
            assert conv.grad.data_ptr() == view.data_ptr(), "the conv weight gradient was not written in place"
        g.finish()
        torch.cuda.synchronize()
        for k, p in model.named_parameters():
            if k in want:
                assert torch.equal(p.grad, want[k]), (name, k)
            assert p.grad.data_ptr() == g.views[[id(q) for q in g.params].index(id(p))].data_ptr()
            assert getattr(p, "_tbns_grad_dst", None) is None      # disarmed after the pass



@pytest.mark.parametrize("n", [1003, 4096, 7])
def test_adamw_flat_kernel_matches_formula(n):
    """tbns_adamw_flat through the C ABI on raw buffers (vector body + scalar tail) against the AdamW formula in fp64"""
    from transformerbasednavierstokesolver_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n)
    p, gr, m, v = (torch.randn(n, generator=g) for _ in range(4))
    v = v.abs()
    lr, b1, b2, eps, wd, t = 3e-3, 0.87, 0.995, 1e-8, 1e-2, 5
    hp = torch.tensor([lr, b1, b2, eps, wd, 1 - b1 ** t, 1 - b2 ** t, 0.0])
    pd, gd, md, vd = (x.double() for x in (p, gr, m, v))
    pd = pd * (1 - lr * wd)
    md = b1 * md + (1 - b1) * gd
    vd = b2 * vd + (1 - b2) * gd * gd
    pd = pd - (lr / (1 - b1 ** t)) * md / (vd.sqrt() / (1 - b2 ** t) ** 0.5 + eps)
    pc, gc, mc, vc, hc = (x.to(dev) for x in (p, gr, m, v, hp))
    _lib.check(lib.tbns_adamw_flat(pc.data_ptr(), gc.data_ptr(), mc.data_ptr(), vc.data_ptr(), hc.data_ptr(), n,
                                   torch.cuda.current_stream().cuda_stream), "tbns_adamw_flat")
    torch.cuda.synchronize()
    assert torch.allclose(pc.cpu().double(), pd, rtol=2e-6, atol=1e-7)
    assert torch.allclose(mc.cpu().double(), md, rtol=2e-6, atol=1e-7)
    assert torch.allclose(vc.cpu().double(), vd, rtol=2e-6, atol=1e-7)
