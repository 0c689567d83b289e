"""CPU tests of the data-parallel host logic (world_size 2, gloo): flat-gradient all-reduce sums gradients so that a DP
run equals a single-process run on the global batch (the reference loss is a SUM over samples, utils/testloss.py:40),
replicas start identical after broadcast, and the batched teacher-forced step equals the literal exp_ns.py loop."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transformerbasednavierstokesolver_b200 import train


class TinyModel(torch.nn.Module):
    """stands in for the Transolver on CPU: same call signature model(x, fx=...) -> [B, N, 1]"""

    def __init__(self, t_in):
        super().__init__()
        self.l1 = torch.nn.Linear(2 + t_in, 16)
        self.l2 = torch.nn.Linear(16, 1)

    def forward(self, x, fx):
        return self.l2(torch.tanh(self.l1(torch.cat((x, fx), -1))))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)            # different init per rank on purpose
    model = TinyModel(4)
    train.broadcast_parameters(model)        # -> identical replicas
    grads = train.FlatGradients(model.parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=1e-5)
    x, fx, yy = train.synthetic_ns_batch(4, 6, 4, 3, seed=7)          # global batch 4
    sl = slice(rank * 2, rank * 2 + 2)                                 # this rank's shard
    for _ in range(3):
        loss = train.train_step(model, opt, None, grads, x[sl], fx[sl], yy[sl], T=3, step=1, batched=True)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put((gathered[0].tolist(), gathered[1].tolist(), float(loss)))   # plain lists: no shared-memory handles to outlive the worker
    dist.destroy_process_group()


def test_dp2_equals_single_process_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    p0, p1, _ = q.get(timeout=120)
    p0, p1 = torch.tensor(p0), torch.tensor(p1)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert torch.equal(p0, p1), "replicas diverged"
    # single process, global batch, same init as rank 0
    torch.manual_seed(100)
    model = TinyModel(4)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=1e-5)
    x, fx, yy = train.synthetic_ns_batch(4, 6, 4, 3, seed=7)
    for _ in range(3):
        train.train_step(model, opt, None, None, x, fx, yy, T=3, step=1, batched=False)
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(p0, ref, rtol=1e-5, atol=1e-6)


def test_batched_teacher_forcing_equals_loop():
    torch.manual_seed(0)
    model = TinyModel(5)
    x, fx, yy = train.synthetic_ns_batch(3, 5, 5, 4, seed=3)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    la = train.train_step(model, opt, None, None, x, fx, yy, T=4, step=1, batched=True)
    ga = [p.grad.clone() for p in model.parameters()]
    lb = train.train_step(model, opt, None, None, x, fx, yy, T=4, step=1, batched=False)
    gb = [p.grad.clone() for p in model.parameters()]
    assert abs(float(la) - float(lb)) < 1e-5 * abs(float(lb))
    for a, b in zip(ga, gb):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


def test_rollout_feeds_predictions_back():
    model = TinyModel(3)
    x, fx, _ = train.synthetic_ns_batch(2, 4, 3, 3, seed=1)
    r = train.rollout(model, x, fx, T=3, step=1)
    assert r.shape == (2, 16, 3)
    f = fx.clone()
    for t in range(3):
        im = model(x, fx=f)
        assert torch.allclose(im[..., 0], r[..., t], atol=1e-6)
        f = torch.cat((f[..., 1:], im), -1)


def test_flat_gradients_gather_equals_accumulate():
    """FlatGradients.begin()/finish() (autograd hands over whole gradient tensors, one multi-tensor copy into the flat
    all-reduce buffer) == zero() + in-place accumulation; a parameter that receives no gradient reads as zero and
    p.grad always ends up as a view of the flat buffer (what the optimizer and the all-reduce see)."""
    torch.manual_seed(0)
    model = TinyModel(4)
    model.unused = torch.nn.Parameter(torch.ones(3))         # never touched by forward
    x, fx, yy = train.synthetic_ns_batch(2, 6, 4, 3, seed=3)
    g = train.FlatGradients(model.parameters())
    g.zero()
    train.step_loss(model, x, fx, yy, 3, 1, True).backward()
    want = g.flat.clone()
    g.flat.fill_(123.0)                                       # stale contents must not survive
    g.begin()
    assert all(p.grad is None for p in model.parameters())
    train.step_loss(model, x, fx, yy, 3, 1, True).backward()
    g.finish()
    off = 0
    for p in g.params:     # every view starts on a 16-byte boundary (FlatAdamW / libtbns read them with vector loads)
        assert p.grad.data_ptr() == g.flat.data_ptr() + 4 * off and p.grad.shape == p.shape and off % 4 == 0
        assert torch.equal(g.flat[off:off + p.numel()], want[off:off + p.numel()])
        off += g.padded(p.numel())
    assert off == g.flat.numel()
    assert float(model.unused.grad.abs().sum()) == 0.0


def test_flat_gradients_arm_destinations_once_per_pass():
    """host logic of the direct gradient destinations (ops._claim_grad): begin() arms every parameter with (view, token), a
    view is handed out once per backward pass and only for a matching shape, finish() disarms; a second begin() re-arms."""
    from transformerbasednavierstokesolver_b200 import ops
    model = TinyModel(3)
    g = train.FlatGradients(model.parameters())
    p = g.params[0]
    g.begin()
    dst = ops._grad_dst(p)
    assert dst is not None and dst[0] is g.views[0]
    a = ops._claim_grad(dst, p.shape)
    assert a.data_ptr() == g.views[0].data_ptr() and a is not g.views[0]          # alias: a fresh tensor object on the view
    b = ops._claim_grad(dst, p.shape)
    assert b.data_ptr() != g.views[0].data_ptr()                                  # second user of the parameter in this pass
    c = ops._claim_grad(ops._grad_dst(g.params[1]), (7, 7))
    assert c.shape == (7, 7)                                                      # shape mismatch -> plain allocation
    assert ops._claim_grad(None, (2, 2)).shape == (2, 2)
    for q in g.params:
        q.grad = torch.zeros_like(q)
    g.finish()
    assert all(getattr(q, "_tbns_grad_dst", None) is None for q in g.params)
    g.begin()
    assert ops._claim_grad(ops._grad_dst(p), p.shape).data_ptr() == g.views[0].data_ptr()   # new pass, new token
    g.finish()
