"""Batch-data-parallel training / rollout harness around the Transolver models (one process per GPU).

Restates the semantics of the reference drivers without their data loading:
  exp_ns.py:191-218                 teacher-forced optimizer step (T model calls, one backward, AdamW + OneCycleLR)
  exp_ns.py:225-241, ns_vorticity_unrolling.py:264-286   closed-loop rollout evaluation
  ns_vorticity_unrolling.py:225-244  unrolled (look_ahead) training through SOL_Transolver_Structured_Mesh_2D
and adds what the reference does not have: gradient all-reduce over NCCL.  The reference loss is a
SUM over the batch (utils/testloss.py:40, size_average=False), so replicas SUM gradients — a DP run equals a
single-GPU run with the same global batch.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist


# diagnostic only (measures what the exchange costs): replicas drift apart without it
_NO_ALLREDUCE = os.environ.get("TBNS_DEBUG_NO_ALLREDUCE", "0") == "1"


def rel_l2_sum(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """TestLoss(size_average=False).rel  (utils/testloss.py:31-42)"""
    n = pred.shape[0]
    diff = torch.linalg.vector_norm(pred.reshape(n, -1) - target.reshape(n, -1), dim=1)
    return (diff / torch.linalg.vector_norm(target.reshape(n, -1), dim=1)).sum()


class FlatGradients:
    """All parameter gradients live in ONE flat fp32 buffer (p.grad are views), so the data-parallel exchange is a single
    NCCL all-reduce (cfg 1: 11.2 M elements = 44.8 MB) and the optimizer reads contiguous memory."""

    ALIGN = 4   # every view starts on a 16-byte boundary (FlatAdamW lays the PARAMETERS out the same way, and libtbns reads
                # weights with 16-byte vector loads / TMA); the padding elements stay zero

    @staticmethod
    def padded(n: int) -> int:
        return -(-n // FlatGradients.ALIGN) * FlatGradients.ALIGN

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(self.padded(p.numel()) for p in self.params)
        self.flat = torch.zeros(total, device=self.params[0].device, dtype=torch.float32)
        off = 0
        self.views = []
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += self.padded(p.numel())

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def begin(self):
        """start of a backward pass: detach the views so autograd hands each parameter its gradient tensor as is
        (no `grad += new` kernel per parameter, no zero fill of the flat buffer)."""
        token = object()   # identifies this backward pass: a view is handed out at most once per pass (ops._claim_grad)
        for p, v in zip(self.params, self.views):
            p.grad = None
            p._tbns_grad_dst = (v, token)

    def finish(self):
        """end of a backward pass: gather the per-parameter gradients into the flat buffer with one multi-tensor copy and
        point p.grad back at the views (parameters that received no gradient read as zero)."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v
            p._tbns_grad_dst = None

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not _NO_ALLREDUCE:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def broadcast_parameters(model: torch.nn.Module, src: int = 0):
    """identical replicas: rank `src` wins (parameters and buffers)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.detach(), src)
        from . import ops
        ops.invalidate_weight_caches()   # the collective wrote into the masters behind autograd's version counters


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (exp_ns.py:172-173: AdamW(lr, weight_decay); amsgrad / maximize off) as ONE libtbns kernel
    over flat buffers: the parameters become views into one fp32 buffer laid out like `grads.flat`, the moments are flat, and
    a step is a single elementwise pass (`tbns_adamw_flat`, ~50 us for the 11.2 M parameters of cfg 1; torch's fused
    multi-tensor AdamW needs ~0.45 ms for the same ~180 tensors).  Hyper-parameters are ordinary `param_groups` entries, so
    `OneCycleLR` drives `lr` and (cycle_momentum) `betas[0]` as with the stock optimizer; before every launch they are copied
    into a small device array that the kernel reads - which also makes the step replayable from a CUDA graph with changing
    lr / beta1 (`GraphedTrainStep` calls `prepare_step()` before each replay).  One parameter group.  Build it AFTER the model
    is on its device (`module.to()` would re-allocate the parameters and detach them from the flat buffer)."""

    def __init__(self, params, grads: FlatGradients, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = [p for p in params if p.requires_grad]
        if [id(p) for p in params] != [id(p) for p in grads.params]:
            raise ValueError("FlatAdamW: parameters must be the ones (and in the order) of the FlatGradients buffer")
        if not params[0].is_cuda:
            raise RuntimeError("transformerbasednavierstokesolver_b200 has no CPU path: FlatAdamW needs CUDA parameters")
        super().__init__(params, dict(lr=float(lr), betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay)))
        self.grads = grads
        dev = params[0].device
        self.flat_p = torch.zeros_like(grads.flat)
        off = 0
        with torch.no_grad():
            for p in params:
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view      # state_dict / checkpoint loading keep working: they copy into the views
                off += grads.padded(p.numel())
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.hp = torch.zeros(8, device=dev, dtype=torch.float32)
        # ring of pinned host slots: the copy of step t is asynchronous, so its source must stay untouched until it has run;
        # the host cannot be 256 optimizer steps ahead of the device (launch-queue depth)
        self._hp_host = torch.zeros(256, 8, dtype=torch.float32).pin_memory()
        self.t = 0
        self._order = tuple(id(p) for p in grads.params)

    # checkpoint / resume: the moments live in the flat buffers, not in `self.state`
    def state_dict(self):
        sd = super().state_dict()
        sd["flat"] = {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "t": self.t}
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        flat = sd.pop("flat")
        super().load_state_dict(sd)
        self.exp_avg.copy_(flat["exp_avg"])
        self.exp_avg_sq.copy_(flat["exp_avg_sq"])
        self.t = int(flat["t"])

    def prepare_step(self):
        """advance the step count and hand this step's hyper-parameters to the device (asynchronous 32-byte copy)"""
        g = self.param_groups[0]
        self.t += 1
        lr = float(g["lr"])
        b1, b2 = (float(b) for b in g["betas"])
        h = self._hp_host[self.t % self._hp_host.shape[0]]
        h[0], h[1], h[2], h[3], h[4] = lr, b1, b2, float(g["eps"]), float(g["weight_decay"])
        h[5], h[6] = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t
        self.hp.copy_(h, non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib
        if closure is not None:
            raise RuntimeError("FlatAdamW: closures are not supported")
        if self._order != tuple(id(p) for p in self.grads.params):
            raise RuntimeError("FlatAdamW: the FlatGradients layout changed (bucketed stage order): rebuild the optimizer")
        if not torch.cuda.is_current_stream_capturing():
            self.prepare_step()       # under capture only the kernel is recorded; the replaying host calls prepare_step()
        from . import ops
        ops._count(1)
        lib = _lib.load()
        _lib.check(lib.tbns_adamw_flat(self.flat_p.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.hp.data_ptr(), self.flat_p.numel(),
                                       torch.cuda.current_stream().cuda_stream), "tbns_adamw_flat")
        return None


def teacher_forced_inputs(fx: torch.Tensor, yy: torch.Tensor, T: int, step: int = 1) -> torch.Tensor:
    """The T/step inputs of exp_ns.py:197-208 are known up front (ground truth is fed back): input t is the window
    cat(fx, yy)[..., t*step : t*step + T_in].  Returns them stacked on the batch axis, call-major: [T/step * B, N, T_in]."""
    T_in = fx.shape[-1]
    full = torch.cat((fx, yy), dim=-1)
    return torch.cat([full[..., t:t + T_in] for t in range(0, T, step)], dim=0)


def step_loss(model, x, fx, yy, T: int, step: int = 1, batched: bool = True) -> torch.Tensor:
    """summed relative-L2 loss of the T/step teacher-forced calls (exp_ns.py:197-208).
    batched=True evaluates them as one batch of T/step*B samples — same math (the calls are independent given the ground
    truth), 1/T as many kernel launches and T x more tokens per launch."""
    bsz = x.shape[0]
    if batched:
        calls = T // step
        fx_all = teacher_forced_inputs(fx, yy, T, step)
        x_all = x.repeat(calls, 1, 1)
        y_all = torch.cat([yy[..., t:t + step] for t in range(0, T, step)], dim=0)
        im = model(x_all, fx=fx_all)
        loss = rel_l2_sum(im.reshape(calls * bsz, -1), y_all.reshape(calls * bsz, -1))
    else:
        loss = 0
        for t in range(0, T, step):
            y = yy[..., t:t + step]
            im = model(x, fx=fx)
            loss = loss + rel_l2_sum(im.reshape(bsz, -1), y.reshape(bsz, -1))
            fx = torch.cat((fx[..., step:], y), dim=-1)
    return loss


def train_step(model, optimizer, scheduler, grads: Optional[FlatGradients], x, fx, yy, T: int, step: int = 1,
               batched: bool = True, max_grad_norm: Optional[float] = None) -> torch.Tensor:
    """One optimizer step with exp_ns.py:191-218 semantics; returns the (local) summed step loss as a 0-d tensor."""
    if grads is not None:
        grads.begin()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss = step_loss(model, x, fx, yy, T, step, batched)
    loss.backward()
    if grads is not None:
        grads.finish()
        grads.all_reduce()
    if max_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)  # global norm, after the all-reduce
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


def unrolled_step_loss(sol_model, x, fx, yy, T: int, step: int = 1, batched: bool = True) -> torch.Tensor:
    """summed relative-L2 loss of one ns_vorticity_unrolling.py:225-244 batch: windows of `look_ahead` CHAINED model calls
    (the prediction is fed back inside a window), loss on the last prediction of each window, ground truth fed back between
    windows.  Because the ground truth is fed back between windows, the input of window t is known up front
    (cat(fx, yy)[..., t : t + T_in]); batched=True stacks the windows on the batch axis, so the step is `look_ahead`
    sequential calls on T/look_ahead * B samples instead of T calls on B samples - same math."""
    bsz = x.shape[0]
    look_ahead = sol_model.n
    off = look_ahead * step
    starts = list(range(0, T - off + 1, off))
    if batched:
        T_in = fx.shape[-1]
        full = torch.cat((fx, yy), dim=-1)
        win = torch.cat([full[..., t:t + T_in] for t in starts], dim=0)
        y = torch.cat([yy[..., t + off - step:t + off] for t in starts], dim=0)
        im = sol_model(x.repeat(len(starts), 1, 1), win)
        return rel_l2_sum(im.reshape(len(starts) * bsz, -1), y.reshape(len(starts) * bsz, -1))
    loss = 0
    for t in starts:
        y = yy[..., t + off - step:t + off]
        im = sol_model(x, fx)
        loss = loss + rel_l2_sum(im.reshape(bsz, -1), y.reshape(bsz, -1))
        fx = torch.cat((fx[..., off:], yy[..., t:t + off]), dim=-1)
    return loss


def unrolled_train_step(sol_model, optimizer, scheduler, grads: Optional[FlatGradients], x, fx, yy, T: int, step: int = 1,
                        batched: bool = False):
    """one optimizer step of ns_vorticity_unrolling.py:225-244 (see unrolled_step_loss)"""
    if grads is not None:
        grads.begin()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss = unrolled_step_loss(sol_model, x, fx, yy, T, step, batched)
    loss.backward()
    if grads is not None:
        grads.finish()
        grads.all_reduce()
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


@torch.no_grad()
def rollout(model: Callable, x, fx, T: int, step: int = 1) -> torch.Tensor:
    """closed-loop autoregressive rollout (predictions fed back; exp_ns.py:225-241, ns_vorticity_unrolling.py:264-286):
    returns [B, N, T].

    Models of this package run the FUSED rollout step: all frames live in one history buffer [B, N, T_in + T]; the input of
    step t is the strided window hist[..., t : t + T_in] (read directly by the packed preprocess, ops.PackedMlpFn) and the
    last layer writes its prediction straight into column T_in + t (`out=`), so the reference's per-step
    `cat(fx[..., step:], im)` and the final `cat(preds)` do not exist.  Any other callable takes the literal loop."""
    if getattr(model, "_packed_preprocess_ok", None) is not None and fx.is_cuda and model._packed_preprocess_ok(x, fx):
        B, N, T_in = fx.shape
        hist = torch.empty(B, N, T_in + T, device=fx.device, dtype=torch.float32)
        hist[..., :T_in] = fx
        for t in range(0, T, step):
            model(x, fx=hist[..., t:t + T_in], out=hist[..., T_in + t:T_in + t + step])
        return hist[..., T_in:]
    preds = []
    for _ in range(0, T, step):
        im = model(x, fx=fx)
        preds.append(im)
        fx = torch.cat((fx[..., step:], im), dim=-1)
    return torch.cat(preds, dim=-1)


def synthetic_ns_batch(batch: int, h: int, T_in: int, T: int, seed: int, device="cpu", pin: bool = False):
    """positions as exp_ns.py:88-94 (meshgrid of linspace(0,1,h)); fields ~ N(0, 0.38^2) (phiflow velocity statistics,
    data_generation.ipynb cell 5).  Returns x [B,h*h,2], fx [B,h*h,T_in], yy [B,h*h,T]."""
    g = torch.Generator().manual_seed(seed)
    lin = torch.linspace(0, 1, h)
    gx, gy = torch.meshgrid(lin, lin, indexing="xy")
    pos = torch.stack((gx.reshape(-1), gy.reshape(-1)), -1).unsqueeze(0).repeat(batch, 1, 1)
    fx = 0.38 * torch.randn(batch, h * h, T_in, generator=g)
    yy = 0.38 * torch.randn(batch, h * h, T, generator=g)
    out = (pos.contiguous(), fx, yy)
    if pin:
        out = tuple(t.pin_memory() for t in out)
    if device != "cpu":
        out = tuple(t.to(device, non_blocking=True) for t in out)
    return out


class DeferredLoss:
    """Per-step loss logging that does not drain the GPU: every step's scalar is copied to pinned host memory
    asynchronously and read one step late (after the NEXT step has been enqueued), so the host never waits for the step
    it has just launched.  `push` after each step, `flush` at the end; `values` holds one float per step."""

    def __init__(self):
        self.host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.ev = [torch.cuda.Event() for _ in range(2)]
        self.n, self.values = 0, []

    def push(self, loss_dev: torch.Tensor):
        i = self.n & 1
        self.host[i].copy_(loss_dev, non_blocking=True)
        self.ev[i].record()
        if self.n > 0:
            self._read((self.n - 1) & 1)
        self.n += 1

    def _read(self, j):
        self.ev[j].synchronize()
        self.values.append(float(self.host[j]))

    def flush(self):
        if self.n > len(self.values):
            self._read((self.n - 1) & 1)
        return self.values


class GraphedTrainStep:
    """Optimizer step replayed from CUDA graphs: at cfg 1 a step is ~370 kernel launches of 3-160 us each, so launch
    latency matters for the eager loop (and dominates it when the teacher-forced calls are not batched).

    buckets == 1: two graphs are captured once - (1) forward + backward + gradient gather, (2) AdamW - and the NCCL gradient
    all-reduce runs eagerly between them (collectives are kept out of capture on purpose: replicas replay independently and
    a captured collective can deadlock against host-side scheduling).

    buckets == K > 1 (data parallel): backward is cut at K-1 block boundaries into K stages, each its own graph
    (`torch.autograd.grad` from one boundary activation to the next).  Stage k's parameter gradients are a contiguous range
    of the flat buffer (`FlatGradients` is re-laid out in stage order) whose all-reduce is launched asynchronously as soon as
    the stage's graph is enqueued, so it runs beside the remaining stages; only the last bucket's all-reduce is exposed.

    Inputs live in static device buffers (`load` copies a host or device batch in, asynchronously).  libtbns kernels are
    plain stream launches on caller-owned memory, so they capture like any other kernel: TMA descriptors are kernel
    parameters and are frozen into the graph together with the static buffer addresses.  The optimizer must be built with
    capturable=True and a tensor lr (the host-side scheduler writes the new lr into that tensor after each replay)."""

    def __init__(self, model, optimizer, scheduler, grads: FlatGradients, example, T: int, step: int = 1,
                 batched: bool = True, warmup: int = 3, buckets: int = 1, loss_fn: Optional[Callable] = None):
        """loss_fn(model, x, fx, yy) -> scalar loss; default: the teacher-forced step loss of exp_ns.py (step_loss).  Pass
        `lambda m, x, fx, yy: unrolled_step_loss(m, x, fx, yy, T, step)` with a SOL_... model for the unrolled driver."""
        self.model, self.opt, self.sched, self.grads = model, optimizer, scheduler, grads
        self.T, self.step_, self.batched = T, step, batched
        self.loss_fn = loss_fn or (lambda m, x, fx, yy: step_loss(m, x, fx, yy, T, step, batched))
        self.static = tuple(torch.empty_like(t, device=next(model.parameters()).device) for t in example)
        self.load(example)
        blocks = list(getattr(model, "blocks", None) or getattr(getattr(model, "transolver_model", None), "blocks", []))
        self.nb = max(1, min(int(buckets), len(blocks)))
        if self.nb > 1:
            self._plan_stages(blocks)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # warm-up outside capture: allocator pools, packed-weight caches, smem opt-ins
                if self.nb > 1:
                    for k in range(self.nb):
                        self._stage(k)
                    self._release()
                else:
                    self._fwd_bwd()
                self.grads.all_reduce()
                self.opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_stage = []
        from . import ops
        self._ops = ops
        ops.begin_capture()   # derived weight copies (bf16 casts, packed projections) are re-recorded inside the graphs
        if self.nb > 1:
            pool = None
            for k in range(self.nb):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    out = self._stage(k)
                    if k == 0:
                        self.loss = out
                pool = g.pool()
                self.g_stage.append(g)
            self._release()
            for hk in self._hooks:      # only needed while the stages were traced: later eager forwards must not be retained
                hk.remove()
            self._hooks = []
            self.g_fb = self.g_stage[0]
        else:
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb):
                self.loss = self._fwd_bwd()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            self.opt.step()
        ops.invalidate_weight_caches()

    # ---- staged backward (buckets > 1) -------------------------------------------------------------------------------
    def _plan_stages(self, blocks):
        n = len(blocks)
        cuts = sorted({round(n * k / self.nb) for k in range(1, self.nb)} - {0, n})   # stage boundaries (block indices)
        self.nb = len(cuts) + 1
        self.cuts = cuts
        # stage 0 runs first in backward: the blocks after the last cut; the last stage owns blocks before the first cut
        # and everything that is not a block (preprocess, placeholder, ...)
        edges = [0] + cuts + [n]
        in_blocks = set()
        stage_params = []
        for k in range(self.nb):
            lo, hi = edges[self.nb - 1 - k], edges[self.nb - k]
            ps = [p for b in blocks[lo:hi] for p in b.parameters() if p.requires_grad]
            in_blocks.update(id(p) for p in ps)
            stage_params.append(ps)
        stage_params[-1] = [p for p in self.model.parameters() if p.requires_grad and id(p) not in in_blocks] + stage_params[-1]
        self.stage_params = stage_params
        # flat gradient buffer in stage order: each stage is one contiguous range
        fg = self.grads
        fg.params = [p for ps in stage_params for p in ps]
        off, fg.views, self.ranges = 0, [], []
        for ps in stage_params:
            start = off
            for p in ps:
                fg.views.append(fg.flat[off:off + p.numel()].view_as(p))
                off += fg.padded(p.numel())
            self.ranges.append((start, off))
        assert off == fg.flat.numel()
        for p, v in zip(fg.params, fg.views):
            p.grad = v
        self.stage_views = []
        i = 0
        for ps in stage_params:
            self.stage_views.append(fg.views[i:i + len(ps)])
            i += len(ps)
        # boundary activations: the input of blocks[c] is the output of blocks[c-1]
        self._acts = {}
        self._hooks = [blocks[c - 1].register_forward_hook(lambda mod, inp, out, c=c: self._acts.__setitem__(c, out)) for c in cuts]

    def _stage(self, k):
        """backward stage k (k = 0 also runs the forward).  Returns the detached loss for k == 0."""
        out = None
        if k == 0:
            x, fx, yy = self.static
            self._loss_t = self.loss_fn(self.model, x, fx, yy)
            out = self._loss_t.detach()
            root, gout = self._loss_t, None
        else:
            root, gout = self._acts[self.cuts[self.nb - 1 - (k - 1) - 1]], self._gact
        last = k == self.nb - 1
        params = self.stage_params[k]
        inputs = list(params) if last else [self._acts[self.cuts[self.nb - 2 - k]]] + list(params)
        res = torch.autograd.grad(root, inputs, grad_outputs=gout, retain_graph=not last, allow_unused=True)
        if not last:
            self._gact, res = res[0], res[1:]
        src, dst = [], []
        for g, v in zip(res, self.stage_views[k]):
            if g is None:
                v.zero_()
            else:
                src.append(g)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        return out

    def _release(self):
        self._loss_t, self._gact = None, None
        self._acts.clear()

    # ---- single-graph backward (buckets == 1) ------------------------------------------------------------------------
    def _fwd_bwd(self):
        x, fx, yy = self.static
        self.grads.begin()
        loss = self.loss_fn(self.model, x, fx, yy)
        loss.backward()
        self.grads.finish()
        return loss.detach()

    def load(self, batch):
        for dst, src in zip(self.static, batch):
            dst.copy_(src, non_blocking=True)

    def __call__(self, batch=None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        if self.nb > 1:
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not _NO_ALLREDUCE
            works = []
            for k, g in enumerate(self.g_stage):
                g.replay()
                if multi:   # this bucket's all-reduce runs beside the remaining backward stages
                    a, b = self.ranges[k]
                    works.append(dist.all_reduce(self.grads.flat[a:b], op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()
        else:
            self.g_fb.replay()
            self.grads.all_reduce()
        if hasattr(self.opt, "prepare_step"):
            self.opt.prepare_step()   # FlatAdamW: this step's lr / betas / bias corrections into the device array the graph reads
        self.g_opt.replay()
        self._ops.invalidate_weight_caches()   # the replay changed the masters without bumping their version counters
        if self.sched is not None:
            self.sched.step()     # host-side schedule; writes the new lr into the device tensor the graph reads
        return self.loss


class GraphedRollout:
    """closed-loop rollout (train.rollout) of a fixed shape replayed from ONE CUDA graph: T model calls whose inputs are
    strided windows of the static frame history and whose outputs land in its next column.  `load` copies a (host or
    device) batch in; `__call__` replays and returns the [B, N, T] view of the predictions inside the static history."""

    def __init__(self, model, example, T: int, step: int = 1, warmup: int = 2):
        from . import ops
        x, fx = example
        dev = next(model.parameters()).device
        self.model, self.T, self.step_, self.T_in = model, T, step, fx.shape[-1]
        self.x = torch.empty_like(x, device=dev)
        self.hist = torch.zeros(fx.shape[0], fx.shape[1], self.T_in + T, device=dev, dtype=torch.float32)
        self.load((x, fx))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.begin_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    @torch.no_grad()
    def _run(self):
        for t in range(0, self.T, self.step_):
            self.model(self.x, fx=self.hist[..., t:t + self.T_in], out=self.hist[..., self.T_in + t:self.T_in + t + self.step_])

    def load(self, batch):
        x, fx = batch
        self.x.copy_(x, non_blocking=True)
        self.hist[..., :self.T_in].copy_(fx, non_blocking=True)

    def __call__(self, batch=None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.hist[..., self.T_in:]
