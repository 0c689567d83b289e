"""Reported baselines for bench.py: the UNMODIFIED reference model (oracle/_ref or /root/reference via reference_shim)
driven through the training step of exp_ns.py:191-218 - on the host cores (`--impl reference`, `cpu_baseline`) and, as the
kernel-level bar SURVEY.md §2.2 / §8d asks for, as stock PyTorch eager on one B200 (`gpu_eager_baseline`).

BASELINE INFRASTRUCTURE ONLY: nothing here is on the product path; the product package never imports it.
"""
from __future__ import annotations

import os
import statistics
import time

import torch

from . import reference_shim as R


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _loss(pred, y):
    """utils/testloss.py:31-42, size_average=False"""
    n = pred.shape[0]
    d = torch.linalg.vector_norm(pred.reshape(n, -1) - y.reshape(n, -1), dim=1)
    return (d / torch.linalg.vector_norm(y.reshape(n, -1), dim=1)).sum()


def build_reference_model(cfg: dict, device, seed: int = 0):
    torch.manual_seed(seed)
    with R.cpu_cuda_identity(force=(torch.device(device).type == "cpu")):
        m = R.transolver_2d().Model(**cfg)
    m = m.to(device)
    if hasattr(m, "pos") and torch.is_tensor(m.pos):
        m.pos = m.pos.to(device)   # a plain attribute in the reference (not a buffer): .to() does not move it
    return m


def reference_step(model, opt, sched, x, fx, yy, T: int, step: int, batched: bool, autocast_dtype=None):
    """one optimizer step with exp_ns.py:191-218 semantics through the reference's own Model.forward.
    batched=False: the literal loop.  batched=True: the ten teacher-forced inputs stacked on the batch axis (the same
    restructuring bench.py's own arm uses; same math)."""
    bsz = x.shape[0]
    dev_type = x.device.type
    ctx = torch.autocast(dev_type, dtype=autocast_dtype) if autocast_dtype is not None else torch.autocast(dev_type, enabled=False)
    with ctx:
        if batched:
            calls = T // step
            full = torch.cat((fx, yy), -1)
            T_in = fx.shape[-1]
            fx_all = torch.cat([full[..., t:t + T_in] for t in range(0, T, step)], 0)
            y_all = torch.cat([yy[..., t:t + step] for t in range(0, T, step)], 0)
            im = model(x.repeat(calls, 1, 1), fx=fx_all)
            loss = _loss(im.float().reshape(calls * bsz, -1), y_all.reshape(calls * bsz, -1))
        else:
            loss = 0
            for t in range(0, T, step):
                y = yy[..., t:t + step]
                im = model(x, fx=fx)
                loss = loss + _loss(im.float().reshape(bsz, -1), y.reshape(bsz, -1))
                fx = torch.cat((fx[..., step:], y), dim=-1)
    opt.zero_grad()
    loss.backward()
    opt.step()
    if sched is not None:
        sched.step()
    return loss.detach()


def cpu_reference_steps(cfg: dict, batch: int, T_in: int, T: int, step: int, steps: int, warmup: int, threads: int | None = None):
    """-> dict(value samples/s, ms_per_step, cores, kind, times).  Real optimizer steps (ten calls, one backward, AdamW) of
    the reference model on the host cores; `threads` defaults to every core this process may use (torchrun's
    OMP_NUM_THREADS=1 is deliberately overridden)."""
    cores = threads or host_cores()
    torch.set_num_threads(cores)
    m = build_reference_model(cfg, "cpu")
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=steps + warmup + 4)
    g = torch.Generator().manual_seed(1)
    N = cfg["H"] * cfg["W"]
    times = []
    for i in range(warmup + steps):
        x = torch.rand(batch, N, 2, generator=g)
        fx = 0.38 * torch.randn(batch, N, T_in, generator=g)
        yy = 0.38 * torch.randn(batch, N, T, generator=g)
        t0 = time.perf_counter()
        reference_step(m, opt, sched, x, fx, yy, T, step, batched=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                kind="reference" if R.available() else "port", median_ms=1e3 * statistics.median(times))


def gpu_eager_baseline(cfg: dict, batch: int, T_in: int, T: int, step: int, device, steps: int = 5, warmup: int = 2):
    """the reference model as stock PyTorch eager (cuDNN / cuBLAS) on one GPU, same optimizer step: fp32 with TF32 tensor
    cores and bf16 autocast, literal ten-call loop and batched.  -> {mode: {samples_per_s, ms_per_step}}"""
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    try:
        N = cfg["H"] * cfg["W"]
        g = torch.Generator().manual_seed(2)
        x = torch.rand(batch, N, 2, generator=g).to(device)
        fx = (0.38 * torch.randn(batch, N, T_in, generator=g)).to(device)
        yy = (0.38 * torch.randn(batch, N, T, generator=g)).to(device)
        for name, dtype in (("tf32", None), ("bf16_autocast", torch.bfloat16)):
            for batched in (False, True):
                m = build_reference_model(cfg, device)
                opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
                for _ in range(warmup):
                    reference_step(m, opt, None, x, fx, yy, T, step, batched, dtype)
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    reference_step(m, opt, None, x, fx, yy, T, step, batched, dtype)
                e1.record()
                torch.cuda.synchronize(device)
                ms = e0.elapsed_time(e1) / steps
                out[f"{name}_{'batched' if batched else 'literal'}"] = {"samples_per_s": batch / (ms / 1e3), "ms_per_step": ms}
                del m, opt
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out
