#!/usr/bin/env python
"""bench.py — NS-64x64 Transolver training throughput (BASELINE.json metric) on N B200s of one node.

Workload (config.workload): BASELINE.json configs[1] — Transolver_Structured_Mesh_2D, 64x64 grid, 8 layers, n_hidden 256,
8 heads, slice_num 32, unified_pos 1, T_in = T = 10, per-GPU batch 2 (scripts/Transolver_NS.sh), bf16 operand mode,
one optimizer step = exp_ns.py:191-218 (10 teacher-forced model calls, summed rel-L2, one backward, AdamW + OneCycleLR),
batch sharded over ranks, gradients summed with one NCCL all-reduce.  Synthetic data, random-init weights.

  python bench.py [--gpus N --steps K --warmup W]          -> one JSON line (ours)
  python bench.py --impl reference [...]                   -> one JSON line: the reference algorithm (oracle port, torch CPU ops)
                                                              timed on this box's host cores (kind "port": the reference is
                                                              Python and cannot travel to the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(space_dim=2, n_layers=8, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
           slice_num=32, ref=8, unified_pos=1, H=64, W=64)
T_IN, T_OUT, STEP, PER_GPU_BATCH = 10, 10, 1, 2
WORKLOAD = "Transolver_Structured_Mesh_2D NS 64x64, 8 layers, n_hidden 256, 8 heads, slice_num 32, T_in=T_out=10, per-GPU batch 2"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """start of the timed region (nvidia-smi itself is started before the warm-up: it needs up to a second to come up)"""
        self.t0 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        window = "timed region"
        rows = [r for t, r in self.rows if self.t0 is None or t >= self.t0]
        if not any(len(r) >= 6 and r[0].replace(".", "").isdigit() for r in rows):
            rows, window = [r for _, r in self.rows], "warm-up + timed region (same load)"   # region shorter than one sample period
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def conv_fprop_flops(batch_tokens: int) -> float:
    C = CFG["n_hidden"]
    return 2.0 * batch_tokens * (9 * C) * (2 * C)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port on host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int):
    """each 'step' = ONE of the 10 teacher-forced model calls of an optimizer step (forward + backward, B=2, fp32) through
    the oracle restatement with torch CPU ops on all host threads; samples/s = B / (10 * t_call)."""
    from oracle import model as OM, physics_attention as O
    O.USE_LIBRARY_CONV = True  # same library conv as the reference (nn.Conv2d), see oracle/physics_attention.py
    torch.manual_seed(0)
    n_threads = torch.get_num_threads()
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
    m = Model(**CFG)  # parameter container only (CPU); compute below is the oracle's
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    N = CFG["H"] * CFG["W"]
    x = torch.rand(PER_GPU_BATCH, N, 2, generator=g)
    fx = 0.38 * torch.randn(PER_GPU_BATCH, N, T_IN, generator=g)
    y = 0.38 * torch.randn(PER_GPU_BATCH, N, 1, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = OM.model_forward(x, fx, sd, CFG["n_layers"], CFG["n_head"], (CFG["H"], CFG["W"]), True, CFG["ref"])
        loss = O.rel_l2_sum(out, y)
        torch.autograd.grad(loss, list(sd.values()), allow_unused=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t_call = statistics.median(times)
    calls = T_OUT // STEP
    return PER_GPU_BATCH / (calls * t_call), t_call, n_threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, t_call, cores = cpu_reference_run(args.steps, args.warmup)
    sample = f"{args.steps} timed x one teacher-forced model call fwd+bwd (B=2, fp32, oracle port, torch CPU ops); step = 10 such calls"
    line = {
        "impl": "reference", "metric": "NS-64x64 train samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_call * (T_OUT // STEP), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": PER_GPU_BATCH, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import ops, train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg.set_default_precision(args.precision)

    torch.manual_seed(1234)
    model = Model(**CFG).to(dev)
    train.broadcast_parameters(model)
    grads = train.FlatGradients(model.parameters())
    graphed = bool(args.graph)
    if args.precision == "bf16":
        # the preprocess MLP (outside the Physics-Attention path, still PyTorch) may use TF32 tensor cores in bf16 mode
        torch.backends.cuda.matmul.allow_tf32 = True
    opt = torch.optim.AdamW(model.parameters(), lr=torch.tensor(1e-3, device=dev) if graphed else 1e-3, weight_decay=1e-5,
                            fused=True, capturable=graphed)
    total_steps = 2 * (args.steps + args.warmup) + 16
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=total_steps)

    h = CFG["H"]
    # a pool of distinct pinned host batches (fresh data every step; rank-specific shard of the global batch)
    pool = [train.synthetic_ns_batch(PER_GPU_BATCH, h, T_IN, T_OUT, seed=1000 * rank + i, pin=True) for i in range(4)]
    dev_pool = [tuple(t.to(dev) for t in b) for b in pool]
    batched = bool(args.batched)
    gstep = None
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()   # long before the timed region: nvidia-smi needs up to a second for its first sample
    if graphed:
        ops.LAUNCHES = 0
        gstep = train.GraphedTrainStep(model, opt, sched, grads, dev_pool[0], T_OUT, STEP, batched=batched, warmup=3,
                                       buckets=args.buckets if world > 1 else 1)
        launches_per_step = ops.LAUNCHES // 4   # 3 eager warm-ups + 1 capture pass
        torch.cuda.synchronize()

    def step_device(i):
        x, fx, yy = dev_pool[i % len(dev_pool)]
        if graphed:
            return gstep((x, fx, yy))           # device->device copy into the static buffers + graph replay
        return train.train_step(model, opt, sched, grads, x, fx, yy, T_OUT, STEP, batched=batched)

    losses = train.DeferredLoss()

    def step_e2e(i):
        if graphed:
            # pinned host -> static device buffers, replay, D2H of this step's loss (read by the host one step late, so the
            # next step is already enqueued while it waits; the last one is read by `finish` inside the timed region)
            losses.push(gstep(pool[i % len(pool)]))
            return None
        x, fx, yy = (t.to(dev, non_blocking=True) for t in pool[i % len(pool)])
        loss = train.train_step(model, opt, sched, grads, x, fx, yy, T_OUT, STEP, batched=batched)
        return float(loss.item())  # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None, finish=None):
        barrier()
        if sampler:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.LAUNCHES = 0
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if graphed:
            ops.LAUNCHES = launches_per_step * steps   # kernels inside the replayed graph
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.LAUNCHES, clocks

    for i in range(args.warmup):
        step_device(i)
    # device-resident timing, with CUDA-event pairs around the dominant kernel (projection conv fprop)
    if not graphed:
        ops.PROFILE = {}
    ms_dev, launches, clocks = timed(step_device, args.steps, sampler)
    prof, ops.PROFILE = ops.PROFILE, None
    prof_ms = ms_dev
    if graphed:
        # kernels inside a replayed graph cannot be bracketed by events: time the dominant kernel in eager replays of the
        # same step (same shapes, same buffers) right after the timed region
        ops.PROFILE = {}
        side_saved, ops._USE_SIDE = ops._USE_SIDE, False   # per-launch times of kernels running alone, not beside the wgrad branch
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for i in range(3):
            x, fx, yy = dev_pool[i % len(dev_pool)]
            train.train_step(model, opt, None, grads, x, fx, yy, T_OUT, STEP, batched=batched)
        pe1.record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        ops._USE_SIDE = side_saved
        prof_ms = pe0.elapsed_time(pe1)
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    losses.flush()
    n_before = len(losses.values)
    ms_e2e, _, _ = timed(step_e2e, args.steps, finish=losses.flush)
    if graphed:
        assert len(losses.values) - n_before == args.steps and all(v == v for v in losses.values), "every step's loss must reach the host"

    replicas_identical = None
    if world > 1:
        # data-parallel invariant: every reduction on the path is fixed-order, so replicas must stay BITWISE identical
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        digest = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::97].double().sum()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        replicas_identical = all(bool(torch.equal(gathered[0], t)) for t in gathered)
    if rank == 0:
        pk = peaks()
        gb = PER_GPU_BATCH * world
        value = gb * args.steps / (ms_dev / 1e3)
        e2e = gb * args.steps / (ms_e2e / 1e3)
        calls = T_OUT // STEP
        tokens_per_launch = PER_GPU_BATCH * h * h * (calls if batched else 1)
        roof = None
        if prof.get("proj_fprop"):
            durs = [a.elapsed_time(b) for a, b in prof["proj_fprop"]]
            avg_ms = sum(durs) / len(durs)
            ach = conv_fprop_flops(tokens_per_launch) / (avg_ms / 1e3) / 1e12
            roof = {"kernel": "projection conv3x3 fprop (implicit GEMM, x|fx fused)", "bound": "tensor", "achieved": ach,
                    "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, one `ncu --set full` capture
                    # (profiles/ncu_r01_conv_fprop_persistent.md); algorithmic minimum 2*M*C + 4*M*2I + 2*9C*2I = 212 MB
                    "traffic": 160.65e6 if batched else None, "traffic_unit": "bytes/launch (ncu)",
                    "peak_source": pk["source"] + " (sustained bf16 cuBLAS)", "avg_launch_ms": avg_ms, "launches_timed": len(durs),
                    # warm launch time per step / graph-timed step time (eager replays only provide the per-launch durations)
                    "share_of_step": (sum(durs) / (3 if graphed else args.steps)) / (ms_dev / args.steps),
                    "timed": "eager replays after the graph-timed region" if graphed else "inside the timed region"}
            for tag in ("proj_dgrad", "proj_wgrad"):
                if prof.get(tag):
                    d2 = [a.elapsed_time(b) for a, b in prof[tag]]
                    roof[tag + "_share_of_step"] = (sum(d2) / (3 if graphed else args.steps)) / (ms_dev / args.steps)
            # warm per-kernel-family device time per step (CUDA events around every tagged libtbns launch, eager replays)
            nprof = 3 if graphed else args.steps
            roof["kernel_ms_per_step"] = {t: round(sum(a.elapsed_time(b) for a, b in v) / nprof, 3) for t, v in sorted(prof.items())}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, t_call, cores = cpu_reference_run(3, 1)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "3 timed x one teacher-forced model call fwd+bwd (B=2, fp32, oracle port on host cores); step = 10 calls"}
        per_step_in = sum(t.numel() * t.element_size() for t in pool[0])
        line = {
            "metric": "NS-64x64 train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": gb, "parallelism": f"dp{world}",
                       "teacher_forced_calls_batched": batched, "cuda_graph": graphed, "allreduce_buckets": (gstep.nb if gstep is not None else 1),
                       "replicas_identical": replicas_identical,
                       "l2": "activations written per step (>1 GB) exceed the 126 MB L2; fresh input batch every step"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": per_step_in, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "loss_readback": ("every step's loss through pinned host memory, read one step late (train.DeferredLoss); all "
                                      "K reads inside the timed region") if graphed else "loss.item() every step"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batched", type=int, default=1, help="evaluate the 10 teacher-forced calls as one batch (same math)")
    ap.add_argument("--graph", type=int, default=1, help="replay the optimizer step from CUDA graphs (fwd+bwd graph, eager NCCL "
                    "all-reduce, optimizer graph); 0 = eager launches")
    ap.add_argument("--buckets", type=int, default=1, help="N > 1 GPUs: cut backward into this many stage graphs and all-reduce each "
                    "stage's gradient range beside the next stages (measured on 2 GPUs: the 44.8 MB exchange costs 0.07 ms of a "
                    "10.8 ms step, less than the extra graph launches, so the default is one bucket)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
