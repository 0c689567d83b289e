#!/usr/bin/env python
"""bench.py — NS-64x64 Transolver training throughput (BASELINE.json metric) on N B200s of one node.

Default workload (config.workload): BASELINE.json configs[1] — Transolver_Structured_Mesh_2D, 64x64 grid, 8 layers,
n_hidden 256, 8 heads, slice_num 32, unified_pos 1, T_in = T = 10, per-GPU batch 2 (scripts/Transolver_NS.sh), bf16 operand
mode, one optimizer step = exp_ns.py:191-218 (10 teacher-forced model calls, summed rel-L2, one backward, AdamW +
OneCycleLR), batch sharded over ranks, gradients summed with one NCCL all-reduce.  Synthetic data, random-init weights.

  python bench.py [--gpus N --steps K --warmup W]          -> one JSON line (ours)
  python bench.py --impl reference [...]                   -> one JSON line: the UNMODIFIED reference model (oracle/_ref,
                                                              byte-compiled by oracle/build_ref.py) running real optimizer steps
                                                              on this box's host cores
  python bench.py --workload unrolled --look-ahead L       -> ns_vorticity_unrolling.py:225-244 training step through
                                                              SOL_Transolver_Structured_Mesh_2D (north_star's driver)
  python bench.py --workload rollout_cfg5 [--gpus N]       -> BASELINE configs[4]: 256x256 grid, 16 layers, slice_num 64,
                                                              closed-loop rollout inference, batch sharded over ranks
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
# rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION / INFO write to stdout) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("TBNS_KEEP_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = "WARN"

CFG = dict(space_dim=2, n_layers=8, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
           slice_num=32, ref=8, unified_pos=1, H=64, W=64)
CFG5 = dict(space_dim=2, n_layers=16, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
            slice_num=64, ref=8, unified_pos=1, H=256, W=256)
T_IN, T_OUT, STEP, PER_GPU_BATCH = 10, 10, 1, 2
WORKLOAD = "Transolver_Structured_Mesh_2D NS 64x64, 8 layers, n_hidden 256, 8 heads, slice_num 32, T_in=T_out=10, per-GPU batch 2"
WORKLOAD5 = "Transolver_Structured_Mesh_2D NS 256x256, 16 layers, n_hidden 256, 8 heads, slice_num 64, 10-step closed-loop rollout"

# kernels of the Physics-Attention module itself (forward + backward) among the CUDA-event tags of ops.py
PA_TAGS = ("proj_fprop", "proj_dgrad", "proj_wgrad", "slice_fwd", "slice_bwd", "token_attn_fwd", "token_attn_bwd", "deslice_out",
           "deslice_dw", "deslice_dP")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def ncu_traffic(kernel_key: str):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written from the ncu report by profiles/summarize_ncu.py) - None when no capture of the shipped binary exists"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except (ValueError, OSError):
            return None
    return None


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


# ------------------------------------------------------------------------------------------------
# algorithmic work (SURVEY.md §8d closed forms, per sample per layer, forward; backward = 2x)
# ------------------------------------------------------------------------------------------------
def pa_flops_fwd(cfg) -> float:
    N, C, H, G = cfg["H"] * cfg["W"], cfg["n_hidden"], cfg["n_head"], cfg["slice_num"]
    D = C // H
    return 2 * 2 * N * 9 * C * C + 3 * (2 * N * C * G) + 6 * H * G * D * D + 4 * H * G * G * D + 2 * N * C * C


def model_flops_fwd(cfg, in_features=74) -> float:
    N, C = cfg["H"] * cfg["W"], cfg["n_hidden"]
    per_layer = pa_flops_fwd(cfg) + 4 * N * cfg["mlp_ratio"] * C * C
    return cfg["n_layers"] * per_layer + 2 * N * (in_features * 2 * C + 2 * C * C) + 2 * N * C * cfg["out_dim"]


def conv_fprop_flops(batch_tokens: int) -> float:
    C = CFG["n_hidden"]
    return 2.0 * batch_tokens * (9 * C) * (2 * C)


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference model on the host cores (oracle/baselines.py)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return   # under torchrun rank 0 alone measures; the other ranks exit without work
    from oracle import baselines as BL
    from oracle import reference_shim as R
    r = BL.cpu_reference_steps(CFG, PER_GPU_BATCH, T_IN, T_OUT, STEP, steps=args.steps, warmup=args.warmup)
    one = None
    if not args.no_cpu_1thread:
        # 1-thread figure on a bounded sample: ONE real optimizer step would take minutes single-threaded, so one model call
        # forward + backward (1/10 of a step's model work) is timed and scaled; stated as such
        torch.set_num_threads(1)
        m = BL.build_reference_model(CFG, "cpu")
        g = torch.Generator().manual_seed(5)
        N = CFG["H"] * CFG["W"]
        x, fx = torch.rand(PER_GPU_BATCH, N, 2, generator=g), 0.38 * torch.randn(PER_GPU_BATCH, N, T_IN, generator=g)
        y = 0.38 * torch.randn(PER_GPU_BATCH, N, 1, generator=g)
        t0 = time.perf_counter()
        BL._loss(m(x, fx=fx).reshape(PER_GPU_BATCH, -1), y.reshape(PER_GPU_BATCH, -1)).backward()
        t_call = time.perf_counter() - t0
        one = {"value": PER_GPU_BATCH / (10 * t_call), "unit": "samples/s", "cores": 1,
               "sample": "one model call forward+backward (B=2) x 10 calls per step, extrapolated"}
        torch.set_num_threads(r["cores"])
    sample = (f"{args.steps} timed real optimizer steps (exp_ns.py:191-218: 10 teacher-forced calls, one backward, AdamW + OneCycleLR), "
              f"B=2, fp32, reference modules from {R.kind()} ({'oracle/_ref' if R.kind() == 'compiled' else R.REF_ROOT})")
    line = {
        "impl": "reference", "metric": "NS-64x64 train samples/sec", "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": PER_GPU_BATCH, "parallelism": "cpu"},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": sample,
                         "one_thread": one},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Harness:
    """process-group / timing plumbing shared by the workloads"""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None
        if self.sampler:
            self.sampler.start()   # long before the timed region: nvidia-smi needs up to a second for its first sample

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, sample=False, finish=None):
        """EXACTLY `steps` calls bracketed by barrier + synchronize on both sides, CUDA events, MAX over ranks -> ms"""
        self.barrier()
        if sample and self.sampler:
            self.sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        clocks = self.sampler.stop() if (sample and self.sampler) else None
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def profile_eager_steps(step_fn, n=3):
    """CUDA-event pairs around every tagged libtbns launch during `n` eager steps, side stream off so that each tagged kernel
    runs alone (kernels inside a replayed graph cannot be bracketed by events) -> {tag: ms per step}, raw event pairs"""
    from transformerbasednavierstokesolver_b200 import ops
    side_saved, ops._USE_SIDE = ops._USE_SIDE, False
    try:
        # one untimed eager replay first: the graphed region ran from a private memory pool, so the first eager step pays for
        # allocator growth (host-synchronous cudaMalloc) and cold caches; kernels are timed after warm-up like everything else
        ops.PROFILE = None
        step_fn(0)
        torch.cuda.synchronize()
        ops.PROFILE = {}
        for i in range(n):
            # keep the stream backlogged while the host enqueues the eager step (~100 ms spin kernel first): an event recorded
            # on an idle stream is stamped at enqueue time, so host launch latency between the two records of a pair would be
            # counted as kernel time
            torch.cuda._sleep(int(0.1 * 1.9e9))
            step_fn(i)
        torch.cuda.synchronize()
    finally:
        prof, ops.PROFILE = ops.PROFILE, None
        ops._USE_SIDE = side_saved
    def robust(v):
        # the launch sequence of a tag is the same in every replay: launch j's duration = median over the n replays (a host
        # hiccup that lets the stream run dry inflates single samples), summed over j
        d = [a.elapsed_time(b) for a, b in v]
        if len(d) % n:
            return sum(d) / n
        m = len(d) // n
        return sum(sorted(d[j + s * m] for s in range(n))[n // 2] for j in range(m))
    per_step = {t: robust(v) for t, v in sorted(prof.items())}
    dump = os.environ.get("TBNS_BENCH_DUMP")   # debugging aid: per-launch durations (us) of one tag on stderr
    if dump and dump in prof:
        print(dump, [round(a.elapsed_time(b) * 1e3, 1) for a, b in prof[dump]], file=sys.stderr)
    return per_step, prof


def run_train(args, unrolled: bool):
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import ops, train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model
    from transformerbasednavierstokesolver_b200.model.SOL_Transolver_Structured_Mesh_2D import SOL_Transolver_Structured_Mesh_2D

    hs = Harness(args)
    dev, world, rank = hs.dev, hs.world, hs.rank
    dist = hs.dist
    pkg.set_default_precision(args.precision)
    torch.manual_seed(1234)
    batched = bool(args.batched)
    if unrolled:
        model = SOL_Transolver_Structured_Mesh_2D(**CFG, step=STEP, look_ahead=args.look_ahead).to(dev)
        loss_fn = lambda m, x, fx, yy: train.unrolled_step_loss(m, x, fx, yy, T_OUT, STEP, batched)   # noqa: E731
        calls = args.look_ahead * len(range(0, T_OUT - args.look_ahead * STEP + 1, args.look_ahead * STEP))
    else:
        model = Model(**CFG).to(dev)
        loss_fn = lambda m, x, fx, yy: train.step_loss(m, x, fx, yy, T_OUT, STEP, batched)   # noqa: E731
        calls = T_OUT // STEP
    train.broadcast_parameters(model)
    grads = train.FlatGradients(model.parameters())
    graphed = bool(args.graph)
    if args.optimizer == "flat" and not (world > 1 and args.buckets > 1):
        # AdamW as one libtbns kernel over flat parameter / gradient / moment buffers (same update rule as torch.optim.AdamW)
        opt = train.FlatAdamW(model.parameters(), grads, lr=1e-3, weight_decay=1e-5)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=torch.tensor(1e-3, device=dev) if graphed else 1e-3, weight_decay=1e-5,
                                fused=True, capturable=graphed)
    total_steps = 2 * (args.steps + args.warmup) + 32
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=total_steps)

    h = CFG["H"]
    # a pool of distinct pinned host batches (fresh data every step; rank-specific shard of the global batch)
    pool = [train.synthetic_ns_batch(PER_GPU_BATCH, h, T_IN, T_OUT, seed=1000 * rank + i, pin=True) for i in range(4)]
    dev_pool = [tuple(t.to(dev) for t in b) for b in pool]
    gstep = None
    launches_per_step = None
    if graphed:
        ops.LAUNCHES = 0
        gstep = train.GraphedTrainStep(model, opt, sched, grads, dev_pool[0], T_OUT, STEP, batched=batched, warmup=3,
                                       buckets=args.buckets if world > 1 else 1, loss_fn=loss_fn)
        launches_per_step = ops.LAUNCHES // 4   # 3 eager warm-ups + 1 capture pass
        torch.cuda.synchronize()

    def eager_step(x, fx, yy, sched_=sched):
        grads.begin()
        loss = loss_fn(model, x, fx, yy)
        loss.backward()
        grads.finish()
        grads.all_reduce()
        opt.step()
        if sched_ is not None:
            sched_.step()
        return loss.detach()

    def step_device(i):
        x, fx, yy = dev_pool[i % len(dev_pool)]
        if graphed:
            return gstep((x, fx, yy))           # device->device copy into the static buffers + graph replay
        return eager_step(x, fx, yy)

    losses = train.DeferredLoss()

    def step_e2e(i):
        if graphed:
            # pinned host -> static device buffers, replay, D2H of this step's loss (read by the host one step late, so the
            # next step is already enqueued while it waits; the last one is read by `finish` inside the timed region)
            losses.push(gstep(pool[i % len(pool)]))
            return None
        x, fx, yy = (t.to(dev, non_blocking=True) for t in pool[i % len(pool)])
        return float(eager_step(x, fx, yy).item())  # D2H read of the step's result

    for i in range(args.warmup):
        step_device(i)
    ops.LAUNCHES = 0
    if not graphed:
        ops.PROFILE = {}
    ms_dev, clocks = hs.timed(step_device, args.steps, sample=True)
    launches = launches_per_step * args.steps if graphed else ops.LAUNCHES
    if graphed:
        kernel_ms, prof = profile_eager_steps(lambda i: eager_step(*dev_pool[i % len(dev_pool)], sched_=None), 3)
        nprof = 3
    else:
        prof, ops.PROFILE = ops.PROFILE, None
        nprof = args.steps
        kernel_ms = {t: sum(a.elapsed_time(b) for a, b in v) / nprof for t, v in sorted(prof.items())}
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    losses.flush()
    n_before = len(losses.values)
    ms_e2e, _ = hs.timed(step_e2e, args.steps, finish=losses.flush)
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

    eager_bar = None
    if rank == 0 and world == 1 and not unrolled and not args.no_eager_baseline:
        try:
            from oracle import baselines as BL
            eager_bar = BL.gpu_eager_baseline(CFG, PER_GPU_BATCH, T_IN, T_OUT, STEP, dev)
            eager_bar["what"] = ("the UNMODIFIED reference model as stock PyTorch eager (cuDNN / cuBLAS) on this GPU, same optimizer step "
                                 "(exp_ns.py:191-218, AdamW): fp32 with TF32 tensor cores and bf16 autocast, literal ten-call loop and "
                                 "the ten calls batched; 5 timed steps after 2 warm-up")
        except Exception as e:   # a reported baseline must not take the measurement down
            eager_bar = {"unavailable": repr(e)[:200]}

    if rank == 0:
        pk = peaks()
        gb = PER_GPU_BATCH * world
        step_ms = ms_dev / args.steps
        value = gb * args.steps / (ms_dev / 1e3)
        e2e = gb * args.steps / (ms_e2e / 1e3)
        windows = calls // args.look_ahead if unrolled else calls
        imgs_per_launch = PER_GPU_BATCH * (windows if batched else 1)
        tokens_per_launch = imgs_per_launch * h * h
        roof = None
        if prof.get("proj_fprop"):
            durs = [a.elapsed_time(b) for a, b in prof["proj_fprop"]]
            avg_ms = sum(durs) / len(durs)
            ach = conv_fprop_flops(tokens_per_launch) / (avg_ms / 1e3) / 1e12
            roof = {"kernel": "projection conv3x3 fprop (implicit GEMM, x|fx fused)", "bound": "tensor", "achieved": ach,
                    # the kernel is timed ALONE in short eager replays at boost clocks: the burst figure is the honest denominator
                    "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": ach / pk["bf16_burst"],
                    "frac_of_sustained_peak": ach / pk["bf16_sustained"],
                    "traffic": ncu_traffic("proj_fprop"), "traffic_unit": "bytes/launch, dram read+write, ncu --set full (profiles/)",
                    "algorithmic_bytes_per_launch": tokens_per_launch * (2 * 256 + 2 * 512) + 2 * 2304 * 512,
                    "peak_source": pk["source"] + " (burst bf16 cuBLAS)", "avg_launch_ms": avg_ms, "launches_timed": len(durs),
                    "flops_per_launch": conv_fprop_flops(tokens_per_launch),
                    "share_of_step": (sum(durs) / nprof) / step_ms,
                    "timed": "eager replays after the graph-timed region, stream kept backlogged (event pairs bracket device time only)" if graphed else "inside the timed region"}
            for tag in ("proj_dgrad", "proj_wgrad"):
                if tag in kernel_ms:
                    roof[tag + "_share_of_step"] = kernel_ms[tag] / step_ms
            roof["kernel_ms_per_step"] = {t: round(v, 3) for t, v in kernel_ms.items()}
        # Physics-Attention as a whole (BASELINE metric 2 / north_star's 60 % bar): algorithmic FLOPs of the module, forward +
        # backward, over the device time of ITS kernels (each timed alone, eager replays)
        pa = None
        if kernel_ms:
            pa_ms = sum(v for t, v in kernel_ms.items() if t in PA_TAGS)
            pa_gflop = 3 * pa_flops_fwd(CFG) * CFG["n_layers"] * PER_GPU_BATCH * calls / 1e9
            if pa_ms > 0:
                pa_t = pa_gflop / pa_ms   # GFLOP / ms = TFLOP/s
                pa = {"pa_tflops": pa_t, "pa_frac_of_burst": pa_t / pk["bf16_burst"], "pa_kernel_ms_per_step": pa_ms,
                      "pa_gflop_per_step": pa_gflop,
                      "step_tflops": 3 * model_flops_fwd(CFG) * PER_GPU_BATCH * calls / 1e9 / step_ms,
                      "kernels": [t for t in PA_TAGS if t in kernel_ms]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not unrolled:
            from oracle import baselines as BL
            from oracle import reference_shim as R
            r = BL.cpu_reference_steps(CFG, PER_GPU_BATCH, T_IN, T_OUT, STEP, steps=2, warmup=1)
            cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": f"2 timed real optimizer steps after 1 warm-up (10 calls, one backward, AdamW), B=2, fp32, reference modules ({R.kind()})"}
        per_step_in = sum(t.numel() * t.element_size() for t in pool[0])
        metric = "NS-64x64 train samples/sec" if not unrolled else f"NS-64x64 unrolled (look_ahead {args.look_ahead}) train samples/sec"
        line = {
            "metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + (f"; ns_vorticity_unrolling step, look_ahead {args.look_ahead}" if unrolled else ""),
                       "global_batch": gb, "parallelism": f"dp{world}",
                       "calls_batched": batched, "images_per_launch": imgs_per_launch, "model_calls_per_step": calls,
                       "cuda_graph": graphed, "allreduce_buckets": (gstep.nb if gstep is not None else 1),
                       "optimizer": ("AdamW as one libtbns kernel over flat buffers (train.FlatAdamW)" if isinstance(opt, train.FlatAdamW)
                                     else "torch.optim.AdamW(fused=True)") + " + OneCycleLR",
                       "replicas_identical": replicas_identical,
                       "l2": "activations written per step (>1 GB) exceed the 126 MB L2; fresh input batch every step"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": per_step_in, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "loss_readback": ("every step's loss through pinned host memory, read one step late (train.DeferredLoss); all "
                                      "K reads inside the timed region") if graphed else "loss.item() every step"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "physics_attention": pa, "cpu_baseline": cpu,
            "gpu_eager_baseline": eager_bar,
        }
        print(json.dumps(line), flush=True)
    hs.close()


def run_rollout_cfg5(args):
    """BASELINE configs[4]: closed-loop rollout inference (ns_vorticity_unrolling.py:264-286) of the scaled NS model, batch
    sharded over ranks; the only exchange is the scalar error metric all-reduced at the end of each rollout."""
    import transformerbasednavierstokesolver_b200 as pkg
    from transformerbasednavierstokesolver_b200 import ops, train
    from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model

    hs = Harness(args)
    dev, world, rank = hs.dev, hs.world, hs.rank
    dist = hs.dist
    pkg.set_default_precision(args.precision)
    torch.manual_seed(4321)
    model = Model(**CFG5).to(dev).eval()
    train.broadcast_parameters(model)
    B, h = args.rollout_batch, CFG5["H"]
    pool = [train.synthetic_ns_batch(B, h, T_IN, T_OUT, seed=7000 * rank + i, pin=True) for i in range(2)]
    dev_pool = [tuple(t.to(dev) for t in b) for b in pool]
    graphed = bool(args.graph)
    ops.LAUNCHES = 0
    runner = train.GraphedRollout(model, dev_pool[0][:2], T_OUT, STEP, warmup=2) if graphed else None
    launches_per_step = ops.LAUNCHES // 3 if graphed else None
    metric_host = torch.empty((), dtype=torch.float32).pin_memory()
    last = {}

    def one(batch, read):
        x, fx, yy = batch
        pred = runner((x, fx)) if graphed else train.rollout(model, x.to(dev, non_blocking=True), fx.to(dev, non_blocking=True), T_OUT, STEP)
        err = train.rel_l2_sum(pred.reshape(B, -1), yy.to(dev, non_blocking=True).reshape(B, -1))   # test_l2_full, :200
        if world > 1:
            dist.all_reduce(err, op=dist.ReduceOp.SUM)     # the only collective of the sharded rollout
        if read:
            metric_host.copy_(err, non_blocking=True)
        last["err"] = err

    for i in range(args.warmup):
        one(dev_pool[i % 2], False)
    ops.LAUNCHES = 0
    ms_dev, clocks = hs.timed(lambda i: one(dev_pool[i % 2], False), args.steps, sample=True)
    launches = launches_per_step * args.steps if graphed else ops.LAUNCHES
    ms_e2e, _ = hs.timed(lambda i: one(pool[i % 2], True), args.steps, finish=torch.cuda.synchronize)
    kernel_ms, prof = profile_eager_steps(lambda i: train.rollout(model, dev_pool[0][0], dev_pool[0][1], T_OUT, STEP), 2)
    if rank == 0:
        pk = peaks()
        gb = B * world
        calls = T_OUT // STEP
        step_ms = ms_dev / args.steps
        frames = gb * calls * args.steps / (ms_dev / 1e3)
        flops_call = model_flops_fwd(CFG5) * B
        roof = None
        if prof.get("proj_fprop"):
            durs = [a.elapsed_time(b) for a, b in prof["proj_fprop"]]
            avg_ms = sum(durs) / len(durs)
            fl = 2.0 * B * h * h * 2304 * 512
            roof = {"kernel": "projection conv3x3 fprop (implicit GEMM, x|fx fused)", "bound": "tensor", "achieved": fl / (avg_ms / 1e3) / 1e12,
                    "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": fl / (avg_ms / 1e3) / 1e12 / pk["bf16_burst"], "traffic": None,
                    "avg_launch_ms": avg_ms, "launches_timed": len(durs), "kernel_ms_per_step": {t: round(v, 3) for t, v in kernel_ms.items()}}
        per_step_in = sum(t.numel() * t.element_size() for t in pool[0])
        line = {
            "metric": "NS-256x256 rollout frames/sec", "value": frames, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "ms_per_model_call": step_ms / calls, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD5, "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{world} (batch shards, metric all-reduce)",
                       "cuda_graph": graphed, "fused_rollout_step": True,
                       "l2": "each model call streams > 1 GB of activations through the 126 MB L2"},
            "e2e": {"value": gb * calls * args.steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": per_step_in,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "model_tflops": flops_call * calls / 1e9 / step_ms, "rollout_error_metric": float(last["err"]) / gb,
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    hs.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "unrolled", "rollout_cfg5"])
    ap.add_argument("--look-ahead", type=int, default=2, help="--workload unrolled: chained model calls per window (1, 2, 4, 8, 10)")
    ap.add_argument("--rollout-batch", type=int, default=2, help="--workload rollout_cfg5: samples per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_exact"])
    ap.add_argument("--batched", type=int, default=1, help="evaluate the teacher-forced calls / windows as one batch (same math)")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from CUDA graphs (fwd+bwd graph, eager NCCL all-reduce, "
                    "optimizer graph); 0 = eager launches")
    ap.add_argument("--buckets", type=int, default=1, help="N > 1 GPUs: cut backward into this many stage graphs and all-reduce each "
                    "stage's gradient range beside the next stages")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"], help="flat: train.FlatAdamW (one libtbns kernel over flat "
                    "buffers); torch: torch.optim.AdamW(fused=True)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-1thread", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        # a CPU step is 3+ s: two untimed steps are enough to page everything in (the line reports the W actually used)
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    elif args.workload == "rollout_cfg5":
        run_rollout_cfg5(args)
    else:
        run_train(args, unrolled=args.workload == "unrolled")


if __name__ == "__main__":
    main()
