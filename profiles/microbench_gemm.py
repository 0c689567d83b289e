"""Micro-benchmark of the tcgen05 GEMM kernels at the bench shapes (B=20 images of 64x64, C=256): CUDA-event timing of
isolated launches with an L2 flush between iterations.  usage: python profiles/microbench_gemm.py [iters]"""
import os
import sys
import torch
sys.path.insert(0, ".")
from transformerbasednavierstokesolver_b200 import ops

dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, Hg, Wg, C, I2 = 20, 64, 64, 256, 512
M = B * Hg * Wg
g = torch.Generator().manual_seed(0)
x16 = torch.randn(M, C, generator=g).to(dev).bfloat16()
Wf16 = (torch.randn(I2, 9 * C, generator=g) / 48).to(dev).bfloat16()
bias = torch.randn(I2, generator=g).to(dev)
XF = torch.empty(M, I2, device=dev)
XF16 = torch.empty(M, I2, device=dev, dtype=torch.bfloat16)
W1 = (torch.randn(C, C, generator=g) / 16).to(dev).bfloat16()
b1 = torch.randn(C, generator=g).to(dev)
pre16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
hid16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
res = torch.randn(M, C, generator=g).to(dev)
out = torch.empty(M, C, device=dev)
dXF16 = torch.randn(M, I2, generator=g).to(dev).bfloat16()
Wd16 = (torch.randn(C, 9 * I2, generator=g) / 68).to(dev).bfloat16()
dWx = torch.empty(C, C, 3, 3, device=dev)
dWfx = torch.empty(C, C, 3, 3, device=dev)
lng = torch.randn(C, generator=g).to(dev)
lnb = torch.randn(C, generator=g).to(dev)
PT16 = (torch.randn(B, C, C, generator=g) / 16).to(dev).bfloat16()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

cases = {
    "conv_fprop 81920x2304x512": (lambda: ops.gemm_tc(x16, Wf16, XF, bias, B, Hg, Wg, C, I2, 9, 0), 2.0 * M * 9 * C * I2),
    "conv_fprop16 81920x2304x512 bf16 out": (lambda: ops.gemm_tc(x16, Wf16, None, bias, B, Hg, Wg, C, I2, 9, 0, C16=XF16), 2.0 * M * 9 * C * I2),
    "conv_dgrad 81920x4608x256": (lambda: ops.gemm_tc(dXF16, Wd16, out, None, B, Hg, Wg, I2, C, 9, 1), 2.0 * M * 9 * C * I2),
    "fc1 81920x256x256 +bias+gelu+pre16+bf16": (lambda: ops.gemm_tc(x16, W1, None, b1, 1, 1, M, C, C, act=1, aux_out=pre16, aux_bf16=1, C16=hid16), 2.0 * M * C * C),
    "dpre 81920x256x256 gelu'(pre16)+bf16": (lambda: ops.gemm_tc(x16, W1, None, None, 1, 1, M, C, C, act=2, aux_in=pre16, aux_bf16=1, C16=hid16), 2.0 * M * C * C),
    "fc1d 81920x256x256 +bias+gelu+gelu'16+bf16 (act 3)": (lambda: ops.gemm_tc(x16, W1, None, b1, 1, 1, M, C, C, act=3, aux_out=pre16, aux_bf16=1, C16=hid16), 2.0 * M * C * C),
    "dpred 81920x256x256 *stored gelu'+bf16 (act 4)": (lambda: ops.gemm_tc(x16, W1, None, None, 1, 1, M, C, C, act=4, aux_in=pre16, aux_bf16=1, C16=hid16), 2.0 * M * C * C),
    "dx2 81920x256x256 fp32 out": (lambda: ops.gemm_tc(x16, W1, out, None, 1, 1, M, C, C), 2.0 * M * C * C),
    "fc2 81920x256x256 +bias+res": (lambda: ops.gemm_tc(hid16, W1, out, b1, 1, 1, M, C, C, residual=res), 2.0 * M * C * C),
    "fc2ln 81920x256x256 +bias+res+LayerNorm": (lambda: ops.gemm_tc(hid16, W1, out, b1, 1, 1, M, C, C, residual=res, ln=(lng, lnb, 1e-5)), 2.0 * M * C * C),
    "deslice_out 20x4096x256x256 batched P +bias+res+LayerNorm": (lambda: ops.gemm_tc(hid16, PT16, out, b1, B, 1, Hg * Wg, C, C, w_batched=1, residual=res, ln=(lng, lnb, 1e-5)), 2.0 * M * C * C),
    "conv_wgrad 2304x512x81920": (lambda: ops.gemm_tc_wgrad(x16, dXF16, B, Hg, Wg, C, I2, taps=9, scatter=(dWx, dWfx), I=C), 2.0 * M * 9 * C * I2),
}
only = os.environ.get("MB_ONLY")          # comma-separated name prefixes (for ncu captures)
warm = int(os.environ.get("MB_WARM", "3"))
if only:
    cases = {k: v for k, v in cases.items() if any(k.startswith(o) for o in only.split(","))}
for name, (fn, flops) in cases.items():
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:45s} median {med*1e3:8.1f} us   {flops/med/1e9:8.1f} TFLOP/s")

if only:
    sys.exit(0)
# slice stage: exact SIMT kernels vs the tcgen05 (tf32) kernels
from transformerbasednavierstokesolver_b200 import _lib
lib = _lib.load()
H, D, G = 8, 32, 32
N = Hg * Wg
XFs = torch.randn(M, I2, generator=g).to(dev)
XFs16 = XFs.bfloat16()
Ws = (torch.randn(G, D, generator=g) * 0.3).to(dev); bs = torch.randn(G, generator=g).to(dev); tau = torch.full((H,), 0.5, device=dev)
groups = lib.tbns_slice_groups(B, N, H)
w16 = torch.empty(B, N, H * G, device=dev, dtype=torch.bfloat16)
part = torch.empty(B * H * groups * G * (D + 1), device=dev)
dw = torch.randn(B, N, H * G, generator=g).to(dev); dTt = torch.randn(B, H, G, D, generator=g).to(dev); ds = torch.randn(B, H, G, generator=g).to(dev)
dXF16o = torch.empty(M, I2, device=dev, dtype=torch.bfloat16)
dw16 = dw.bfloat16()
dWs_p = torch.empty(B * H * groups, G * (D + 1), device=dev); dtau_p = torch.empty(B * H * groups, device=dev); dbc = torch.empty(B * groups, H * 2 * D, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
cases2 = {
    "slice_fwd SIMT": lambda: lib.tbns_pa_slice_fwd(P(XFs), P(Ws), P(bs), P(tau), None, P(w16), P(part), B, N, H, D, G, 1, st),
    "slice_fwd tcgen05": lambda: lib.tbns_pa_slice_fwd_tc(P(XFs16), P(Ws), P(bs), P(tau), P(w16), P(part), B, N, H, D, G, 1, st),
    "slice_bwd SIMT": lambda: lib.tbns_pa_slice_bwd(P(XFs), P(Ws), P(bs), P(tau), P(dw), P(dTt), P(ds), None, P(dXF16o), P(dWs_p), P(dtau_p), P(dbc), B, N, H, D, G, 1, st),
    "slice_bwd tcgen05": lambda: lib.tbns_pa_slice_bwd_tc(P(XFs16), P(Ws), P(bs), P(tau), P(dw16), P(dTt), P(ds), P(dXF16o), P(dWs_p), P(dtau_p), B, N, H, D, G, 1, st),
}
for name, fn in cases2.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:45s} median {ts[len(ts)//2]*1e3:8.1f} us")
