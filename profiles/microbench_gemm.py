"""Micro-benchmark of the tcgen05 GEMM kernels at the bench shapes (B=20 images of 64x64, C=256): CUDA-event timing of
isolated launches with an L2 flush between iterations.  usage: python profiles/microbench_gemm.py [iters]"""
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
W1 = (torch.randn(C, C, generator=g) / 16).to(dev).bfloat16()
b1 = torch.randn(C, generator=g).to(dev)
pre = torch.empty(M, C, device=dev)
hid16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
res = torch.randn(M, C, generator=g).to(dev)
out = torch.empty(M, C, device=dev)
dXF16 = torch.randn(M, I2, generator=g).to(dev).bfloat16()
dWx = torch.empty(C, C, 3, 3, device=dev)
dWfx = torch.empty(C, C, 3, 3, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

cases = {
    "conv_fprop 81920x2304x512": (lambda: ops.gemm_tc(x16, Wf16, XF, bias, B, Hg, Wg, C, I2, 9, 0), 2.0 * M * 9 * C * I2),
    "fc1 81920x256x256 +bias+gelu+pre+bf16": (lambda: ops.gemm_tc(x16, W1, None, b1, 1, 1, M, C, C, act=1, aux_out=pre, C16=hid16), 2.0 * M * C * C),
    "fc2 81920x256x256 +bias+res": (lambda: ops.gemm_tc(hid16, W1, out, b1, 1, 1, M, C, C, residual=res), 2.0 * M * C * C),
    "conv_wgrad 2304x512x81920": (lambda: ops.gemm_tc_wgrad(x16, dXF16, B, Hg, Wg, C, I2, taps=9, scatter=(dWx, dWfx), I=C), 2.0 * M * 9 * C * I2),
}
for name, (fn, flops) in cases.items():
    for _ in range(3):
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
