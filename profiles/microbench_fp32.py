"""fp32 mode: 3xTF32 tcgen05 engine (TBNS_PREC_FP32) against the fp32 FMA engine (TBNS_PREC_FP32_EXACT) at the cfg-1 bench shapes.
    python profiles/microbench_fp32.py      (CUDA-event medians, L2 flushed between launches)"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from transformerbasednavierstokesolver_b200 import ops  # noqa: E402
from transformerbasednavierstokesolver_b200._lib import TBNS_PREC_FP32, TBNS_PREC_FP32_EXACT  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=7):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


Bimg, Hg, Wg, C, I2 = 20, 64, 64, 256, 512
M = Bimg * Hg * Wg
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, C, device=dev, generator=g)
Wf = torch.randn(I2, 9 * C, device=dev, generator=g) * 0.02
XF = torch.empty(M, I2, device=dev)
dXF = torch.randn(M, I2, device=dev, generator=g)
Wd = torch.randn(C, 9 * I2, device=dev, generator=g) * 0.02
dx = torch.empty(M, C, device=dev)
dW = torch.empty(9 * C, I2, device=dev)
W1 = torch.randn(C, C, device=dev, generator=g) * 0.05
y = torch.empty(M, C, device=dev)
cases = {
    "conv_fprop 81920x2304x512": (2.0 * M * 9 * C * I2, lambda p: ops.gemm(M=M, N=I2, K=9 * C, A=x, lda=C, a_kind=0, B=Wf, ldb=9 * C, b_kind=0, C=XF, ldc=I2,
                                                                       conv_mode=1, Hg=Hg, Wg=Wg, Cin=C, precision=p)),
    "conv_dgrad 81920x4608x256": (2.0 * M * 9 * C * I2, lambda p: ops.gemm(M=M, N=C, K=9 * I2, A=dXF, lda=I2, a_kind=0, B=Wd, ldb=9 * I2, b_kind=0, C=dx, ldc=C,
                                                                       conv_mode=1, Hg=Hg, Wg=Wg, Cin=I2, flip=1, precision=p)),
    "conv_wgrad 2304x512x81920": (2.0 * M * 9 * C * I2, lambda p: ops.gemm(M=9 * C, N=I2, K=M, A=x, lda=C, a_kind=1, B=dXF, ldb=I2, b_kind=1, C=dW, ldc=I2,
                                                                       conv_mode=2, Hg=Hg, Wg=Wg, Cin=C, split_k=ops._split_k(9 * C, I2, M), precision=p)),
    "linear 81920x256x256": (2.0 * M * C * C, lambda p: ops.gemm(M=M, N=C, K=C, A=x, lda=C, a_kind=0, B=W1, ldb=C, b_kind=0, C=y, ldc=C, precision=p)),
    "linear_wgrad 256x256x81920": (2.0 * M * C * C, lambda p: ops.gemm(M=C, N=C, K=M, A=x, lda=C, a_kind=1, B=dx, ldb=C, b_kind=1, C=W1.clone(), ldc=C,
                                                                     split_k=ops._split_k(C, C, M), precision=p)),
}
import os
only = os.environ.get("MB_ONLY")
for name, (flops, fn) in cases.items():
    if only and not name.startswith(only):
        continue
    fn(TBNS_PREC_FP32); fn(TBNS_PREC_FP32_EXACT); torch.cuda.synchronize()
    t3 = timeit(lambda: fn(TBNS_PREC_FP32))
    te = timeit(lambda: fn(TBNS_PREC_FP32_EXACT), n=3)
    print(f"{name:32s} 3xTF32 tcgen05 {t3:9.1f} us ({flops / t3 / 1e6:7.1f} TFLOP/s fp32-equivalent)   fp32 FMA {te:9.1f} us ({flops / te / 1e6:6.1f} TFLOP/s)   x{te / t3:.1f}",
          flush=True)
