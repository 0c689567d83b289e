"""Auto-encoder variant at the bench shape (B=20 images of 64x64, C=256, 8 heads, slice_num 32), bf16 mode: CUDA-event timings
of the last encoder block's encode / decode (forward, and forward+backward).  usage: python profiles/microbench_autoencoder.py"""
import sys
import torch
sys.path.insert(0, ".")
import transformerbasednavierstokesolver_b200 as pkg
from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh2D_Encoder import Transolver_Encoder_block

pkg.set_default_precision("bf16")
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = Transolver_Encoder_block(num_heads=8, hidden_dim=256, dropout=0.0, mlp_ratio=1, last_layer=True, out_dim=1, slice_num=32, H=64, W=64).to(dev)
h = torch.randn(20, 4096, 256, device=dev, requires_grad=True)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def enc():
    with torch.no_grad():
        return blk.encode(h)


def dec():
    with torch.no_grad():
        return blk.decode(code)


def train_step():
    out = blk.decode(blk.encode(h))
    out.sum().backward()


code = enc()
print(f"encode  fwd (LN1 + projections + slice + token attention)      {timed(enc):7.3f} ms")
print(f"decode  fwd (project_slice + 2 x deslice/to_out + MLP + head)  {timed(dec):7.3f} ms")
print(f"decode(encode) fwd + bwd                                        {timed(train_step):7.3f} ms")
