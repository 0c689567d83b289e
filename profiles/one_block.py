"""One Transolver block forward+backward at the bench shape (B images of 64x64, C=256, 8 heads, 32 slices), bf16 mode.
Driver for ncu captures of individual kernels:  ncu -k regex:<kernel> ... python profiles/one_block.py [iters] [B]
(B = 20: the batched teacher-forced step; B = 2: one literal / unrolled call)"""
import sys
import torch
sys.path.insert(0, ".")
import transformerbasednavierstokesolver_b200 as tbns
from transformerbasednavierstokesolver_b200.model._blocks import Transolver_block

tbns.set_default_precision("bf16")
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
blk = Transolver_block(num_heads=8, hidden_dim=256, dropout=0.0, act="gelu", mlp_ratio=1, slice_num=32, H=64, W=64).to(dev)
x = torch.randn(B, 4096, 256, device=dev, requires_grad=True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    y = blk(x)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
