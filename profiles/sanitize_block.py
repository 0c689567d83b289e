"""Small-shape Transolver block forward+backward in bf16 mode (tensor-core kernels: CTA-pair GEMM / wgrad, persistent GEMM,
tf32 slice stage, token stage, LayerNorm) - the program compute-sanitizer is pointed at:
   compute-sanitizer --tool {memcheck,racecheck,synccheck} python profiles/sanitize_block.py"""
import sys
import torch
sys.path.insert(0, ".")
import transformerbasednavierstokesolver_b200 as tbns
from transformerbasednavierstokesolver_b200 import ops
from transformerbasednavierstokesolver_b200.model._blocks import Transolver_block

ops._USE_SIDE = False            # one stream: the tools serialise kernels anyway
tbns.set_default_precision("bf16")
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = Transolver_block(num_heads=8, hidden_dim=256, dropout=0.0, act="gelu", mlp_ratio=1, slice_num=32, H=16, W=24, last_layer=True).to(dev)
x = torch.randn(2, 16 * 24, 256, device=dev, requires_grad=True)
y = blk(x)
y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(x.grad.abs().mean()))
