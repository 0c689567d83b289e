"""Training sanity on one B200: the cfg-1 model overfits one fixed synthetic batch; loss curves of the bf16 tensor-core path and
the exact fp32 SIMT path are printed side by side (same seed, same data).  usage: python profiles/train_sanity.py [steps]"""
import sys
import torch
sys.path.insert(0, ".")
import transformerbasednavierstokesolver_b200 as pkg
from transformerbasednavierstokesolver_b200 import train
from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
CFG = dict(space_dim=2, n_layers=8, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
           slice_num=32, ref=8, unified_pos=1, H=64, W=64)
curves = {}
for prec in ("bf16", "fp32"):
    pkg.set_default_precision(prec)
    torch.manual_seed(0)
    m = Model(**CFG).to(dev)
    g = train.FlatGradients(m.parameters())
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
    x, fx, yy = train.synthetic_ns_batch(2, 64, 10, 10, seed=1, device=dev)
    # smooth, learnable target: next frames are a fixed linear mix of the input frames
    mix = torch.randn(10, 10, generator=torch.Generator().manual_seed(2)).to(dev) / 10 ** 0.5
    yy = fx @ mix
    n = steps if prec == "bf16" else min(steps, 12)   # the SIMT fp32 engine is ~10x slower: a short prefix is enough to compare
    losses = []
    for i in range(n):
        losses.append(float(train.train_step(m, opt, None, g, x, fx, yy, 10, 1, batched=True)) / 20.0)
    curves[prec] = losses
print("step  bf16(tcgen05)   fp32(SIMT)")
for i in range(len(curves["bf16"])):
    b = curves["bf16"][i]
    f = curves["fp32"][i] if i < len(curves["fp32"]) else None
    if i < 12 or i % 10 == 0 or i == len(curves["bf16"]) - 1:
        print(f"{i:4d}  {b:12.5f}  " + (f"{f:12.5f}" if f is not None else ""))
