"""cfg-1 model (8 layers, C=256, 64x64, batched teacher-forced step of 20 images): full flat gradient of the bf16 tensor-core
path vs the exact fp32 SIMT path on identical weights and data; then a short SGD run in both modes from the same start."""
import sys
import torch
sys.path.insert(0, ".")
import transformerbasednavierstokesolver_b200 as pkg
from transformerbasednavierstokesolver_b200 import train
from transformerbasednavierstokesolver_b200.model.Transolver_Structured_Mesh_2D import Model

dev = torch.device("cuda:0")
CFG = dict(space_dim=2, n_layers=8, n_hidden=256, dropout=0.0, n_head=8, Time_Input=False, mlp_ratio=1, fun_dim=10, out_dim=1,
           slice_num=32, ref=8, unified_pos=1, H=64, W=64)
x, fx, yy = train.synthetic_ns_batch(2, 64, 10, 10, seed=1, device=dev)
mix = torch.randn(10, 10, generator=torch.Generator().manual_seed(2)).to(dev) / 10 ** 0.5
yy = fx @ mix
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for prec in ("fp32", "bf16"):
    pkg.set_default_precision(prec)
    torch.manual_seed(0)
    m = Model(**CFG).to(dev)
    g = train.FlatGradients(m.parameters())
    g.zero()
    loss = train.step_loss(m, x, fx, yy, 10, 1, True)
    loss.backward()
    res[prec] = (float(loss), g.flat.clone(), [(n, p.grad.clone()) for n, p in m.named_parameters()])
    opt = torch.optim.SGD(m.parameters(), lr=2e-3)
    curve = []
    for i in range(8 if prec == "fp32" else 30):
        curve.append(float(train.train_step(m, opt, None, g, x, fx, yy, 10, 1, batched=True)) / 20)
    res[prec + "_curve"] = curve
a, b = res["fp32"][1].double(), res["bf16"][1].double()
print(f"loss fp32 {res['fp32'][0]:.6f}  bf16 {res['bf16'][0]:.6f}")
print(f"flat gradient ({a.numel()} params): rel-L2 {float((a-b).norm()/a.norm()):.3e}  cosine {float((a@b)/(a.norm()*b.norm())):.6f}")
worst = sorted(((float((q.double()-p.double()).norm()/(p.double().norm()+1e-30)), n) for (n, p), (_, q) in zip(res["fp32"][2], res["bf16"][2])), reverse=True)[:5]
print("largest per-parameter rel-L2:", [(f"{e:.2e}", n) for e, n in worst])
print("SGD lr 2e-3, per-call loss:  step  fp32  bf16")
for i, v in enumerate(res["bf16_curve"]):
    f = res["fp32_curve"][i] if i < len(res["fp32_curve"]) else None
    if i < 8 or i % 5 == 0 or i == 29:
        print(f"  {i:3d}  " + (f"{f:.5f}" if f is not None else "   -   ") + f"  {v:.5f}")
