import sys, torch
sys.path.insert(0, "/root/repo")
from transformerbasednavierstokesolver_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B, N, H, G, D = 1, 256, 1, 32, 32
g = torch.Generator().manual_seed(1)
XF = torch.randn(B * N, 2 * H * D, generator=g).to(dev)
Ws = (torch.randn(G, D, generator=g) * 0.4).to(dev)
bs = torch.randn(G, generator=g).to(dev)
tau = torch.linspace(0.5, 0.5, H).to(dev)
groups = lib.tbns_slice_groups(B, N, H)
st = torch.cuda.current_stream().cuda_stream
w_ref = torch.empty(B, N, H * G, device=dev)
part_ref = torch.empty(B * H * groups * G * (D + 1), device=dev)
_lib.check(lib.tbns_pa_slice_fwd(XF.data_ptr(), Ws.data_ptr(), bs.data_ptr(), tau.data_ptr(), w_ref.data_ptr(), None, part_ref.data_ptr(), B, N, H, D, G, 1, st), "f")
w16 = torch.zeros(B, N, H * G, device=dev, dtype=torch.bfloat16)
part = torch.full((B * H * groups * G * (D + 1),), float("nan"), device=dev)
_lib.check(lib.tbns_pa_slice_fwd_tc(XF.data_ptr(), Ws.data_ptr(), bs.data_ptr(), tau.data_ptr(), w16.data_ptr(), part.data_ptr(), B, N, H, D, G, 1, st), "tc")
torch.cuda.synchronize()
p = part.view(groups, G, D + 1).sum(0).cpu(); pr = part_ref.view(groups, G, D + 1).sum(0).cpu()
torch.set_printoptions(precision=4, linewidth=200)
print("groups", groups)
print("tc  row0", p[0, :8], "row1", p[1, :8])
print("ref row0", pr[0, :8], "row1", pr[1, :8])
print("tc col0 over g", p[:8, 0]); print("ref col0 over g", pr[:8, 0])
# is tc some permutation/transposition of ref?
print("match transpose?", float((p[:, :32] - pr[:, :32].t()).abs().max()), "ref absmax", float(pr.abs().max()))
print("nonzero frac", float((p[:, :32] != 0).float().mean()))
