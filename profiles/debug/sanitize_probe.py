"""Small, ragged shapes through every round-2 kernel, meant to run under compute-sanitizer (memcheck / racecheck):
pool_tc forward + dq (P = 6, 32, 64; bags of 1, 37, 129, 300, 1000 rows; shared and per-bag queries), the dz pass with
the fused dq~, modularity prep on tcgen05 (tiles spanning two bags, worst-case sized buffers), the whole hot path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import imp_b200
from imp_b200 import kernels, model as M, step as S, modularity as MOD

dev = "cuda"
torch.manual_seed(0)
for lens, P, shared in (([1], 6, False), ([37, 129, 300], 32, True), ([1000, 64], 64, False), ([257, 640, 130], 16, False)):
    total = sum(lens)
    h = torch.relu(torch.randn(total, 256)).bfloat16().to(dev)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    qt = (torch.randn(1 if shared else len(lens), P, 256) * 0.08).to(dev)
    pooled, lse = kernels.pool_fwd(h, cu, max(lens), qt)
    dp = torch.randn(len(lens), P, 256, device=dev)
    delta = (dp.bfloat16().float() * pooled).sum(-1).contiguous()
    dq, _ = kernels.pool_bwd(h, cu, max(lens), [qt], [dp], [lse], [delta], 0, want_dz=False)
    if P <= 32:
        db1 = torch.empty(256, device=dev)
        dq2, dz = kernels.pool_bwd(h, cu, max(lens), [qt, qt], [dp, dp], [lse, lse], [delta, delta], 0, want_dz=True, db1=db1)
    c1 = torch.randn(len(lens), min(P, 32), 256, device=dev)
    c2 = torch.randn(len(lens), 7, 256, device=dev)
    # a one-patch bag has an empty graph (e = 0): the loss is 0/0 here exactly as in the reference (utils.py:201,222)
    t = MOD.modularity_terms(h, cu, max(lens), c1, c2) if min(lens) > 1 else torch.zeros(1, device=dev)
    torch.cuda.synchronize()
    print(lens, P, shared, 'pooled', bool(torch.isfinite(pooled).all()), 'lse', bool(torch.isfinite(lse).all()), 'dq', bool(torch.isfinite(dq).all()), 'terms', bool(torch.isfinite(t).all()), (bool(torch.isfinite(dq2).all()), bool(torch.isfinite(dz.float()).all()), bool(torch.isfinite(db1).all())) if P <= 32 else None, flush=True)
    assert torch.isfinite(pooled).all() and torch.isfinite(dq).all() and torch.isfinite(t).all()
# the whole hot path from the reference batch layout (worst-case sized packed buffer: rows past cu[B] belong to no bag)
net = M.IMPHotPath(n_proto=6, dropout=0.25, seed=0).to(dev)
runner = S.HotPathStep(net).to(dev).train()
img = torch.full((3, 512, 512), -10000.0)
for i, n in enumerate((300, 2, 417)):
    img[i, :n] = torch.randn(n, 512)
loss = runner({"img": img.to(dev), "omic": torch.rand(3, 3354, device=dev)}, torch.randn(3, 6, 256, device=dev), torch.randn(3, 7, 256, device=dev))
loss.backward()
torch.cuda.synchronize()
assert torch.isfinite(loss).item()
print("sanitize probe ok")
