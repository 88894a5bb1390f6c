"""Bring-up probe for pool_tc.cu: structured inputs that show WHICH patch rows reach the accumulators with WHICH weights."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import imp_b200
from imp_b200 import kernels

torch.set_printoptions(linewidth=220, precision=3, sci_mode=False)
dev = "cuda"

def run(n, P=4, ramp=False):
    h = torch.zeros(n, 256)
    h[torch.arange(n), torch.arange(n) % 256] = 1.0
    q = torch.zeros(1, P, 256)
    if ramp:
        q[0, 0] = torch.arange(256) / 64.0          # S_n = q[n % 256]: weights identify the row
    cu = torch.tensor([0, n], dtype=torch.int32)
    pooled, lse = kernels.pool_fwd(h.bfloat16().to(dev), cu.to(dev), n, q.to(dev))
    torch.cuda.synchronize()
    s = (q[0, 0][torch.arange(n) % 256]).double()
    w = torch.exp(s - s.max())
    ref = torch.zeros(256, dtype=torch.double)
    ref.index_add_(0, torch.arange(n) % 256, w)
    ref = ref / w.sum()
    got = pooled[0, 0].double().cpu()
    err = (got - ref).abs().max().item()
    print("n=%d ramp=%s  max abs err %.3e  lse %.5f (ref %.5f)" % (n, ramp, err, lse[0, 0].item(), (s.max() + torch.log(w.sum())).item()))
    if err > 1e-3:
        ratio = got / ref.clamp_min(1e-30)
        for blk in range(0, 256, 32):
            print("   f %3d..: got*N " % blk, (got[blk:blk + 32] * n).numpy().round(2).tolist())
            if ramp:
                print("            ref*N ", (ref[blk:blk + 32] * n).numpy().round(2).tolist())

for n in (64, 128, 192, 256, 512, 1024):
    run(n)
for n in (128, 256, 512):
    run(n, ramp=True)
os.environ["X"] = "1"
