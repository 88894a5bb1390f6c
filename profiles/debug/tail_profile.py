"""Kernel-level profile of one eager drop-in model step (torch.profiler, CUDA activity): which launches make up the
token tail.  Usage (GPU box): python profiles/debug/tail_profile.py [bags] [patches] > gpurun_out/tail_profile.txt"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from types import SimpleNamespace as NS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048          # small bags: the tail does not depend on the bag size
P = 32
dev = torch.device("cuda:0")
import imp_b200  # noqa
from imp_b200 import survival
from imp_b200.registry import build_model
import imp_b200.umeml_gan  # noqa
cfg = NS(DATASET=NS(ROOT=".", PATH=NS(DIM=512), OMIC=NS(DIM=sum(bench.GROUP_SIZES))),
         MODEL=NS(DROPOUT=0.25, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small",
                  UMEML=NS(PROTOTYPES=P, REGISTERS=3, GENE_GROUP_INDEXES=None, IMPORTANCE_LOG="defer")),
         TRAINER=NS(PREC="fp32"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
os.chdir(os.path.join(ROOT, "gpurun_out"))
model = build_model("umeml_gan", verbose=False, cfg=cfg, num_classes=4, omic_sizes=1000).to(dev).train()
model.plot_set = "prof"
x = torch.randn(B * N, 512, device=dev).bfloat16()
cu = torch.arange(0, B + 1, device=dev, dtype=torch.int32) * N
omic = torch.randn(B, sum(bench.GROUP_SIZES), device=dev)
batch = {"x_packed": x, "cu_seqlens": cu, "max_len": N, "omic": omic, "patient_id": [str(i) for i in range(B)]}
y = torch.randint(0, 4, (B,), device=dev)
c = torch.randint(0, 2, (B,), device=dev)


def step():
    for p in model.parameters():
        p.grad = None
    out = model(batch)
    loss = survival.nll_loss_new(out, y, c) + out[5] + out[1]
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
cnt, tot = collections.Counter(), collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name[:90]
        cnt[name] += 1
        tot[name] += e.device_time if hasattr(e, "device_time") else e.cuda_time
print("kernels %d, device time %.3f ms" % (sum(cnt.values()), sum(tot.values()) / 1e3))
for name, t in tot.most_common(45):
    print("%8.1f us  %5d x  %s" % (t, cnt[name], name))
