"""Kernel-level profile of the sharded giant-bag step (configs[3]) on rank 0: where the time above sweep / world goes.
torchrun --nproc-per-node N profiles/debug/giant_profile.py > gpurun_out/giant_profile.txt"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import imp_b200  # noqa
from imp_b200 import model as M, modularity as MOD, ops, parallel as PAR

n, P = 120000, 32
torch.manual_seed(1234)
net = M.IMPHotPath(n_proto=P, dropout=0.25, seed=0).to(dev).train()
a, b = PAR.shard_bounds(n, world)[rank]
x = torch.randn(b - a, 512, device=dev).bfloat16()
cot = torch.randn(1, P, 256, device=dev) * 1e-2
grp = dist.group.WORLD if world > 1 else None
params = list(net.parameters())
blocks = [ops.block_params(blk) for blk in net.proto_g_blocks]
cu = torch.tensor([0, b - a], dtype=torch.int32, device=dev)


def step():
    for p in params:
        p.grad = None
    c, h = ops.proto_fusion(x, cu, b - a, net.p_proto, net.path_net[0].weight, net.path_net[0].bias, blocks,
                            p_drop=net.dropout, seed=net._seed(), shard_group=grp)
    loss = (c * cot).sum()
    if grp is not None:
        loss = loss + MOD.modularity_terms_sharded(h, a, n, c, group=grp)[0, 0]
    else:
        loss = loss + MOD.modularity_terms(h, cu, b - a, c)[0, 0]
    loss.backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    cnt, tot = collections.Counter(), collections.Counter()
    t_min, t_max = None, None
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            name = e.name[:100]
            cnt[name] += 1
            tot[name] += e.device_time
            s, f = e.time_range.start, e.time_range.end
            t_min = s if t_min is None else min(t_min, s)
            t_max = f if t_max is None else max(t_max, f)
    print("world %d: %.3f ms per step (events); profiled 3 steps: span %.3f ms, kernel time %.3f ms, %d device activities"
          % (world, ms, (t_max - t_min) / 1e3, sum(tot.values()) / 1e3, sum(cnt.values())))
    for name, t in tot.most_common(30):
        print("%9.1f us  %4d x  %s" % (t / 3, cnt[name] // 3 if cnt[name] >= 3 else cnt[name], name))
if world > 1:
    dist.destroy_process_group()
