"""2 GPUs, NCCL: (a) a bag sharded over ranks (LSE merge of the partial softmax states, all-reduced
dq~/dW1) reproduces the single-GPU tokens and gradients; (b) slide-parallel gradient all-reduce equals
the single-GPU gradients over all slides; (c) the modularity loss of one bag sharded by rows over the ranks
(all-gather of xh / assignments, all-reduce of the partial traces and token gradients) equals the single-GPU
loss and gradients; (d) the sharded step with its collectives captured in one CUDA graph equals the eager step.
Skipped with fewer than 2 GPUs."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import imp_b200  # noqa: F401
        from imp_b200 import ops, parallel as P, step as S
        from util_hotpath import block_tensors, make_bags, make_params, rel
        dev = torch.device("cuda", rank)
        params = make_params(0)
        g = torch.Generator().manual_seed(5)
        npatch, nproto = 5000, 16
        bag = make_bags([npatch], 1)[0].bfloat16()
        p_proto = ((torch.rand(1, nproto, 256, generator=g) * 2 - 1) / nproto).to(dev)
        cot = torch.randn(1, nproto, 256, generator=g).to(dev)

        def run(x, shard_group):
            leaves = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
            cu = torch.tensor([0, x.shape[0]], dtype=torch.int32, device=dev)
            blocks = [block_tensors(leaves, 0), block_tensors(leaves, 1)]
            c, _ = ops.proto_fusion(x.to(dev).contiguous(), cu, max(1, x.shape[0]), p_proto, leaves["path_net.0.weight"],
                                    leaves["path_net.0.bias"], blocks, shard_group=shard_group)
            (c * cot).sum().backward()
            return c.detach(), {k: v.grad for k, v in leaves.items()}

        a, b = P.shard_bounds(npatch, world)[rank]
        c_sh, g_sh = run(bag[a:b], dist.group.WORLD)
        c_full, g_full = run(bag, None)
        err_c = rel(c_sh, c_full)
        err_g = max(rel(g_sh[k], g_full[k]) for k in g_full)

        # slide parallel
        lens = [700, 300, 512, 900]
        bags = [t.bfloat16() for t in make_bags(lens, 2)]
        cot2 = torch.randn(len(lens), nproto, 256, generator=g).to(dev)

        def run_slides(idx):
            leaves = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(v.clone().to(dev)) for k, v in params.items()})
            L = {k: leaves[k.replace(".", "_")] for k in params}
            x = torch.cat([bags[i] for i in idx]).to(dev).contiguous()
            cu = ops._cu_from_lengths([lens[i] for i in idx], dev)
            c, _ = ops.proto_fusion(x, cu, max(lens), p_proto, L["path_net.0.weight"], L["path_net.0.bias"],
                                    [block_tensors(L, 0), block_tensors(L, 1)])
            (c * cot2[idx]).sum().backward()
            return leaves

        mine = P.assign_slides(lens, world)[rank]
        lv = run_slides(mine)
        S.allreduce_gradients(lv, world)
        full = run_slides(list(range(len(lens))))
        err_dp = max(rel(lv[k].grad * world, full[k].grad) for k in full.keys())

        # modularity of one bag sharded by rows (imp_modularity_prepare / _execute around the collectives)
        from imp_b200 import modularity as MOD
        nmod = 3000 + 37
        gm = torch.Generator().manual_seed(9)
        hm = torch.relu(torch.randn(nmod, 256, generator=gm) + 0.5 * torch.randn(1, 256, generator=gm)).bfloat16().to(dev)
        cp = torch.randn(1, 16, 256, generator=gm).to(dev)
        co = torch.randn(1, 7, 256, generator=gm).to(dev)
        cp_s, co_s = cp.clone().requires_grad_(True), co.clone().requires_grad_(True)
        a2, b2 = P.shard_bounds(nmod, world)[rank]
        t_sh = MOD.modularity_terms_sharded(hm[a2:b2].contiguous(), a2, nmod, cp_s, co_s, group=dist.group.WORLD)
        (t_sh[0, 0] + 2.0 * t_sh[0, 1]).backward()
        cp_f, co_f = cp.clone().requires_grad_(True), co.clone().requires_grad_(True)
        cu_m = torch.tensor([0, nmod], dtype=torch.int32, device=dev)
        t_full = MOD.modularity_terms(hm, cu_m, nmod, cp_f, co_f)
        (t_full[0, 0] + 2.0 * t_full[0, 1]).backward()
        err_m = max(((t_sh - t_full).abs() / (t_full.abs() + 1e-6)).max().item(), rel(cp_s.grad, cp_f.grad), rel(co_s.grad, co_f.grad))
        # (d) the sharded training step (pooling + modularity, collectives included) replayed from ONE CUDA graph equals
        #     the eager step
        leaves = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(v.clone().to(dev)) for k, v in params.items()})
        L = {k: leaves[k.replace(".", "_")] for k in params}
        xs = bag[a:b].to(dev).contiguous()
        cu_s = torch.tensor([0, b - a], dtype=torch.int32, device=dev)
        blocks_g = [block_tensors(L, 0), block_tensors(L, 1)]

        def loss_fn():
            c, h = ops.proto_fusion(xs, cu_s, max(1, b - a), p_proto, L["path_net.0.weight"], L["path_net.0.bias"], blocks_g,
                                    shard_group=dist.group.WORLD)
            return (c * cot).sum() + MOD.modularity_terms_sharded(h, a, npatch, c, group=dist.group.WORLD)[0, 0]

        plist = list(leaves.parameters())
        loss_e = loss_fn()
        loss_e.backward()
        ref = {k: v.grad.clone() for k, v in leaves.items() if v.grad is not None}
        loss_e = float(loss_e)
        gs = S.GraphedStep(None).capture_fn(loss_fn, plist, dev)
        for _ in range(2):
            loss_g = gs.replay()
        torch.cuda.synchronize()
        err_graph = max([abs(float(loss_g) - loss_e) / abs(loss_e)] + [rel(leaves[k].grad, ref[k]) for k in ref])
        gs.close()
        q.put((rank, err_c, err_g, err_dp, err_m, err_graph))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_bag_and_slide_parallel():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_c, err_g, err_dp, err_m, err_graph in res:
        assert err_c < 2e-4, (rank, err_c)          # same kernels, different split of the softmax
        assert err_g < 2e-3, (rank, err_g)
        assert err_dp < 2e-3, (rank, err_dp)
        assert err_m < 5e-4, (rank, err_m)          # same kernels, row blocks split over the ranks
        assert err_graph < 1e-3, (rank, err_graph)  # same launches from a graph; fp32 atomics land in another order
