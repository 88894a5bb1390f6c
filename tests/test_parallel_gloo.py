"""CPU, world_size 2 over gloo: the host-side multi-GPU logic.
  * slides data-parallel: per-rank gradients + allreduce_gradients == single-process gradients over all slides;
  * giant bag: per-shard partial pooling states gathered in rank order merge to the full softmax pooling.
The arithmetic on each rank is the CPU oracle (the kernels need a GPU); what is under test is the
partitioning, the collectives and their ordering."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import imp_b200  # noqa: F401
        from imp_b200 import parallel as P, step as S
        from oracle import imp_oracle as O
        from util_hotpath import make_bags, make_params
        torch.set_num_threads(2)
        params = make_params(0)
        lens = [96, 160, 64, 200]
        bags = make_bags(lens, 0)
        g = torch.Generator().manual_seed(5)
        p_proto = (torch.rand(6, 256, generator=g) * 2 - 1) / 6
        cot = torch.randn(len(lens), 6, 256, generator=g)

        # ---- data parallel over slides ----
        mine = P.assign_slides(lens, world)[rank]
        ref_local = O.hot_path_step([bags[i] for i in mine], params, p_proto, with_modularity=False, grad_seed=cot[mine])
        holder = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(v.clone()) for k, v in params.items()})
        for k, v in params.items():
            holder[k.replace(".", "_")].grad = ref_local["grads"][k].clone()
        S.allreduce_gradients(holder, world)
        full = O.hot_path_step(bags, params, p_proto, with_modularity=False, grad_seed=cot)
        worst = 0.0
        for k in params:
            got = holder[k.replace(".", "_")].grad * world           # mean over ranks -> sum over slides
            worst = max(worst, ((got - full["grads"][k]).norm() / full["grads"][k].norm().clamp_min(1e-30)).item())

        # ---- giant bag sharded over ranks ----
        h = torch.relu(torch.randn(1000, 256, generator=torch.Generator().manual_seed(9)))
        qt = torch.randn(6, 256, generator=torch.Generator().manual_seed(10)) * 0.1
        a, b = P.shard_bounds(1000, world)[rank]
        m, l, acc = O.pool_partial(h[a:b], qt)
        pooled_r, lse_r = (acc / l[:, None]).unsqueeze(0), (m + torch.log(l)).unsqueeze(0)
        part_p, part_l = P.gather_partials(pooled_r, lse_r)
        parts = [(part_l[0, r], torch.ones(6), part_p[0, r]) for r in range(world)]
        merged, lse = O.lse_merge(parts)
        ref_pool, ref_lse = O.lse_merge([O.pool_partial(h, qt)])
        err_pool = ((merged - ref_pool).norm() / ref_pool.norm()).item()
        err_lse = (lse - ref_lse).abs().max().item()
        q.put((rank, worst, err_pool, err_lse, tuple(part_p.shape)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_data_parallel_and_sharded_bag():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst, err_pool, err_lse, shape in res:
        assert worst < 1e-5, (rank, worst)
        assert err_pool < 1e-5 and err_lse < 1e-5, (rank, err_pool, err_lse)
        assert shape == (1, world, 6, 256)
