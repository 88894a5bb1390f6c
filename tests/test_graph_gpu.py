"""CUDA-graph replay of the hot-path step (step.GraphedStep): the replayed loss and gradients equal the
eagerly launched step on the same inputs, refilling the static inputs changes the result accordingly, and
with dropout on every replay draws a new keep-mask through the device seed word (imp_set_seed_offset)."""
import pytest
import torch

from util_hotpath import rel

pytestmark = pytest.mark.gpu


def _setup(dropout):
    from imp_b200 import model as M, step as S
    from oracle import imp_oracle as O
    dev = torch.device("cuda")
    torch.manual_seed(0)
    lens, P = [700, 400, 555], 16
    net = M.IMPHotPath(n_proto=P, dropout=dropout, seed=0).to(dev)
    runner = S.HotPathStep(net).to(dev).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(sum(lens), 512, generator=g).bfloat16().to(dev)
    cu = torch.tensor([0, 700, 1100, 1655], dtype=torch.int32, device=dev)
    batch = {"x_packed": x, "cu_seqlens": cu, "max_len": max(lens), "omic": torch.rand(len(lens), sum(O.GROUP_SIZES), generator=g).to(dev)}
    cot_p = (torch.randn(len(lens), P, 256, generator=g) * 1e-2).to(dev)
    cot_o = (torch.randn(len(lens), 7, 256, generator=g) * 1e-2).to(dev)
    return S, runner, batch, cot_p, cot_o


def _eager(runner, params, batch, cot_p, cot_o):
    for p in params:
        p.grad = None
    loss = runner(batch, cot_p, cot_o)
    loss.backward()
    out = (loss.item(), [p.grad.clone() for p in params])
    del loss                      # a live autograd graph pins AccumulateGrad nodes to this stream and breaks capture
    return out


def test_graph_replay_matches_eager_step():
    S, runner, batch, cot_p, cot_o = _setup(0.0)
    params = list(runner.parameters())
    ref_loss, ref = _eager(runner, params, batch, cot_p, cot_o)
    cot_p.mul_(2.0)
    _, ref2 = _eager(runner, params, batch, cot_p, cot_o)
    cot_p.mul_(0.5)
    gs = S.GraphedStep(runner).capture(batch, cot_p, cot_o)
    try:
        for _ in range(2):
            out = gs.replay()
        torch.cuda.synchronize()
        assert abs(out.item() - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-6
        for p, r in zip(params, ref):
            assert rel(p.grad, r) < 1e-3      # float atomics in the modularity sums reorder between launches
        cot_p.mul_(2.0)           # refill a static input in place: the replay must see it
        gs.replay()
        torch.cuda.synchronize()
        for p, r in zip(params, ref2):
            assert rel(p.grad, r) < 1e-3      # float atomics in the modularity sums reorder between launches
    finally:
        gs.close()


def test_graph_replay_draws_a_new_dropout_mask():
    S, runner, batch, cot_p, cot_o = _setup(0.25)
    runner.with_modularity = False
    gs = S.GraphedStep(runner).capture(batch, cot_p, cot_o)
    try:
        gs.replay()
        g1 = runner.model.path_net[0].weight.grad.clone()
        gs.replay()
        g2 = runner.model.path_net[0].weight.grad.clone()
        torch.cuda.synchronize()
        assert torch.isfinite(g1).all() and torch.isfinite(g2).all()
        assert rel(g1, g2) > 1e-3          # different masks -> different gradients
    finally:
        gs.close()
