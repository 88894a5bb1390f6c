"""A2/A3 softmax pooling kernels vs the CPU oracle (oracle/imp_oracle.py) on the same inputs.
Tolerance: operands are bf16 on the device (h, and the P-wide operands inside the MMA), fp32
statistics; north_star asks 1e-3 relative on outputs/gradients, measured here as relative
Frobenius error against the oracle evaluated on the same bf16-rounded h."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _make(lens, P, seed, shared_q):
    g = torch.Generator().manual_seed(seed)
    total = sum(lens)
    h = torch.relu(torch.randn(total, 256, generator=g)).bfloat16()
    nb = len(lens)
    qt = torch.randn(1 if shared_q else nb, P, 256, generator=g) * 0.08
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    return h, qt, cu


@pytest.mark.parametrize("lens,P,shared", [([64], 16, True), ([1], 6, False), ([1000, 37, 513], 6, False),
                                           ([4096], 16, True), ([2500, 3000], 32, False), ([300, 200], 64, True),
                                           ([16384 + 11], 32, True)])
def test_pool_fwd_vs_oracle(lens, P, shared):
    from imp_b200 import kernels
    from oracle import imp_oracle as O
    h, qt, cu = _make(lens, P, 1, shared)
    pooled, lse = kernels.pool_fwd(h.cuda(), cu.cuda(), max(lens), qt.cuda())
    torch.cuda.synchronize()
    for b, n in enumerate(lens):
        hb = h[cu[b]:cu[b + 1]].float()
        q = qt[0 if shared else b].bfloat16().float()        # the kernel rounds q~ to bf16
        ref_pool, ref_lse = O.lse_merge([O.pool_partial(hb, q)])
        assert _rel(pooled[b].cpu(), ref_pool) < 2e-3, (b, _rel(pooled[b].cpu(), ref_pool))
        assert (lse[b].cpu() - ref_lse).abs().max().item() < 2e-3


def _oracle_bwd(hb, qts, dps, relu_mask, keep_scale):
    hb = hb.double().requires_grad_(True)
    qs = [q.double().requires_grad_(True) for q in qts]
    tot = 0
    for q, dp in zip(qs, dps):
        a = torch.softmax(q @ hb.t(), dim=1)
        tot = tot + ((a @ hb) * dp.double()).sum()
    grads = torch.autograd.grad(tot, [hb] + qs)
    dh = grads[0]
    if relu_mask:
        dh = dh * (hb.detach() > 0) * keep_scale
    return dh, grads[1:]


@pytest.mark.parametrize("lens,P,nblk,want_dz", [([64], 16, 1, False), ([700, 129], 6, 1, True),
                                                 ([700, 129], 6, 2, True), ([2048], 32, 2, True),
                                                 ([1500, 1000], 32, 1, False), ([333], 64, 1, True),
                                                 ([8192 + 7], 32, 2, True)])
def test_pool_bwd_vs_oracle(lens, P, nblk, want_dz):
    from imp_b200 import kernels
    from oracle import imp_oracle as O
    h, _, cu = _make(lens, P, 2, False)
    nb = len(lens)
    g = torch.Generator().manual_seed(7)
    qts = [(torch.randn(nb, P, 256, generator=g) * 0.08).bfloat16().float() for _ in range(nblk)]
    dps = [(torch.randn(nb, P, 256, generator=g)).bfloat16().float() for _ in range(nblk)]
    hd, cud = h.cuda(), cu.cuda()
    lses, deltas = [], []
    for k in range(nblk):
        pooled, lse = kernels.pool_fwd(hd, cud, max(lens), qts[k].cuda())
        lses.append(lse)
        deltas.append((dps[k].cuda() * pooled).sum(-1).contiguous())
    db1 = torch.zeros(256, device="cuda")
    dq_block = nblk - 1 if not want_dz else 0
    dq, dz = kernels.pool_bwd(hd, cud, max(lens), [q.cuda() for q in qts], [d.cuda() for d in dps], lses, deltas,
                              dq_block, want_dz, relu_mask=True, keep_scale=1.25, db1=db1)
    torch.cuda.synchronize()
    ref_dz = []
    for b in range(nb):
        hb = h[cu[b]:cu[b + 1]].float()
        dh, dqs = _oracle_bwd(hb, [q[b] for q in qts], [d[b] for d in dps], True, 1.25)
        ref_dz.append(dh)
        assert _rel(dq[b].cpu(), dqs[dq_block]) < 5e-3, (b, _rel(dq[b].cpu(), dqs[dq_block]))
    if want_dz:
        ref = torch.cat(ref_dz)
        assert _rel(dz.cpu().float(), ref) < 6e-3, _rel(dz.cpu().float(), ref)      # dz is stored in bf16
        assert _rel(db1.cpu(), ref.sum(0)) < 5e-3
