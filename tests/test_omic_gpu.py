"""A7/A8 omic encoders and missing-omics masks vs the CPU oracle; fp32 arithmetic -> 1e-5 relative."""
import pytest
import torch

from util_hotpath import rel

pytestmark = pytest.mark.gpu


def _setup(bsz, seed=0, sizes=None):
    from oracle import imp_oracle as O
    sizes = sizes or O.GROUP_SIZES
    g = torch.Generator().manual_seed(seed)
    G = sum(sizes)
    perm = torch.randperm(G, generator=g).tolist()
    groups, o = [], 0
    for s in sizes:
        groups.append(perm[o:o + s]); o += s
    x = torch.rand(bsz, G, generator=g)
    ws = [torch.randn(256, s, generator=g) / s ** 0.5 for s in sizes]
    bs = [0.1 * torch.randn(256, generator=g) for _ in sizes]
    return x, groups, ws, bs, g


@pytest.mark.parametrize("bsz", [1, 2, 13, 64])
def test_omic_encode_fwd_bwd(bsz):
    from imp_b200 import omics
    from oracle import imp_oracle as O
    x, groups, ws, bs, g = _setup(bsz)
    mask = (torch.rand(x.shape, generator=g) < 0.3).int()
    means = torch.rand(x.shape[1], generator=g)
    enc = omics.OmicEncoders(groups, dropout=0.0).cuda()
    with torch.no_grad():
        for k, m in enumerate(enc.omic_net):
            m[0].weight.copy_(ws[k]); m[0].bias.copy_(bs[k])
    out = enc(x.cuda(), mask.cuda(), means.cuda())
    seed = torch.randn(out.shape, generator=g)
    (out * seed.cuda()).sum().backward()
    torch.cuda.synchronize()
    wl = [w.clone().requires_grad_(True) for w in ws]
    bl = [b.clone().requires_grad_(True) for b in bs]
    ref = O.omic_encode(O.impute_missing_genes(x, mask, means), groups, wl, bl)
    (ref * seed).sum().backward()
    assert rel(out, ref) < 1e-5
    for k, m in enumerate(enc.omic_net):
        assert rel(m[0].weight.grad, wl[k].grad) < 1e-5, k
        assert rel(m[0].bias.grad, bl[k].grad) < 1e-5, k
    # without a mask
    out2 = enc(x.cuda())
    assert rel(out2, O.omic_encode(x, groups, ws, bs)) < 1e-5


def test_omic_dropout_statistics():
    from imp_b200 import omics
    x, groups, ws, bs, g = _setup(64)
    enc = omics.OmicEncoders(groups, dropout=0.25).cuda().train()
    out = enc(x.cuda())
    enc.eval()
    ref = enc(x.cuda())
    pos = ref > 1e-4
    kept = ((out != 0) & pos).sum().item() / pos.sum().item()
    assert abs(kept - 0.75) < 0.02, kept


@pytest.mark.parametrize("use_wo,use_mask", [(True, False), (False, True), (True, True), (False, False)])
def test_blend_missing_omics(use_wo, use_mask):
    from imp_b200 import omics
    from oracle import imp_oracle as O
    g = torch.Generator().manual_seed(3)
    bsz = 9
    h = torch.randn(bsz, 7, 256, generator=g)
    gen = torch.randn(bsz, 7, 256, generator=g)
    wo = (torch.rand(bsz, generator=g) < 0.5).int() if use_wo else None
    mask = (torch.rand(bsz, 3354, generator=g) < 0.5).int() if use_mask else None
    hd, gd = h.cuda().requires_grad_(True), gen.cuda().requires_grad_(True)
    out = omics.blend_missing_omics(hd, gd, wo.cuda() if use_wo else None, mask.cuda() if use_mask else None)
    hr, gr = h.clone().requires_grad_(True), gen.clone().requires_grad_(True)
    ref = O.blend_missing_omics(hr, gr, wo, mask)
    assert rel(out, ref) < 1e-6
    seed = torch.randn(ref.shape, generator=g)
    if out.requires_grad:
        (out * seed.cuda()).sum().backward()
        (ref * seed).sum().backward()
        assert rel(hd.grad, hr.grad) < 1e-6
        if gr.grad is not None and gr.grad.abs().sum() > 0:
            assert rel(gd.grad, gr.grad) < 1e-6
    # all-zero masks are no-ops, like the reference's "if sum > 0" guards
    z = omics.blend_missing_omics(h.cuda(), gen.cuda(), torch.zeros(bsz, dtype=torch.int32).cuda(),
                                  torch.zeros(bsz, 3354, dtype=torch.int32).cuda())
    assert torch.equal(z.cpu(), h)
