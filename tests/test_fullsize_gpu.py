"""Properties at BASELINE.json's full sizes, where the CPU oracle is too slow to be the checker:
  * 16384x512 bag, P=32: softmax pooling is invariant to the order of the patches, and pooling the whole
    bag equals the log-sum-exp merge of two halves (imp_lse_merge);
  * 16384 patches: the modularity loss is invariant to a permutation of the patches;
  * dW1 is linear in dz (imp_pathnet_dw of a sum equals the sum);
  * 2^20 x 512 k-means (config 5): every row's chosen centroid is optimal within 1e-3 in fp64 on a
    random sample, assignments are permutation-equivariant, the encode of the centroids themselves is the
    identity map, and sums/counts of the Lloyd update add up to the data."""
import pytest
import torch

from util_hotpath import block_tensors, make_params, rel

pytestmark = pytest.mark.gpu


def test_pooling_permutation_and_split_invariance_16k():
    from imp_b200 import kernels
    n, p = 16384, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    h = torch.relu(torch.randn(n, 256, device="cuda", generator=g)).bfloat16()
    qt = (torch.randn(1, p, 256, device="cuda", generator=g) * 0.08)
    cu = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    pooled, lse = kernels.pool_fwd(h, cu, n, qt)
    perm = torch.randperm(n, device="cuda", generator=g)
    pooled_p, lse_p = kernels.pool_fwd(h[perm].contiguous(), cu, n, qt)
    assert rel(pooled_p, pooled) < 2e-4 and (lse_p - lse).abs().max().item() < 1e-4
    half = n // 2
    cu2 = torch.tensor([0, half, n], dtype=torch.int32, device="cuda")
    pp, ll = kernels.pool_fwd(h, cu2, half, qt)                                   # two "bags" = two shards
    merged, mlse = kernels.lse_merge(pp.unsqueeze(0).contiguous(), ll.unsqueeze(0).contiguous())
    assert rel(merged, pooled) < 2e-4 and (mlse - lse).abs().max().item() < 1e-4


def test_modularity_permutation_invariance_16k():
    from imp_b200 import modularity as M
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.randn(8, 256, device="cuda", generator=g)
    h = torch.relu((torch.rand(n, 8, device="cuda", generator=g) ** 3) @ base + 0.3 * torch.randn(n, 256, device="cuda", generator=g)).bfloat16()
    c = torch.randn(1, 32, 256, device="cuda", generator=g).requires_grad_(True)
    co = torch.randn(1, 7, 256, device="cuda", generator=g)
    cu = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    a = M.modularity_terms(h, cu, n, c, co)
    (ga,) = torch.autograd.grad(a[0, 0], c)
    perm = torch.randperm(n, device="cuda", generator=g)
    b = M.modularity_terms(h[perm].contiguous(), cu, n, c, co)
    (gb,) = torch.autograd.grad(b[0, 0], c)
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 2e-4 * a.abs().max().item()
    assert rel(gb, ga) < 1e-3


def test_pathnet_dw_linearity_16k():
    from imp_b200 import kernels
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n, 512, device="cuda", generator=g).bfloat16()
    d1 = torch.randn(n, 256, device="cuda", generator=g).bfloat16()
    d2 = torch.randn(n, 256, device="cuda", generator=g).bfloat16()
    w1, w2 = kernels.pathnet_dw(d1, x), kernels.pathnet_dw(d2, x)
    w12 = kernels.pathnet_dw((d1.float() + d2.float()).bfloat16(), x)
    assert rel(w12, w1 + w2) < 3e-3          # the sum is re-rounded to bf16


def test_kmeans_one_million_rows():
    from imp_b200 import kernels
    n, k, d = 1 << 20, 32, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, d, device="cuda", generator=g)
    pick = torch.randperm(n, device="cuda", generator=g)[:k]
    mu = x[pick].clone()
    assign, dist = kernels.kmeans_assign(x, mu, want_dist=True)
    assert assign[pick].cpu().tolist() == list(range(k))               # a centroid row is its own nearest centroid
    assert int(assign.min()) >= 0 and int(assign.max()) < k
    # optimality on a random sample, in fp64
    idx = torch.randperm(n, device="cuda", generator=g)[:50000]
    d64 = torch.cdist(x[idx].double(), mu.double()) ** 2
    best = d64.min(dim=1).values
    chosen = d64.gather(1, assign[idx].long()[:, None])[:, 0]
    assert (chosen - best).max().item() <= 1e-3
    srt = d64.sort(dim=1).values
    clear = (srt[:, 1] - srt[:, 0]) > 1e-3
    assert torch.equal(assign[idx][clear].long(), d64.argmin(dim=1)[clear])
    assert (dist[idx] - best.float()).abs().max().item() < 2e-2       # fp32 distance of magnitude ~1e3
    # permutation equivariance
    perm = torch.randperm(n, device="cuda", generator=g)
    assert torch.equal(kernels.kmeans_assign(x[perm].contiguous(), mu), assign[perm])
    # checksum of the Lloyd update: counts add to n, sums add to the column sums of x
    sums, counts = kernels.kmeans_update(x, assign, k)
    assert int(counts.sum()) == n
    assert rel(sums.sum(0), x.double().sum(0)) < 1e-4
