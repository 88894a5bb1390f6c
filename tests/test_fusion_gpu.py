"""Fused hot path (A0-A3: strip, path_net, two prototype blocks) forward + backward vs the CPU
oracle on the same bf16-rounded features and W1.  north_star tolerance: 1e-3 relative for
outputs and gradients; measured as relative Frobenius error per tensor.  h is stored in bf16 on
the device (the oracle keeps fp32), which is the dominant difference."""
import pytest
import torch

from util_hotpath import block_tensors, make_bags, make_params, rel

pytestmark = pytest.mark.gpu


def _run_device(bags, params, p_proto, grad_seed, p_drop=0.0):
    from imp_b200 import ops
    dev = "cuda"
    leaves = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
    lens = [b.shape[0] for b in bags]
    x = torch.cat(bags).to(dev).bfloat16().contiguous()
    cu = ops._cu_from_lengths(lens, dev)
    blocks = [block_tensors(leaves, 0), block_tensors(leaves, 1)]
    c, h = ops.proto_fusion(x, cu, max(lens), p_proto.to(dev), leaves["path_net.0.weight"],
                            leaves["path_net.0.bias"], blocks, p_drop=p_drop, seed=11)
    (c * grad_seed.to(dev)).sum().backward()
    torch.cuda.synchronize()
    return c.detach().cpu(), {k: v.grad.cpu() for k, v in leaves.items()}, h


@pytest.mark.parametrize("lens,P", [([4096], 16), ([700, 1300, 64], 6), ([2048, 2048], 32)])
def test_fusion_fwd_bwd_vs_oracle(lens, P):
    from oracle import imp_oracle as O
    params = make_params(0)
    bags = [b.bfloat16().float() for b in make_bags(lens, 0)]
    params["path_net.0.weight"] = params["path_net.0.weight"].bfloat16().float()
    g = torch.Generator().manual_seed(5)
    p_proto = (torch.rand(1, P, 256, generator=g) * 2 - 1) / P
    grad_seed = torch.randn(len(lens), P, 256, generator=g)
    ref = O.hot_path_step(bags, params, p_proto[0], with_modularity=False, grad_seed=grad_seed)
    c, grads, _ = _run_device(bags, params, p_proto, grad_seed)
    assert rel(c, ref["c"]) < 1e-3, rel(c, ref["c"])
    worst = {k: rel(grads[k], ref["grads"][k]) for k in grads}
    bad = {k: v for k, v in worst.items() if v > 5e-3}
    assert not bad, worst


def test_fusion_dropout_mask_is_regenerated_in_backward():
    """With dropout on, dW1 must vanish exactly where the forward dropped: rows of dz are masked by h>0."""
    params = make_params(1)
    bags = make_bags([1000], 3)
    g = torch.Generator().manual_seed(5)
    p_proto = (torch.rand(1, 16, 256, generator=g) * 2 - 1) / 16
    grad_seed = torch.randn(1, 16, 256, generator=g)
    c, grads, h = _run_device(bags, params, p_proto, grad_seed, p_drop=0.25)
    frac_zero = (h == 0).float().mean().item()
    assert 0.55 < frac_zero < 0.70, frac_zero        # relu (~50%) and dropout (25% of the rest)
    assert torch.isfinite(c).all() and all(torch.isfinite(v).all() for v in grads.values())


def test_strip_and_pack_exact():
    from imp_b200 import kernels, ops
    from oracle import imp_oracle as O
    lens = [37, 10000, 1, 512, 9999]
    npad = 10000
    g = torch.Generator().manual_seed(0)
    img = torch.full((len(lens), npad, 512), O.SENTINEL)
    for i, n in enumerate(lens):
        img[i, :n] = torch.randn(n, 512, generator=g)
    img[3, 300, 17] = O.SENTINEL                      # a sentinel inside a row ends the bag there (reference rule)
    want = [O.bag_length(img[i]) for i in range(len(lens))]
    assert want == [37, 10000, 1, 300, 9999]
    lengths, cu = kernels.bag_lengths(img.cuda())
    assert lengths.cpu().tolist() == want
    assert cu.cpu().tolist() == [0] + torch.tensor(want).cumsum(0).tolist()
    x, cu2, max_len = ops.strip_and_pack(img.cuda())
    ref = torch.cat([img[i, :n] for i, n in enumerate(want)]).bfloat16()
    assert torch.equal(x[:ref.shape[0]].cpu(), ref)
    x3, cu3, ml3 = ops.strip_and_pack(img.cuda(), lengths=want)
    assert torch.equal(x3.cpu(), ref) and ml3 == 10000 and torch.equal(cu3, cu2)


def test_reference_batch_layout_without_host_lengths_matches_packed_bags():
    """The reference batch layout (B,Npad,512) with -10000 row padding, lengths found on the device (no host
    sync): the packed buffer is sized for the worst case, so rows past cu[B] belong to no bag.  Tokens and every
    gradient must equal the run on exactly packed bags, also when the allocator hands out dirty memory."""
    from imp_b200 import ops
    from oracle import imp_oracle as O
    dev = "cuda"
    params = make_params(0)
    lens, npad, P = [300, 200, 417], 512, 16
    bags = [b.bfloat16().float() for b in make_bags(lens, 4)]
    img = torch.full((len(lens), npad, 512), O.SENTINEL)
    for i, b in enumerate(bags):
        img[i, :b.shape[0]] = b
    g = torch.Generator().manual_seed(2)
    p_proto = (torch.rand(1, P, 256, generator=g) * 2 - 1) / P
    cot = torch.randn(len(lens), P, 256, generator=g)

    def run(x, cu, max_len):
        leaves = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
        c, _ = ops.proto_fusion(x, cu, max_len, p_proto.to(dev), leaves["path_net.0.weight"], leaves["path_net.0.bias"],
                                [block_tensors(leaves, 0), block_tensors(leaves, 1)])
        (c * cot.to(dev)).sum().backward()
        return c.detach(), {k: v.grad for k, v in leaves.items()}

    x_ref = torch.cat(bags).to(dev).bfloat16().contiguous()
    c_ref, g_ref = run(x_ref, ops._cu_from_lengths(lens, dev), max(lens))
    poison = torch.full((64 << 20,), float("nan"), device=dev)      # dirty the allocator's free blocks
    del poison
    x, cu, max_len = ops.strip_and_pack(img.to(dev))                   # lengths from the device sentinel scan
    assert x.shape[0] == len(lens) * npad and cu.tolist() == [0, 300, 500, 917]
    c, gr = run(x, cu, max_len)
    assert rel(c, c_ref) < 1e-5
    for k in g_ref:
        assert torch.isfinite(gr[k]).all(), k
        assert rel(gr[k], g_ref[k]) < 1e-4, (k, rel(gr[k], g_ref[k]))
