"""Parity at the BASELINE.json shapes, checked by the oracle run in fp64 ON THE GPU (the oracle is plain torch and
follows its inputs' device): configs[1] = 16 384 x 512 patches, 32 prototypes + 7 omic tokens, forward and every
gradient; configs[3] = 120 000 patches, tokens, modularity loss and token gradients.

Two checkers per quantity (tests/util_hotpath.check writes every measured error to the parity ledger):
  * ``oracle``          imp_oracle.hot_path_step / modularity in fp64: the reference arithmetic.  Outputs (tokens,
                        losses) are held to the north-star 1e-3.  Gradients are held to the per-tensor floor that the
                        bf16 operands of the tensor-core path impose (FLOOR below; attribution in profiles/r02_parity.md).
  * ``rounding model``  oracle/rounding_model.py = the same step in fp64 with the six bf16 rounding points of the
                        device path switched on: everything, gradients included, is held to 1e-3 against it.
"""
import pytest
import torch

from util_hotpath import Ledger, block_tensors, check, make_bags, make_params, record, rel

pytestmark = pytest.mark.gpu

# Bounds against the FULL-PRECISION oracle (against the rounding model the bound is 1e-3 for every tensor).
# Every gradient of the prototype blocks meets the north-star 1e-3 (measured <= 7e-5).  dW1 / db1 are sums over
# patches of dz^T x, and dz = [h > 0] (E . G) is formed from bf16 tensor-core operands: the attribution test below
# measures what each rounding point costs them -- dpooled in G 1.7e-3 (the same rounded cotangent multiplies every
# patch of a bag, so it does not average out), E = [dS | a] 0.9e-3, dz storage 0.9e-3, q~ 0.3e-3, h 0.2e-3; 2.1e-3 in
# quadrature, which is what the kernels show.  FLOOR is that figure with margin; profiles/r02_parity.md has the table.
FLOOR = 3e-3


def _floor(name):
    return FLOOR if name.startswith("path_net") else 1e-3


def _floor_note(name):
    return "bf16 operand floor of dz (see attribution)" if name.startswith("path_net") else ""


def _device_step(bags, params, p_proto, cot):
    from imp_b200 import ops
    dev = "cuda"
    leaves = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
    lens = [b.shape[0] for b in bags]
    x = torch.cat(bags).to(dev).bfloat16().contiguous()
    cu = ops._cu_from_lengths(lens, dev)
    c, h = ops.proto_fusion(x, cu, max(lens), p_proto.to(dev), leaves["path_net.0.weight"], leaves["path_net.0.bias"],
                            [block_tensors(leaves, 0), block_tensors(leaves, 1)])
    (c * cot.to(dev)).sum().backward()
    torch.cuda.synchronize()
    return c.detach(), {k: v.grad for k, v in leaves.items()}, h, cu


def _inputs(lens, P, seed):
    params = make_params(seed)
    params["path_net.0.weight"] = params["path_net.0.weight"].bfloat16().float()
    bags = [b.bfloat16().float() for b in make_bags(lens, seed)]
    g = torch.Generator().manual_seed(seed + 5)
    p_proto = (torch.rand(1, P, 256, generator=g) * 2 - 1) / P
    cot = torch.randn(len(lens), P, 256, generator=g)
    return params, bags, p_proto, cot


@pytest.mark.parametrize("lens,P", [([16384], 32), ([16384, 9000], 32)])
def test_fusion_at_headline_shape(lens, P):
    from oracle import imp_oracle as O, rounding_model as R
    name = "headline_fusion[%s,P=%d]" % ("+".join(map(str, lens)), P)
    params, bags, p_proto, cot = _inputs(lens, P, 0)
    c, grads, _, _ = _device_step(bags, params, p_proto, cot)
    dev = "cuda"
    params64 = {k: v.to(dev).double() for k, v in params.items()}
    bags64 = [b.to(dev).double() for b in bags]
    exact = O.hot_path_step(bags64, params64, p_proto[0].to(dev).double(), with_modularity=False,
                            grad_seed=cot.to(dev).double())
    model = R.hot_path_step_rounded(bags64, params64, p_proto[0].to(dev), cot.to(dev))
    L = Ledger(name)
    L.add("tokens", rel(c, exact["c"]), 1e-3, "oracle fp64")
    L.add("tokens", rel(c, model["c"]), 1e-3, "rounding model")
    for k in sorted(grads):
        L.add("grad " + k, rel(grads[k], model["grads"][k]), 1e-3, "rounding model")
    for k in sorted(grads):
        L.add("grad " + k, rel(grads[k], exact["grads"][k]), _floor(k), "oracle fp64", _floor_note(k))
    L.assert_ok()


def test_rounding_point_attribution():
    """Which bf16 rounding point costs what: the rounding model with ONE point on, against the exact oracle
    (ledger only; the sum in quadrature reproduces the all-points error the kernels show)."""
    from oracle import imp_oracle as O, rounding_model as R
    params, bags, p_proto, cot = _inputs([4096], 32, 1)
    dev = "cuda"
    params64 = {k: v.to(dev).double() for k, v in params.items()}
    bags64 = [b.to(dev).double() for b in bags]
    exact = O.hot_path_step(bags64, params64, p_proto[0].to(dev).double(), with_modularity=False,
                            grad_seed=cot.to(dev).double())
    none = R.hot_path_step_rounded(bags64, params64, p_proto[0].to(dev), cot.to(dev), points=())
    check("attribution", "model without roundings == oracle (worst grad)",
          max(rel(none["grads"][k], exact["grads"][k]) for k in none["grads"]), 1e-9, "oracle fp64")
    for pt in R.ALL_POINTS + ("all",):
        m = R.hot_path_step_rounded(bags64, params64, p_proto[0].to(dev), cot.to(dev),
                                    points=R.ALL_POINTS if pt == "all" else (pt,))
        record("attribution[N=4096,P=32]", "tokens, point=" + pt, rel(m["c"], exact["c"]), 1e-3, "oracle fp64")
        for k in ("path_net.0.weight", "path_net.0.bias", "proto_g_blocks.0.cross_attn.in_proj_weight",
                  "proto_g_blocks.1.cross_attn.in_proj_weight", "proto_g_blocks.1.norm1.weight"):
            record("attribution[N=4096,P=32]", "grad %s, point=%s" % (k, pt), rel(m["grads"][k], exact["grads"][k]),
                   FLOOR, "oracle fp64")


def _modularity_case(n, p, q, seed, chunk):
    """Loss and token gradient of both token groups at full size, held to the north-star 1e-3 against the
    full-precision oracle at 16 384 and at 120 000 patches (a sweep CTA covers at most 16 384 columns, so the 32-bit
    fixed-point accumulators of T keep the same resolution on giant bags; before that cap the 120 000-patch gradient
    sat at 2.1e-3; profiles/r02_parity.md)."""
    from imp_b200 import modularity as M
    from oracle import imp_oracle as O
    L = Ledger("modularity[N=%d,P=%d+%d]" % (n, p, q))
    params, bags, p_proto, cot = _inputs([n], p, seed)
    c, _, h, cu = _device_step(bags, params, p_proto, cot)          # h: what path_net really produces (bf16, >= 0)
    g = torch.Generator().manual_seed(seed + 9)
    c2 = torch.randn(1, q, 256, generator=g).cuda()                 # signed: keeps the assignment away from saturation
    c1d = c.clone().requires_grad_(True)
    c2d = c2.clone().requires_grad_(True)
    loss = M.modularity_terms(h, cu, n, c1d, c2d)
    (loss[0, 0] + loss[0, 1]).backward()
    torch.cuda.synchronize()
    h64 = h.double()
    gtol = 1e-3
    for tag, cd, cref, li in (("proto", c1d, c[0].double(), 0), ("omic", c2d, c2[0].double(), 1)):
        for gram_bf16, who in ((False, "oracle fp64"), (True, "oracle fp64, bf16 Gram operand")):
            ref, dref = O.modularity(cref, h64, chunk=chunk, gram_bf16=gram_bf16)
            L.add("loss " + tag, abs(loss[0, li].item() - ref.item()) / max(abs(ref.item()), 1e-12), 1e-3, who)
            L.add("grad " + tag, rel(cd.grad[0], dref), gtol, who)
    L.assert_ok()


def test_modularity_at_headline_shape():
    _modularity_case(16384, 32, 7, 2, 1024)


def test_modularity_at_giant_shape():
    _modularity_case(120000, 32, 7, 3, 2048)


def test_fusion_at_giant_shape():
    """120 000 patches: tokens and the gradients of the prototype blocks against the fp64 oracle."""
    from oracle import imp_oracle as O, rounding_model as R
    params, bags, p_proto, cot = _inputs([120000], 32, 4)
    c, grads, _, _ = _device_step(bags, params, p_proto, cot)
    dev = "cuda"
    params64 = {k: v.to(dev).double() for k, v in params.items()}
    bags64 = [b.to(dev).double() for b in bags]
    exact = O.hot_path_step(bags64, params64, p_proto[0].to(dev).double(), with_modularity=False,
                            grad_seed=cot.to(dev).double())
    model = R.hot_path_step_rounded(bags64, params64, p_proto[0].to(dev), cot.to(dev))
    L = Ledger("giant_fusion[N=120000]")
    L.add("tokens", rel(c, exact["c"]), 1e-3, "oracle fp64")
    for k in sorted(grads):
        L.add("grad " + k, rel(grads[k], model["grads"][k]), 1e-3, "rounding model")
        L.add("grad " + k, rel(grads[k], exact["grads"][k]), _floor(k), "oracle fp64", _floor_note(k))
    L.assert_ok()
