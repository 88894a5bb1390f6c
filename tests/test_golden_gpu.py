"""The CUDA path on the committed golden fixtures: tests/golden/*.npz hold inputs and outputs of the EXECUTED,
unmodified reference (tests/golden/make_golden.py).  They are fed straight into the reference-signature modules
(``ops.PathProtoGenerator``, ``ops.MultiheadAttention`` incl. the raw logits, ``modularity.compute_modularity``,
``model.IMPHotPath``) and the results compared with what the reference produced.

The fixtures are fp32 and NOT bf16-representable, so unlike the oracle tests the comparison also pays for
rounding the inputs (patch tokens, W1) to bf16: outputs are held to the north-star 1e-3, gradients to the
bf16-operand floor (see tests/test_headline_parity_gpu.py); every measured error goes to the parity ledger."""
import os

import numpy as np
import pytest
import torch

from util_hotpath import Ledger, block_tensors, check, make_omic_params, make_params, rel

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLOOR = 5e-3


def load(name):
    z = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _load_block(mod, params, b):
    in_w, in_b, out_w, out_b, ln_w, ln_b = block_tensors(params, b)
    with torch.no_grad():
        mod.cross_attn.in_proj_weight.copy_(in_w); mod.cross_attn.in_proj_bias.copy_(in_b)
        mod.cross_attn.out_proj.weight.copy_(out_w); mod.cross_attn.out_proj.bias.copy_(out_b)
        mod.norm1.weight.copy_(ln_w); mod.norm1.bias.copy_(ln_b)


def test_path_proto_generator_module_on_golden():
    """umeml_gan.py:65-80 through the module with the reference signature: forward, and the gradients into the
    patch tokens h, the prototypes c and every parameter."""
    from imp_b200 import ops
    z = load("proto_block_P16_N333.npz")
    params = make_params(int(z["param_seed"]))
    mod = ops.PathProtoGenerator(256).cuda()
    _load_block(mod, params, 0)
    h = z["in_h"].cuda().requires_grad_(True)
    c = z["in_c"].cuda().requires_grad_(True)
    out = mod(h, c)
    (out * z["cot"].cuda()).sum().backward()
    name = "golden proto_block_P16_N333"
    check(name, "out", rel(out, z["out"]), 1e-3, "executed reference")
    check(name, "grad h", rel(h.grad, z["grad.h"]), FLOOR, "executed reference", "bf16 operand floor")
    check(name, "grad c", rel(c.grad, z["grad.c"]), FLOOR, "executed reference", "bf16 operand floor")
    got = {"cross_attn.in_proj_weight": mod.cross_attn.in_proj_weight.grad, "cross_attn.in_proj_bias": mod.cross_attn.in_proj_bias.grad,
           "cross_attn.out_proj.weight": mod.cross_attn.out_proj.weight.grad, "cross_attn.out_proj.bias": mod.cross_attn.out_proj.bias.grad,
           "norm1.weight": mod.norm1.weight.grad, "norm1.bias": mod.norm1.bias.grad}
    for k, g in got.items():
        ref = z["grad." + k]
        if k == "cross_attn.in_proj_bias":
            # the key bias shifts every logit of a prototype equally and cancels in the softmax: the reference
            # gradient of that slice is round-off noise around zero (|.| ~ 1e-9); compare the q and v slices
            g = torch.cat([g[:256], g[512:]]); ref = torch.cat([ref[:256], ref[512:]])
        check(name, "grad " + k, rel(g, ref), FLOOR, "executed reference", "bf16 operand floor")


def test_multihead_attention_module_and_raw_logits_on_golden():
    """blocks.py:441-526 / attention.py:236-547: (L,B,E) layout, returns (attn_output, RAW pre-softmax logits)."""
    from imp_b200 import ops
    from oracle import imp_oracle as O
    z = load("proto_block_P16_N333.npz")
    params = make_params(int(z["param_seed"]))
    in_w, in_b, out_w, out_b, _, _ = block_tensors(params, 0)
    mha = ops.MultiheadAttention(256, 1).cuda()
    with torch.no_grad():
        mha.in_proj_weight.copy_(in_w); mha.in_proj_bias.copy_(in_b)
        mha.out_proj.weight.copy_(out_w); mha.out_proj.bias.copy_(out_b)
    q = z["in_c"].transpose(0, 1).cuda()          # (L=P, B=1, E)
    k = z["in_h"].transpose(0, 1).cuda()          # (S=N, B=1, E)
    out, raw = mha(q, k, k)
    ref_o, ref_s = O.cross_attention(z["in_c"][0], z["in_h"][0], in_w, in_b, out_w, out_b, return_raw=True)
    assert out.shape == (16, 1, 256) and raw.shape == (1, 1, 16, 333)
    check("golden mha", "attn_output", rel(out[:, 0], ref_o), 1e-3, "oracle (pinned by proto_block golden)")
    check("golden mha", "raw logits", rel(raw[0, 0], ref_s), 1e-3, "oracle (pinned by proto_block golden)")


@pytest.mark.parametrize("name", ["modularity_P6_N300.npz", "modularity_P16_N512.npz", "modularity_P7_N257.npz"])
def test_compute_modularity_on_golden(name):
    """ops/utils.py:205-228 with the reference signature compute_modularity(c (1,P,D), x (1,N,D))."""
    from imp_b200 import modularity as M
    z = load(name)
    c = z["c"].cuda().requires_grad_(True)
    loss = M.compute_modularity(c, z["x"].cuda())
    loss.backward()
    ref = z["loss"].item()
    # -100 x the difference of two traces in [0,1]: 1e-3 relative + 5e-5 absolute (tests/test_modularity_gpu.py)
    err = max(0.0, abs(loss.item() - ref) - 5e-5) / abs(ref)
    L = Ledger("golden " + name)
    # the fixture's x is signed fp32 noise on a few hundred patches: the loss is the difference of two nearly equal
    # traces and x is rounded to bf16 on entry, so the bounds are those of the small-graph oracle tests
    # (tests/test_modularity_gpu.py: 1e-2 on the gradient) rather than the full-size ones
    L.add("loss", err, 3e-3, "executed reference", "x rounded to bf16 on entry; small signed graph")
    L.add("grad c", rel(c.grad, z["grad_c"]), 1e-2, "executed reference", "x rounded to bf16 on entry; small signed graph")
    # the same call against the oracle on the bf16-rounded x (what the device actually sees)
    from oracle import imp_oracle as O
    xr = z["x"][0].bfloat16().float()
    ref2, dref2 = O.modularity(z["c"][0], xr, chunk=128)
    L.add("loss", max(0.0, abs(loss.item() - ref2.item()) - 5e-5) / abs(ref2.item()), 1e-3, "oracle on bf16-rounded x")
    L.add("grad c", rel(c.grad[0], dref2), 1e-2, "oracle on bf16-rounded x", "small signed graph")
    L.assert_ok()


def test_chain_on_golden():
    """path_net -> two prototype blocks -> modularity, every gradient (the reference run of make_golden.py)."""
    from imp_b200 import modularity as M, ops
    z = load("chain_P16_N384.npz")
    params = make_params(int(z["param_seed"]))
    dev = "cuda"
    leaves = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
    x = z["x"][0].to(dev).bfloat16().contiguous()
    cu = torch.tensor([0, 384], dtype=torch.int32, device=dev)
    c, h = ops.proto_fusion(x, cu, 384, z["p_proto"].to(dev), leaves["path_net.0.weight"], leaves["path_net.0.bias"],
                            [block_tensors(leaves, 0), block_tensors(leaves, 1)])
    mod = M.modularity_terms(h, cu, 384, c)[0, 0]
    ((c * z["cot"].to(dev)).sum() + mod).backward()
    L = Ledger("golden chain_P16_N384")
    L.add("tokens", rel(c, z["c_out"]), 1e-3, "executed reference")
    ref = z["modularity"].item()
    L.add("modularity", max(0.0, abs(mod.item() - ref) - 5e-5) / abs(ref), 1e-3, "executed reference")
    # fp32 fixture vs bf16 device inputs: rounding x and W1 flips the ReLU mask of the few activations next to zero, each
    # flip switches a full-size dz entry (relative Frobenius error of dW1 ~ sqrt(flipped fraction) = 3-4e-2, independent
    # of the bag size; see tests/test_model_gpu.py); the other tensors see the small-graph modularity gradient (1e-2 class)
    for k, v in leaves.items():
        L.add("grad " + k, rel(v.grad, z["grad." + k]), 5e-2, "executed reference", "fp32 fixture rounded to bf16; modularity-dominated cotangent on 384 patches")
    # the pooling path alone on the same fixture inputs (no modularity): oracle on the bf16-rounded inputs
    from oracle import imp_oracle as O
    pr = dict(params); pr["path_net.0.weight"] = pr["path_net.0.weight"].bfloat16().float()
    exact = O.hot_path_step([z["x"][0].bfloat16().float()], pr, z["p_proto"][0], with_modularity=False, grad_seed=z["cot"])
    leaves2 = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
    c2, _ = ops.proto_fusion(x, cu, 384, z["p_proto"].to(dev), leaves2["path_net.0.weight"], leaves2["path_net.0.bias"],
                             [block_tensors(leaves2, 0), block_tensors(leaves2, 1)])
    (c2 * z["cot"].to(dev)).sum().backward()
    for k, v in leaves2.items():
        L.add("grad " + k + " (no modularity)", rel(v.grad, exact["grads"][k]), 3e-3 if k.startswith("path_net") else 1e-3,
              "oracle on the bf16-rounded fixture")
    L.assert_ok()


def test_model_hot_path_on_golden():
    """UMEML_GAN (P = 6) eval forward of the executed reference: sentinel strip, path_net, prototype blocks,
    imputation + six omic encoders (with insample masks) through ``IMPHotPath``."""
    from imp_b200 import kernels, model as M
    z = load("model_P6_eval.npz")
    seed = int(z["param_seed"])
    params = make_params(seed)
    params.update(make_omic_params(seed))
    groups = [z["group_%d" % k].tolist() for k in range(6)]
    net = M.IMPHotPath(n_proto=6, dropout=0.25, gene_group_indexes=groups).cuda().eval()
    missing, unexpected = net.load_state_dict({k: v for k, v in params.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    net.p_proto = z["p_proto_init"].cuda()
    net.omic_means = z["omic_means"].cuda()
    out = net({"img": z["img"].cuda(), "omic": z["omic"].cuda(),
               "insample_without_omic": z["insample_without_omic"].cuda()})
    name = "golden model_P6_eval"
    assert (out["cu_seqlens"][1:] - out["cu_seqlens"][:-1]).tolist() == z["lens"].tolist()
    check(name, "p_proto_out", rel(out["p_proto"], z["p_proto_out"]), 1e-3, "executed reference")
    check(name, "h_omic_bag", rel(out["h_omic_bag"], z["h_omic_bag"]), 1e-5, "executed reference")
    o = 0
    for i, n in enumerate(z["lens"].tolist()):
        check(name, "h_path_bag_%d" % i, rel(out["h"][o:o + n].float(), z["h_path_bag_%d" % i]), 4e-3, "executed reference",
              "h is stored in bf16 (half-ulp 2^-9 = 2e-3 per element)")
        o += n
    post, _ = kernels.omic_blend(z["h_omic_pre"].cuda(), z["h_omic_gen"].cuda(), z["without_omic"].to(torch.int32).cuda(),
                                 z["insample_without_omic"].to(torch.int32).cuda())
    check(name, "h_omic_post", rel(post, z["h_omic_post"]), 1e-6, "executed reference")
