"""CPU: the oracle (oracle/imp_oracle.py) against the golden vectors produced by EXECUTING the
unmodified reference (tests/golden/make_golden.py).  fp32 on both sides; tolerances are fp32
round-off of different summation orders."""
import os

import numpy as np
import pytest
import torch

from oracle import imp_oracle as O
from util_hotpath import block_tensors, make_omic_params, make_params, rel

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_proto_block_matches_reference():
    z = load("proto_block_P16_N333.npz")
    params = make_params(int(z["param_seed"]))
    names = ["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "norm1.weight", "norm1.bias"]
    blk = {n: t.clone().requires_grad_(True) for n, t in zip(names, block_tensors(params, 0))}
    h = z["in_h"][0].clone().requires_grad_(True)
    c = z["in_c"][0].clone().requires_grad_(True)
    out = O.proto_block(h, c, blk)
    assert rel(out, z["out"][0]) < 1e-5
    (out * z["cot"][0]).sum().backward()
    assert rel(h.grad, z["grad.h"][0]) < 1e-4
    assert rel(c.grad, z["grad.c"][0]) < 1e-4
    for n in names:
        key = "grad.cross_attn." + n if not n.startswith("norm1") else "grad." + n
        assert rel(blk[n].grad, z[key]) < 1e-4, n


def test_folded_query_is_exact_refactor():
    """S = q~ h^T + const_p reproduces softmax(q k^T)."""
    z = load("proto_block_P16_N333.npz")
    params = make_params(int(z["param_seed"]))
    in_w, in_b, out_w, out_b, _, _ = block_tensors(params, 0)
    h, c = z["in_h"][0], z["in_c"][0]
    ref = O.cross_attention(c, h, in_w, in_b, out_w, out_b)
    qt = O.folded_query(c, in_w, in_b)
    pooled, _ = O.lse_merge([O.pool_partial(h, qt)])
    d = 256
    alt = (pooled @ in_w[2 * d:].t() + in_b[2 * d:]) @ out_w.t() + out_b
    assert rel(alt, ref) < 1e-5


@pytest.mark.parametrize("name", ["modularity_P6_N300.npz", "modularity_P16_N512.npz", "modularity_P7_N257.npz"])
def test_modularity_matches_reference(name):
    z = load(name)
    x, c = z["x"][0], z["c"][0]
    lit = O.modularity_literal(c.clone().requires_grad_(True), x)
    assert abs(lit.item() - z["loss"].item()) <= 1e-5 * abs(z["loss"].item()) + 1e-6
    loss, dc = O.modularity(c, x, chunk=128)
    assert abs(loss.item() - z["loss"].item()) <= 2e-5 * abs(z["loss"].item()) + 1e-6
    assert rel(dc, z["grad_c"][0]) < 2e-4


def test_chain_matches_reference():
    z = load("chain_P16_N384.npz")
    params = make_params(int(z["param_seed"]))
    ref = O.hot_path_step([z["x"][0]], params, z["p_proto"][0], with_modularity=True, grad_seed=z["cot"])
    assert rel(ref["c"], z["c_out"]) < 1e-5
    assert abs(ref["modularity"].item() - z["modularity"].item()) <= 2e-5 * abs(z["modularity"].item())
    for k, g in ref["grads"].items():
        assert rel(g, z["grad." + k]) < 3e-4, (k, rel(g, z["grad." + k]))


def test_model_level_hot_path_matches_reference():
    """UMEML_GAN (P=6) eval forward with without_omic / insample masks: strip, path_net, prototype blocks,
    imputation + omic encoders, and the missing-omics blend."""
    z = load("model_P6_eval.npz")
    seed = int(z["param_seed"])
    params = make_params(seed)
    params.update(make_omic_params(seed))
    lens = [O.bag_length(z["img"][i]) for i in range(z["img"].shape[0])]
    assert lens == z["lens"].tolist()
    bags = O.strip_bags(z["img"])
    blocks = []
    for b in range(2):
        names = ["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "norm1.weight", "norm1.bias"]
        blocks.append(dict(zip(names, block_tensors(params, b))))
    for i, x in enumerate(bags):
        c, h = O.prototype_pool(x, z["p_proto_init"][0], params["path_net.0.weight"], params["path_net.0.bias"], blocks)
        assert rel(h, z["h_path_bag_%d" % i]) < 1e-5
        assert rel(c, z["p_proto_out"][i]) < 1e-5
    groups = [z["group_%d" % k].tolist() for k in range(6)]
    xo = O.impute_missing_genes(z["omic"], z["insample_without_omic"], z["omic_means"])
    ho = O.omic_encode(xo, groups, [params["omic_net.%d.0.weight" % k] for k in range(6)],
                       [params["omic_net.%d.0.bias" % k] for k in range(6)])
    assert rel(ho, z["h_omic_bag"]) < 1e-5
    post = O.blend_missing_omics(z["h_omic_pre"], z["h_omic_gen"], z["without_omic"], z["insample_without_omic"])
    assert rel(post, z["h_omic_post"]) < 1e-6


def test_distance_and_argmin_match_reference():
    z = load("distance_N96_K8.npz")
    a = O.kmeans_assign(z["x"], z["mu"])
    assert torch.equal(a.long(), z["argmin"].long())


def test_rounding_model_without_roundings_is_the_oracle():
    """oracle/rounding_model.py with no rounding point switched on must reproduce hot_path_step (and therefore
    the executed reference through the chain fixture); with all six on it must stay within the bf16 floor."""
    from oracle import rounding_model as R
    z = load("chain_P16_N384.npz")
    params = make_params(int(z["param_seed"]))
    exact = O.hot_path_step([z["x"][0].double()], {k: v.double() for k, v in params.items()}, z["p_proto"][0].double(),
                            with_modularity=False, grad_seed=z["cot"].double())
    plain = R.hot_path_step_rounded([z["x"][0]], params, z["p_proto"][0], z["cot"], points=())
    assert rel(plain["c"], exact["c"]) < 1e-12
    for k in exact["grads"]:
        assert rel(plain["grads"][k], exact["grads"][k]) < 1e-10, k
    rounded = R.hot_path_step_rounded([z["x"][0]], params, z["p_proto"][0], z["cot"])
    assert rel(rounded["c"], exact["c"]) < 1e-3
    worst = max(rel(rounded["grads"][k], exact["grads"][k]) for k in exact["grads"])
    assert 1e-5 < worst < 1e-2, worst
