"""Generate the golden fixtures in this directory by EXECUTING the unmodified reference
(/root/reference, through oracle/ref_harness.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own; these files pin the oracle
(tests/test_oracle_golden.py) and, through it, the CUDA kernels.  Everything is fp32 on CPU with
dropout 0 (the reference RNG stream of nn.Dropout cannot be matched)."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_harness as R  # noqa: E402
from util_hotpath import make_omic_params, make_params  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_num_threads(8)


def npz(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def sd_of(module, prefix=""):
    return {prefix + k: v.detach().clone() for k, v in module.state_dict().items()}


def gen_proto_block(mod, P, N, seed):
    """PathProtoGenerator (umeml_gan.py:65-80) forward + backward."""
    torch.manual_seed(seed)
    blk = mod.PathProtoGenerator(dim=256)
    params = make_params(seed)                               # weights are NOT stored: rebuilt from the seed
    blk.load_state_dict({k[len("proto_g_blocks.0."):]: v for k, v in params.items() if k.startswith("proto_g_blocks.0.")})
    h = torch.relu(torch.randn(1, N, 256)).requires_grad_(True)
    c = ((torch.rand(1, P, 256) * 2 - 1) / P).requires_grad_(True)
    out = blk(h, c)
    cot = torch.randn_like(out)
    grads = torch.autograd.grad((out * cot).sum(), [h, c] + list(blk.parameters()))
    names = ["h", "c"] + [n for n, _ in blk.named_parameters()]
    arrs = {"in_h": h, "in_c": c, "out": out, "cot": cot, "param_seed": seed}
    arrs.update({"grad." + n: g for n, g in zip(names, grads)})
    npz("proto_block_P%d_N%d.npz" % (P, N), **arrs)


def gen_modularity(ops, P, N, seed):
    """compute_modularity (ops/utils.py:205-228) value and gradient wrt c."""
    torch.manual_seed(seed)
    base = torch.randn(6, 256)
    x = torch.relu((torch.rand(N, 6) ** 3) @ base + 0.3 * torch.randn(N, 256)).unsqueeze(0)
    c = torch.randn(1, P, 256, requires_grad=True)
    with R.cpu_cuda_noop():
        loss = ops.compute_modularity(c, x)
    (g,) = torch.autograd.grad(loss, c)
    npz("modularity_P%d_N%d.npz" % (P, N), x=x, c=c, loss=loss, grad_c=g)


def gen_chain(mod, ops, P, N, seed):
    """path_net -> 2 x PathProtoGenerator -> compute_modularity, the ops-level hot path of
    umeml_gan.py:410,425-434,520 at a P the shipped model cannot run."""
    torch.manual_seed(seed)
    path_net = torch.nn.Sequential(torch.nn.Linear(512, 256), torch.nn.ReLU(), torch.nn.Dropout(0.0))
    blocks = torch.nn.ModuleList([mod.PathProtoGenerator(dim=256) for _ in range(2)])
    pp = make_params(seed)
    path_net.load_state_dict({k[len("path_net."):]: v for k, v in pp.items() if k.startswith("path_net.")})
    blocks.load_state_dict({k[len("proto_g_blocks."):]: v for k, v in pp.items() if k.startswith("proto_g_blocks.")})
    x = torch.randn(1, N, 512)
    p_proto = (torch.rand(1, P, 256) * 2 - 1) / P
    h = path_net(x)
    c = p_proto
    for b in blocks:
        c = b(h, c)
    with R.cpu_cuda_noop():
        mod_loss = ops.compute_modularity(c, h)
    cot = torch.randn_like(c)
    total = (c * cot).sum() + mod_loss
    params = dict(list(path_net.named_parameters(prefix="path_net")) + list(blocks.named_parameters(prefix="proto_g_blocks")))
    grads = torch.autograd.grad(total, list(params.values()))
    arrs = {"x": x, "p_proto": p_proto, "cot": cot, "c_out": c, "modularity": mod_loss, "param_seed": seed}
    arrs.update({"grad." + k: g for k, g in zip(params.keys(), grads)})
    npz("chain_P%d_N%d.npz" % (P, N), **arrs)


def gen_model(seed):
    """UMEML_GAN (P = 6) eval forward on a padded batch with missing-omics masks; hot-path
    intermediates captured with hooks (umeml_gan.py:380-434,500-511)."""
    model = R.build_reference_model(seed=seed, dropout=0.0)
    pp = make_params(seed)
    pp.update(make_omic_params(seed))
    missing = model.load_state_dict(pp, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    model.eval()
    torch.manual_seed(seed + 1)
    lens, npad, G = [96, 57, 120], 128, sum(R.GROUP_SIZES)   # every bag shorter than npad: the reference is undefined otherwise (umeml_gan.py:405-410)
    img = torch.full((len(lens), npad, 512), -10000.0)
    for i, n in enumerate(lens):
        img[i, :n] = torch.randn(n, 512)
    omic = torch.rand(len(lens), G)
    model.omic_means = torch.rand(G)
    without = torch.tensor([0, 1, 0])
    insample = (torch.rand(len(lens), G) < 0.3).int()
    cap = {}
    hooks = [
        model.path_net.register_forward_hook(lambda m, i, o: cap.setdefault("h_path_bag", []).append(o.detach().clone())),
        model.proto_g_blocks[1].register_forward_hook(lambda m, i, o: cap.setdefault("p_proto", []).append(o.detach().clone())),
        model.layer_norm_o.register_forward_hook(lambda m, i, o: cap.__setitem__("h_omic_pre", o.detach().clone())),
        model.gan_generator_p2o.register_forward_hook(lambda m, i, o: cap.__setitem__("h_omic_gen", o.detach().clone())),
        model.bottleattn.register_forward_pre_hook(lambda m, a: cap.__setitem__("h_omic_post", cap.get("h_omic_post", a[1].detach().clone()))),
    ]
    for k, net in enumerate(model.omic_net):
        hooks.append(net.register_forward_hook(lambda m, i, o, k=k: cap.__setitem__("omic_%d" % k, o.detach().clone())))
    batch = {"img": img, "omic": omic, "patient_id": ["a", "b", "c"], "without_omic": without,
             "insample_without_omic": insample}
    with torch.no_grad():
        logits = R.run_reference_forward(model, batch, train=False)
    for h in hooks:
        h.remove()
    arrs = {"img": img, "omic": omic, "omic_means": model.omic_means, "without_omic": without,
            "insample_without_omic": insample, "p_proto_init": model.p_proto, "logits": logits,
            "lens": np.array(lens), "p_proto_out": torch.cat(cap["p_proto"][-len(lens):], 0),
            "h_omic_bag": torch.cat([cap["omic_%d" % k] for k in range(6)], dim=1),
            "h_omic_pre": cap["h_omic_pre"], "h_omic_gen": cap["h_omic_gen"], "h_omic_post": cap["h_omic_post"]}
    for i in range(len(lens)):
        arrs["h_path_bag_%d" % i] = cap["h_path_bag"][i][0]
    arrs["param_seed"] = seed
    groups = [np.array(ix) for ix in model.gene_group_indexes]
    for k, ix in enumerate(groups):
        arrs["group_%d" % k] = ix
    npz("model_P6_eval.npz", **arrs)


def gen_model_full(seed):
    """The WHOLE UMEML_GAN (P = 6) with every parameter set by util_hotpath.fill_state(seed): eval forward with masks,
    train forward (7-tuple) + the gradients of NLL + KD + modularity, the cca tuple, and one GAN phase (in-forward
    optimiser steps).  Nystrom output dropout (hard-coded 0.1, ops/blocks.py:262) is switched off on the module for
    the train-mode runs so that they are deterministic."""
    import json
    from util_hotpath import fill_state
    model = R.build_reference_model(seed=seed, dropout=0.0)
    fill_state(model, seed)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    torch.manual_seed(seed + 1)
    lens, npad, G = [96, 57, 120], 128, sum(R.GROUP_SIZES)
    img = torch.full((len(lens), npad, 512), -10000.0)
    for i, n in enumerate(lens):
        img[i, :n] = torch.randn(n, 512)
    omic = torch.rand(len(lens), G)
    model.omic_means = torch.rand(G)
    without = torch.tensor([0, 1, 0])
    insample = (torch.rand(len(lens), G) < 0.3).int()
    label, cens = torch.tensor([1, 3, 0]), torch.tensor([0, 1, 0])
    arrs = {"img": img, "omic": omic, "omic_means": model.omic_means, "without_omic": without, "insample_without_omic": insample,
            "label": label, "censorship": cens, "p_proto_init": model.p_proto.clone(), "param_seed": seed, "lens": np.array(lens)}
    for k, ix in enumerate(model.gene_group_indexes):
        arrs["group_%d" % k] = np.array(ix)
    arrs["state_keys"] = np.array(json.dumps({k: list(v.shape) for k, v in model.state_dict().items()}))
    cap = {}
    hooks = [model.proto_g_blocks[1].register_forward_hook(lambda m, i, o: cap.setdefault("p", []).append(o.detach().clone())),
             model.layer_norm_p.register_forward_hook(lambda m, i, o: cap.__setitem__("h_path", o.detach().clone())),
             model.layer_norm_o.register_forward_hook(lambda m, i, o: cap.__setitem__("h_omic", o.detach().clone()))]
    for k, net in enumerate(model.omic_net):
        hooks.append(net.register_forward_hook(lambda m, i, o, k=k: cap.__setitem__("omic_%d" % k, o.detach().clone())))
    # (1) eval with both kinds of masks
    with torch.no_grad():
        logits = R.run_reference_forward(model, {"img": img, "omic": omic, "patient_id": ["a", "b", "c"], "without_omic": without,
                                                 "insample_without_omic": insample}, train=False)
    arrs.update({"eval.logits": logits, "eval.p_proto": torch.cat(cap["p"][-len(lens):], 0), "eval.h_path": cap["h_path"],
                 "eval.h_omic": cap["h_omic"], "eval.h_omic_bag": torch.cat([cap["omic_%d" % k] for k in range(6)], dim=1)})
    # (2) eval without omics at all (generator output replaces the omic tokens, umeml_gan.py:506-507)
    with torch.no_grad():
        arrs["eval_noomic.logits"] = R.run_reference_forward(model, {"img": img, "omic": None, "patient_id": ["a", "b", "c"],
                                                                     "insample_without_omic": torch.zeros(len(lens), G)}, train=False)
    # (3) train step: loss = NLL + KD + modularity (mbtrain.py:182-186), gradients
    cap.clear()
    model.zero_grad()
    out = R.run_reference_forward(model, {"img": img, "omic": omic, "patient_id": ["a", "b", "c"]}, train=True)
    loss_mod = R.load_model_module()
    import importlib
    nll = importlib.import_module("medmm.loss.loss").nll_loss_new
    loss_nomod = nll(logits=out, Y=label, c=cens) + out[-2]
    loss_nomod.backward(retain_graph=True)                     # the step without the ill-conditioned small-bag modularity term
    names = ("classifier.weight", "path_net.0.weight", "bottleattn.linear_p.weight", "omic_encoder.0.attn.attn.to_qkv.weight",
             "proto_g_blocks.1.cross_attn.in_proj_weight", "omic_net.4.0.weight", "explainer_path.weight", "p_encoder_token")
    for k in names:
        arrs["train.grad_nomod." + k] = dict(model.named_parameters())[k].grad.clone()
    model.zero_grad()
    loss = loss_nomod + out[1]
    loss.backward()
    arrs.update({"train.logits": out[0], "train.modular_loss": out[1], "train.loss_kd": out[5], "train.importance_path": out[6],
                 "train.loss": loss, "train.p_proto": torch.cat(cap["p"][-len(lens):], 0), "train.h_omic": cap["h_omic"]})
    for k in ("classifier.weight", "path_net.0.weight", "bottleattn.linear_p.weight", "omic_encoder.0.attn.attn.to_qkv.weight",
              "proto_g_blocks.1.cross_attn.in_proj_weight", "omic_net.4.0.weight", "explainer_path.weight", "p_encoder_token"):
        arrs["train.grad." + k] = dict(model.named_parameters())[k].grad.clone()
    # (4) cca tuple
    model.cca = True
    with torch.no_grad():
        cca = R.run_reference_forward(model, {"img": img, "omic": omic, "patient_id": ["a", "b", "c"]}, train=False)
    model.cca = False
    arrs.update({"cca.h_path": cca[0], "cca.h_omic": cca[1], "cca.p_proto_before": cca[2], "cca.h_omic_bag_before": cca[3]})
    # (5) GAN phase + replace_ratio swap: three in-forward optimiser steps, numpy RNG for the swap
    model.train_gan, model.replace_ratio = True, 0.5
    np.random.seed(1234)
    out = R.run_reference_forward(model, {"img": img, "omic": omic, "patient_id": ["a", "b", "c"]}, train=True)
    model.train_gan, model.replace_ratio = False, 0
    arrs.update({"gan.gen_loss": out[2], "gan.dis_p_loss": out[3], "gan.dis_o_loss": out[4], "gan.logits": out[0],
                 "gan.p2o_w0_after": model.gan_generator_p2o.net[0].weight.detach()[:8, :64].clone(),
                 "gan.dis_o_w0_after": model.gan_discriminator_o.layers[0].weight.detach()[:8, :64].clone()})
    for h in hooks:
        h.remove()
    npz("model_P6_full.npz", **arrs)


def gen_umeml(seed):
    """The non-GAN UMEML (models/umeml.py:83-221), one slide per batch as the reference requires: eval logits, the
    train pair (logits, modular_loss) and gradients incl. the learnable prototypes."""
    import importlib
    import json
    from types import SimpleNamespace as NS
    from util_hotpath import fill_state
    R.load_model_module()
    mod = importlib.import_module("medmm.modeling.models.umeml")
    cfg = NS(DATASET=NS(ROOT=".", PATH=NS(DIM=512), OMIC=NS(DIM=1000)),
             MODEL=NS(DROPOUT=0.0, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small", UMEML=NS(PROTOTYPES=6, REGISTERS=3)))
    torch.manual_seed(seed)
    model = mod.UMEML(cfg, num_classes=4, omic_sizes=1000).float()
    fill_state(model, seed)
    with torch.no_grad():
        model.p_proto.uniform_(-1.0 / 6, 1.0 / 6, generator=torch.Generator().manual_seed(seed + 3))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn(1, 200, 512, generator=g)
    omic = torch.rand(1, 1000, generator=g)
    arrs = {"img": img, "omic": omic, "param_seed": seed, "p_proto": model.p_proto.detach().clone(),
            "state_keys": np.array(json.dumps({k: list(v.shape) for k, v in model.state_dict().items()}))}
    model.eval()
    with torch.no_grad(), R.cpu_cuda_noop():
        arrs["eval.logits"] = model({"img": img, "omic": omic})
    model.train()
    with R.cpu_cuda_noop():
        logits, modular = model({"img": img, "omic": omic})
    loss = logits.square().sum() + modular
    loss_nomod = logits.square().sum()
    names = ("p_proto", "path_net.0.weight", "omic_net.0.weight", "classifier.weight", "bottleattn.bottle_tokens")
    loss_nomod.backward(retain_graph=True)
    for k in names:
        arrs["train.grad_nomod." + k] = dict(model.named_parameters())[k].grad.clone()
    model.zero_grad()
    loss.backward()
    arrs.update({"train.logits": logits, "train.modular_loss": modular})
    for k in names:
        arrs["train.grad." + k] = dict(model.named_parameters())[k].grad.clone()
    npz("umeml_P6_N200.npz", **arrs)


def gen_distance(seed):
    """euclidean_squared_distance (metrics/distance.py:46-61): the only reference arithmetic behind A9."""
    dist = R.load_distance()
    torch.manual_seed(seed)
    x, mu = torch.randn(96, 512), torch.randn(8, 512)
    d = dist.euclidean_squared_distance(x, mu)
    npz("distance_N96_K8.npz", x=x, mu=mu, dist=d, argmin=d.argmin(dim=1))


if __name__ == "__main__":
    ops = R.load_ops()
    mod = R.load_model_module()
    gen_proto_block(mod, 16, 333, 1)
    gen_modularity(ops, 6, 300, 2)
    gen_modularity(ops, 16, 512, 3)
    gen_modularity(ops, 7, 257, 4)
    gen_chain(mod, ops, 16, 384, 5)
    gen_model(6)
    gen_model_full(8)
    gen_umeml(9)
    gen_distance(7)
