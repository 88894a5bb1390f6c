"""CPU: the drop-in ``umeml_gan`` model (registry name, constructor, state_dict, token-level tail) against the fixture
produced by EXECUTING the reference UMEML_GAN with the same seeded parameters (tests/golden/model_P6_full.npz,
make_golden.gen_model_full).  The hot path needs the GPU (tests/test_model_gpu.py); here the tail is fed the reference's
own hot-path outputs, so everything after umeml_gan.py:434 is compared in isolation: Nystrom layers, bottleneck
pairing without host round trips, fusion MLP, explainers / importance / KD loss, the second pass, the cca tuple,
the GAN phase with its in-forward optimiser steps and the replace_ratio swap."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from util_hotpath import fill_state, rel

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    z = np.load(os.path.join(G, "model_P6_full.npz"))
    return {k: torch.from_numpy(z[k]) if z[k].dtype.kind in "fiu" else z[k] for k in z.files}


def make_cfg(n_proto=6, dropout=0.0, prec="fp32", groups=None):
    return SimpleNamespace(
        DATASET=SimpleNamespace(ROOT=".", PATH=SimpleNamespace(DIM=512), OMIC=SimpleNamespace(DIM=3354)),
        MODEL=SimpleNamespace(DROPOUT=dropout, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small",
                              UMEML=SimpleNamespace(PROTOTYPES=n_proto, REGISTERS=3, GENE_GROUP_INDEXES=groups)),
        TRAINER=SimpleNamespace(PREC=prec))


def build(z, tmp_path=None):
    from imp_b200.registry import build_model
    import imp_b200.umeml_gan  # noqa: F401  registers "umeml_gan"
    groups = [z["group_%d" % k].tolist() for k in range(6)]
    model = build_model("umeml_gan", verbose=False, cfg=make_cfg(groups=groups), num_classes=4, omic_sizes=1000)
    fill_state(model, int(z["param_seed"]))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.p_proto = z["p_proto_init"].clone()
    model.omic_means = z["omic_means"]
    return model


def test_registry_and_state_dict_contract():
    """Same registry name and kwargs as the reference (models/build.py:3-11, umeml_gan.py:704-706) and, at P = 6, the
    reference's full state_dict key set with identical shapes (g_omic_net is the harness-patched layer)."""
    z = load()
    model = build(z)
    ref = json.loads(str(z["state_keys"]))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert set(mine) == set(ref), (sorted(set(ref) - set(mine)), sorted(set(mine) - set(ref)))
    diff = {k: (mine[k], ref[k]) for k in ref if mine[k] != ref[k] and not k.startswith("g_omic_net")}
    assert not diff, diff
    assert mine["g_omic_net.0.weight"] == [256, 1000]                 # the reference's real shape (umeml_gan.py:285)
    from imp_b200.umeml_gan import UMEML_GAN
    assert isinstance(model, UMEML_GAN) and "p_proto" not in mine     # plain tensor as in the reference (:310-315)
    for attr in ("train_gan", "replace_ratio", "cca", "plot_set", "omic_means"):    # trainer-mutated attributes
        assert hasattr(model, attr)


def test_eval_tail_matches_reference_logits(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    z = load()
    model = build(z).eval()
    batch = {"patient_id": ["a", "b", "c"], "without_omic": z["without_omic"], "insample_without_omic": z["insample_without_omic"]}
    with torch.no_grad():
        logits = model.token_tail(z["eval.p_proto"], z["eval.h_omic_bag"], batch)
    assert rel(logits, z["eval.logits"]) < 1e-4, rel(logits, z["eval.logits"])
    rows = open(tmp_path / "train_path.txt").read().strip().splitlines()       # importance log, one row per slide (:576-587)
    assert len(rows) == 3 and len(rows[0].split()) == 6
    with torch.no_grad():                                                         # no omics at all: generator tokens (:506-507)
        logits = model.token_tail(z["eval.p_proto"], None, {"patient_id": None})
    assert rel(logits, z["eval_noomic.logits"]) < 1e-4


def test_cca_tuple(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    z = load()
    model = build(z).eval()
    model.cca = True
    with torch.no_grad():
        out = model.token_tail(z["cca.p_proto_before"], z["cca.h_omic_bag_before"], {"patient_id": None})
    assert len(out) == 5 and out[4] == "cca"
    assert rel(out[0], z["cca.h_path"]) < 1e-4 and rel(out[1], z["cca.h_omic"]) < 1e-4
    assert torch.equal(out[2], z["cca.p_proto_before"]) and torch.equal(out[3], z["cca.h_omic_bag_before"])


def _oracle_modularity_terms(bags_h):
    from oracle import imp_oracle as O

    def terms(h, cu, max_len, c_proto, c_omic=None, temp=0.1):
        rows = []
        for j, hb in enumerate(bags_h):
            vals = []
            for c in (c_proto[j], c_omic[j]):
                vals.append(O.modularity_literal(c, hb, temp))
            rows.append(torch.stack(vals))
        return torch.stack(rows)
    return terms


def test_train_tuple_and_gradients_of_the_tail(tmp_path, monkeypatch):
    """Train mode: (logits, modular_loss, 0, 0, 0, loss_kd, importance) and the gradients of NLL + KD + modularity into
    tail parameters.  The modularity term needs the kernels; on the CPU it is supplied by the oracle (test-only patch)."""
    monkeypatch.chdir(tmp_path)
    from imp_b200 import survival, umeml_gan as U
    from oracle import imp_oracle as O
    z = load()
    model = build(z).train()
    bags = O.strip_bags(z["img"])
    w1, b1 = model.path_net[0].weight.detach(), model.path_net[0].bias.detach()
    bags_h = [O.path_net(x, w1, b1) for x in bags]
    monkeypatch.setattr(U._mod, "modularity_terms", _oracle_modularity_terms(bags_h))
    p_proto = z["train.p_proto"].clone().requires_grad_(True)
    # the omic tokens before the omic encoder are the same in eval and train (dropout 0, no masks in the encoders)
    out = model.token_tail(p_proto, _h_omic_bag(model, z), {"patient_id": ["a", "b", "c"]}, hot={"h": None, "cu_seqlens": None, "max_len": 0})
    assert len(out) == 7 and out[2] == 0 and out[3] == 0 and out[4] == 0
    assert rel(out[0], z["train.logits"]) < 1e-4
    assert abs(out[1].item() - z["train.modular_loss"].item()) <= 1e-4 * abs(z["train.modular_loss"].item()) + 1e-5
    assert abs(out[5].item() - z["train.loss_kd"].item()) <= 1e-4 * abs(z["train.loss_kd"].item()) + 1e-7
    assert rel(out[6], z["train.importance_path"]) < 1e-4
    loss = survival.nll_loss_new(out, z["label"], z["censorship"]) + out[5] + out[1]
    assert abs(loss.item() - z["train.loss"].item()) <= 1e-4 * abs(z["train.loss"].item())
    loss.backward()
    named = dict(model.named_parameters())
    for k in ("classifier.weight", "bottleattn.linear_p.weight", "omic_encoder.0.attn.attn.to_qkv.weight", "explainer_path.weight",
              "p_encoder_token", "omic_net.4.0.weight"):
        if named[k].grad is None:
            continue
        assert rel(named[k].grad, z["train.grad." + k]) < 2e-3, (k, rel(named[k].grad, z["train.grad." + k]))


def _h_omic_bag(model, z):
    from oracle import imp_oracle as O
    groups = [z["group_%d" % k].tolist() for k in range(6)]
    ws = [m[0].weight for m in model.omic_net]
    bs = [m[0].bias for m in model.omic_net]
    return O.omic_encode(z["omic"], groups, ws, bs)


def test_gan_phase_steps_the_same_parameters(tmp_path, monkeypatch):
    """train_gan: generator + two discriminator Adam steps inside the forward (umeml_gan.py:461-490), then the
    replace_ratio swap with the numpy RNG (:492-497).  Losses and the stepped weights must match the reference."""
    monkeypatch.chdir(tmp_path)
    from imp_b200 import umeml_gan as U
    from oracle import imp_oracle as O
    z = load()
    model = build(z).train()
    bags = O.strip_bags(z["img"])
    bags_h = [O.path_net(x, model.path_net[0].weight.detach(), model.path_net[0].bias.detach()) for x in bags]
    monkeypatch.setattr(U._mod, "modularity_terms", _oracle_modularity_terms(bags_h))
    hot = {"h": None, "cu_seqlens": None, "max_len": 0}
    batch = {"patient_id": ["a", "b", "c"]}
    # the reference ran one plain train step before the GAN step: its Adam states are still fresh, parameters unchanged
    model.train_gan, model.replace_ratio = True, 0.5
    np.random.seed(1234)
    out = model.token_tail(z["train.p_proto"], _h_omic_bag(model, z).detach(), batch, hot=hot)
    for name, idx in (("gan.gen_loss", 2), ("gan.dis_p_loss", 3), ("gan.dis_o_loss", 4)):
        assert abs(float(out[idx]) - z[name].item()) <= 1e-4 * abs(z[name].item()), (name, float(out[idx]), z[name].item())
    assert rel(model.gan_generator_p2o.net[0].weight.detach()[:8, :64], z["gan.p2o_w0_after"]) < 1e-5
    assert rel(model.gan_discriminator_o.layers[0].weight.detach()[:8, :64], z["gan.dis_o_w0_after"]) < 1e-5
    assert rel(out[0], z["gan.logits"]) < 1e-3


def test_greedy_pairs_equals_sorted_scan():
    """Device-side pairing == the reference's host scan over the descending-sorted similarities (umeml_gan.py:174-186)."""
    from imp_b200.token_tail import greedy_pairs
    g = torch.Generator().manual_seed(0)
    sim = torch.randn(16, 9, 7, generator=g)
    ip, io = greedy_pairs(sim, 3)
    for b in range(sim.shape[0]):
        s = sim[b].numpy()
        order = np.argsort(s, axis=None)[::-1]
        used_p, used_o, pairs = set(), set(), []
        for idx in order:
            a, c = np.unravel_index(idx, s.shape)
            if a not in used_p and c not in used_o:
                pairs.append((a, c)); used_p.add(a); used_o.add(c)
            if len(pairs) == 3:
                break
        assert pairs == list(zip(ip[b].tolist(), io[b].tolist()))


def test_survival_loss_and_risk_match_reference_formulas():
    """loss/loss.py:28-95 and evaluation/evaluator.py:369-382 restated on the device-agnostic side."""
    from imp_b200 import survival
    from oracle import imp_oracle as O
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(5, 4, generator=g)
    y, c = torch.tensor([0, 1, 2, 3, 1]), torch.tensor([0, 1, 0, 1, 1])
    hz = torch.sigmoid(logits)
    S = torch.cumprod(1 - hz, dim=1)
    Sp = torch.cat([torch.ones(5, 1), S], 1)
    ref = (-(1 - c.view(-1, 1)) * (torch.log(Sp.gather(1, y.view(-1, 1)).clamp(min=1e-7)) + torch.log(hz.gather(1, y.view(-1, 1)).clamp(min=1e-7)))
           - c.view(-1, 1) * torch.log(Sp.gather(1, y.view(-1, 1) + 1).clamp(min=1e-7))).mean()
    assert abs(survival.nll_loss_new((logits,), y, c).item() - ref.item()) < 1e-6
    assert torch.allclose(survival.risk(logits), O.survival_risk(logits), atol=1e-6)


def test_umeml_registry_and_state_dict_contract():
    """The non-GAN ``umeml`` (models/umeml.py:83-221): same registry name, same state_dict keys and shapes, p_proto a
    Parameter."""
    from imp_b200.registry import build_model
    import imp_b200.umeml  # noqa: F401
    z = np.load(os.path.join(G, "umeml_P6_N200.npz"))
    cfg = make_cfg()
    cfg.DATASET.OMIC.DIM = 1000
    model = build_model("umeml", verbose=False, cfg=cfg, num_classes=4, omic_sizes=1000)
    ref = json.loads(str(z["state_keys"]))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref, (sorted(set(ref) ^ set(mine)), {k: (mine.get(k), ref.get(k)) for k in ref if mine.get(k) != ref.get(k)})
    assert isinstance(model.p_proto, torch.nn.Parameter)


@pytest.mark.parametrize("n", [7, 14, 40, 127])
def test_short_sequence_nystrom_equals_the_dense_form(n):
    """token_tail.nystrom_short (block algebra on the m - n zero-padded tokens) against the literal padded 128 x 128
    form (attention.py:46-161), outputs and gradients, in fp64."""
    from imp_b200 import token_tail as T
    torch.manual_seed(n)
    att = T.NystromAttention(dim=256, dim_head=32, heads=8, num_landmarks=128, pinv_iterations=6, residual=True, dropout=0.0).double()
    x = torch.randn(2, n, 256, dtype=torch.double, requires_grad=True)
    w = torch.randn(256, dtype=torch.double)
    y1, y2 = att(x), att.forward_dense(x)
    g1 = torch.autograd.grad((y1 * w).sum(), [x] + list(att.parameters()))
    g2 = torch.autograd.grad((y2 * w).sum(), [x] + list(att.parameters()))
    assert rel(y1, y2) < 1e-12
    assert max(rel(a, b) for a, b in zip(g1, g2)) < 1e-11
