"""Shared synthetic-input builders for the parity tests (SURVEY.md 8(d) shapes)."""
import math

import torch

PARAM_SHAPES = {
    "path_net.0.weight": (256, 512), "path_net.0.bias": (256,),
}
for _b in range(2):
    _p = "proto_g_blocks.%d." % _b
    PARAM_SHAPES.update({
        _p + "cross_attn.in_proj_weight": (768, 256), _p + "cross_attn.in_proj_bias": (768,),
        _p + "cross_attn.out_proj.weight": (256, 256), _p + "cross_attn.out_proj.bias": (256,),
        _p + "norm1.weight": (256,), _p + "norm1.bias": (256,),
    })


def make_params(seed=0, bias_scale=0.02):
    """Reference initialisers (nn.Linear default, xavier_uniform in_proj) with non-zero biases so that
    every gradient path is exercised."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in PARAM_SHAPES.items():
        if k.endswith("norm1.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            out[k] = bias_scale * torch.randn(shp, generator=g)
        else:
            bound = math.sqrt(6.0 / (shp[0] + shp[1])) if "in_proj" in k else 1.0 / math.sqrt(shp[1])
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return out


def make_bags(lens, seed=0):
    g = torch.Generator().manual_seed(seed + 100)
    return [torch.randn(n, 512, generator=g) for n in lens]


def block_tensors(params, b):
    p = "proto_g_blocks.%d." % b
    return [params[p + "cross_attn.in_proj_weight"], params[p + "cross_attn.in_proj_bias"],
            params[p + "cross_attn.out_proj.weight"], params[p + "cross_attn.out_proj.bias"],
            params[p + "norm1.weight"], params[p + "norm1.bias"]]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


GROUP_SIZES = [82, 330, 513, 440, 1538, 451]


def make_omic_params(seed=0):
    """omic_net.{k}.0.{weight,bias} (umeml_gan.py:274-283), deterministic from the seed."""
    g = torch.Generator().manual_seed(seed + 777)
    out = {}
    for k, s in enumerate(GROUP_SIZES):
        out["omic_net.%d.0.weight" % k] = (torch.rand(256, s, generator=g) * 2 - 1) / math.sqrt(s)
        out["omic_net.%d.0.bias" % k] = 0.05 * torch.randn(256, generator=g)
    return out


# --------------------------------------------------------------------------------------------
# parity ledger: every GPU parity test appends its measured errors here so that the numbers behind
# the asserts are visible (gpurun_out/r02_parity.jsonl on the GPU box -> profiles/r02_parity.md)
# --------------------------------------------------------------------------------------------
def record(test: str, quantity: str, err: float, tol: float, against: str, note: str = ""):
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r02_parity.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, "quantity": quantity, "err": float(err), "tol": float(tol),
                                "against": against, "note": note}) + "\n")
    except OSError:
        pass
    return err


def check(test: str, quantity: str, err: float, tol: float, against: str, note: str = ""):
    record(test, quantity, err, tol, against, note)
    assert err <= tol, "%s / %s: %.3e > %.1e (vs %s)" % (test, quantity, err, tol, against)


class Ledger:
    """Collects the checks of one test so that EVERY measured error reaches the parity ledger before the first
    failing one raises."""

    def __init__(self, test: str):
        self.test, self.bad = test, []

    def add(self, quantity: str, err: float, tol: float, against: str, note: str = ""):
        record(self.test, quantity, err, tol, against, note)
        if not err <= tol:
            self.bad.append("%s: %.3e > %.1e (vs %s)" % (quantity, err, tol, against))

    def assert_ok(self):
        assert not self.bad, "%s: %s" % (self.test, "; ".join(self.bad))


def fill_state(module, seed: int):
    """Deterministic parameters for EVERY entry of ``module.state_dict()`` from (seed, key name): used both on the
    executed reference (tests/golden/make_golden.py) and on this repo's drop-in model, so that a fixture only has to
    store the seed.  Weights U(-1/sqrt(fan_in), 1/sqrt(fan_in)), biases 0.02 N(0,1), LayerNorm gains 1 + 0.1 N(0,1),
    learned tokens U(0,1)."""
    import zlib
    sd = module.state_dict()
    out = {}
    for k in sorted(sd):
        t = sd[k]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        if not t.dtype.is_floating_point:
            out[k] = t.clone()
        elif k.endswith("_token") or k.endswith("bottle_tokens"):
            out[k] = torch.rand(t.shape, generator=g)
        elif t.dim() >= 2:
            fan_in = int(torch.tensor(t.shape[1:]).prod())
            out[k] = (torch.rand(t.shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
        elif "norm" in k and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        else:
            out[k] = 0.02 * torch.randn(t.shape, generator=g)
    module.load_state_dict(out, strict=True)
    return out
