"""CPU, build container only: the oracle against the reference executed live from /root/reference
(oracle/ref_harness.py).  Skipped where the reference tree is absent (the GPU box)."""
import warnings

import pytest
import torch

from oracle import imp_oracle as O
from oracle import ref_harness as R
from util_hotpath import rel

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
warnings.filterwarnings("ignore")


@pytest.mark.parametrize("n,p", [(64, 6), (500, 16), (700, 32)])
def test_compute_modularity_live(n, p):
    ops = R.load_ops()
    g = torch.Generator().manual_seed(n)
    x = torch.relu(torch.randn(1, n, 256, generator=g))
    c = torch.randn(1, p, 256, generator=g, requires_grad=True)
    with R.cpu_cuda_noop():
        ref = ops.compute_modularity(c, x)
    (gref,) = torch.autograd.grad(ref, c)
    loss, dc = O.modularity(c[0].detach(), x[0], chunk=200)
    # the loss is -100 x a difference of two O(1) traces: fp32 cancellation noise of the reference itself is ~1e-5
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item()) + 2e-5
    assert rel(dc, gref[0]) < 2e-4


def test_multihead_attention_live():
    ops = R.load_ops()
    torch.manual_seed(0)
    mha = ops.MultiheadAttention(embed_dim=256, num_heads=1)
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    c, h = torch.randn(7, 1, 256), torch.randn(300, 1, 256)
    out, raw = mha(c, h, h)
    o, s = O.cross_attention(c[:, 0], h[:, 0], mha.in_proj_weight.detach(), mha.in_proj_bias.detach(),
                             mha.out_proj.weight.detach(), mha.out_proj.bias.detach(), return_raw=True)
    assert rel(o, out[:, 0]) < 1e-5
    assert rel(s, raw.reshape(7, 300)) < 1e-5          # the module returns RAW logits (attention.py:535-538)


def test_sentinel_strip_live_semantics():
    """umeml_gan.py:404-409: first row-major hit of -10000 decides the row."""
    x = torch.randn(20, 8)
    x[11:] = -10000.0
    x[7, 5] = -10000.0
    idx = torch.nonzero(x == -10000)
    assert int(idx[0][0]) == O.bag_length(x) == 7


def test_register_into_the_reference_registry():
    """INTEGRATION.md option 0: ``register_into(MODEL_REGISTRY)`` puts this repo's factories under the reference's own
    registry names, and the reference's ``build_model`` then returns the B200 model class."""
    R.load_model_module()
    import importlib
    build = importlib.import_module("medmm.modeling.models.build")
    from imp_b200.registry import register_into
    from imp_b200.umeml_gan import UMEML_GAN
    register_into(build.MODEL_REGISTRY)
    assert {"umeml_gan", "umeml"} <= set(build.MODEL_REGISTRY.registered_names())
    from types import SimpleNamespace as NS
    cfg = NS(DATASET=NS(ROOT=".", PATH=NS(DIM=512), OMIC=NS(DIM=3354)),
             MODEL=NS(DROPOUT=0.25, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small", UMEML=NS(PROTOTYPES=6, REGISTERS=3)))
    model = build.build_model("umeml_gan", verbose=False, cfg=cfg, num_classes=4, omic_sizes=1000)
    assert isinstance(model, UMEML_GAN)
