"""Loader for the UNMODIFIED reference (helenypzhang/Interpretable-Multimodal-Prototyping).

TEST INFRASTRUCTURE ONLY.  This file is used in the build container (where
``/root/reference`` is mounted read-only) to
  * validate the restated oracle in ``oracle/imp_oracle.py`` against the real code, and
  * generate the committed fixtures under ``tests/golden/`` (``tests/golden/make_golden.py``).
Nothing in the product package imports it, and nothing that runs on the GPU box
may call it: the reference tree does not exist there (``available()`` is False).

The reference does not import as-is; the harness
supplies, without touching the reference sources:
  1. ``sys.modules`` stubs (with a ModuleSpec) for packages that are absent here:
     timm (DropPath -> identity; the reference uses drop_path=0), ftfy, h5py,
     matplotlib, sksurv, yacs, openpyxl, and the HF-hub dependent
     ``extract_prototype_with_plip_train`` (returns zeros(1,6,256); the model
     overwrites it with uniform noise anyway, umeml_gan.py:310-315);
  2. a nested ``SimpleNamespace`` cfg with the keys ``UMEML_GAN.__init__`` reads
     (umeml_gan.py:243-262);
  3. synthetic CSV fixtures in a scratch CWD so that the six gene groups have
     exactly [82,330,513,440,1538,451] members (umeml_gan.py:274,347-369);
  4. PATCH 1: ``g_omic_net`` -> Linear(G,256) (its output is dead code, :421-422);
  5. PATCH 2: ``torch.Tensor.cuda`` -> identity while running on CPU
     (ops/utils.py:198 hard-codes ``.cuda()``).
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import os
import sys
import tempfile
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("IMP_REFERENCE_ROOT", "/root/reference")
GROUP_SIZES = [82, 330, 513, 440, 1538, 451]          # umeml_gan.py:274
SIG_COLUMNS = ["Tumor Suppressor Genes", "Oncogenes", "Protein Kinases",
               "Cell Differentiation Markers", "Transcription Factors",
               "Cytokines and Growth Factors"]        # umeml_gan.py:350-355


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "medmm", "modeling", "ops", "utils.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    import torch
    import torch.nn as nn

    class DropPath(nn.Module):           # timm.models.layers.DropPath; p=0 in the reference
        def __init__(self, p: float = 0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    def _have(mod: str) -> bool:
        try:
            importlib.import_module(mod)
            return True
        except Exception:
            return False

    if not _have("timm"):
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", DropPath=DropPath)
    for name in ("ftfy", "h5py", "openpyxl", "yacs"):
        if not _have(name):
            _stub(name)
    if "yacs" in sys.modules and not hasattr(sys.modules["yacs"], "config"):
        class CfgNode(dict):
            pass
        sys.modules["yacs"].config = _stub("yacs.config", CfgNode=CfgNode)
    if not _have("matplotlib"):
        _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.pyplot")
        _stub("matplotlib.patches")
    if not _have("sksurv"):
        _stub("sksurv")
        _stub("sksurv.util", Surv=object)
        _stub("sksurv.metrics", concordance_index_censored=None,
              concordance_index_ipcw=None, integrated_brier_score=None,
              cumulative_dynamic_auc=None)

    def get_path_prorotypes():           # extract_prototype_with_plip_train.py:35-112 needs HF hub
        return torch.zeros(1, 6, 256)

    _stub("extract_prototype_with_plip_train", get_path_prorotypes=get_path_prorotypes)


_LOADED = {}


def load_ops():
    """Import ``medmm.modeling.ops`` from the reference tree (attention/blocks/utils)."""
    if "ops" in _LOADED:
        return _LOADED["ops"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # medmm/modeling/__init__.py pulls in every baseline; import the ops sub-package only,
    # by pre-seeding light-weight parent packages.
    for pkg in ("medmm", "medmm.modeling"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split("."))]
            m.__spec__ = importlib.machinery.ModuleSpec(pkg, loader=None, is_package=True)
            sys.modules[pkg] = m
    ops = importlib.import_module("medmm.modeling.ops")
    _LOADED["ops"] = ops
    return ops


def load_model_module():
    """Import ``medmm.modeling.models.umeml_gan`` (the module, not the registry factory)."""
    if "model" in _LOADED:
        return _LOADED["model"]
    load_ops()
    importlib.import_module("medmm.utils")
    importlib.import_module("medmm.modeling.models")
    mod = sys.modules["medmm.modeling.models.umeml_gan"]
    _LOADED["model"] = mod
    return mod


def load_distance():
    """``medmm.metrics.distance`` (euclidean_squared_distance, distance.py:46-61)."""
    load_ops()
    if "medmm.metrics" not in sys.modules:
        m = types.ModuleType("medmm.metrics")
        m.__path__ = [os.path.join(REFERENCE_ROOT, "medmm", "metrics")]
        m.__spec__ = importlib.machinery.ModuleSpec("medmm.metrics", loader=None, is_package=True)
        sys.modules["medmm.metrics"] = m
    return importlib.import_module("medmm.metrics.distance")


def make_cfg(n_proto: int = 6, dropout: float = 0.0, omic_dim: int = sum(GROUP_SIZES),
             fusion: str = "concat", n_reg: int = 3):
    """The keys UMEML_GAN.__init__ reads (umeml_gan.py:243-262)."""
    return SimpleNamespace(
        DATASET=SimpleNamespace(ROOT=".", PATH=SimpleNamespace(DIM=512),
                                OMIC=SimpleNamespace(DIM=omic_dim)),
        MODEL=SimpleNamespace(DROPOUT=dropout, HIDDEN_DIM=256, PROJECT_DIM=256,
                              FUSION=fusion, SIZE="small",
                              UMEML=SimpleNamespace(PROTOTYPES=n_proto, REGISTERS=n_reg)),
    )


def write_gene_fixtures(root: str) -> None:
    """signatures.csv + one molecular CSV such that group k matches GROUP_SIZES[k] rows,
    disjoint and contiguous."""
    import pandas as pd

    lab = os.path.join(root, "DATASET", "tcga_glioma", "labels")
    mol = os.path.join(root, "DATASET", "tcga_glioma", "molecular")
    os.makedirs(lab, exist_ok=True)
    os.makedirs(mol, exist_ok=True)
    cols = {}
    names = []
    longest = max(GROUP_SIZES)
    for k, (col, g) in enumerate(zip(SIG_COLUMNS, GROUP_SIZES)):
        genes = ["G%d_%04d" % (k, i) for i in range(g)]
        names += genes
        cols[col] = genes + [None] * (longest - g)
    pd.DataFrame(cols).to_csv(os.path.join(lab, "signatures.csv"), index=False)
    pd.DataFrame({"gene_name": names, "value": [0.0] * len(names)}).to_csv(
        os.path.join(mol, "TCGA-02-0047-01A-01-BS1.csv"), index=False)


@contextlib.contextmanager
def scratch_cwd():
    """Reference __init__/forward read and write relative paths (umeml_gan.py:348,365,576-587)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="imp_ref_") as d:
        write_gene_fixtures(d)
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


@contextlib.contextmanager
def cpu_cuda_noop():
    """PATCH 2: ops/utils.py:198 calls ``torch.eye(...).cuda()`` unconditionally."""
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build_reference_model(seed: int = 0, num_classes: int = 4, dropout: float = 0.0,
                          fusion: str = "concat"):
    """UMEML_GAN at the only configuration the shipped code can run (P = 6, D3/D7)."""
    import torch
    import torch.nn as nn

    mod = load_model_module()
    torch.manual_seed(seed)
    cfg = make_cfg(n_proto=6, dropout=dropout, fusion=fusion)
    with scratch_cwd():
        model = mod.UMEML_GAN(cfg, num_classes=num_classes, omic_sizes=1000)
    # PATCH 1 (D7): g_omic_net is Linear(1000,256) applied to all G genes; output unused.
    g = torch.Generator().manual_seed(seed + 12345)
    lin = nn.Linear(sum(GROUP_SIZES), 256)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.01)
        lin.bias.zero_()
    model.g_omic_net = nn.Sequential(lin, nn.ReLU(), nn.Dropout(dropout))
    model.cca = False                     # trainer sets it (trainer.py:620-626)
    model.float()
    return model


def run_reference_forward(model, batch, train: bool):
    """One forward under the CWD/.cuda() harness.  Returns whatever the model returns."""
    model.train(train)
    with scratch_cwd(), cpu_cuda_noop():
        return model(batch)
