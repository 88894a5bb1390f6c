"""A1 path_net on tcgen05 vs a torch fp32 reference of the same op (bf16-rounded operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_fwd(x, w, b):
    return torch.relu(x.float() @ w.float().t() + b)


@pytest.mark.parametrize("rows", [128, 1, 77, 1000, 4096 + 33, 148 * 128 * 5 + 57])
def test_pathnet_fwd(rows):
    from imp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, 512, device="cuda", generator=g).bfloat16()
    w = (torch.randn(256, 512, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    guard = torch.full((rows + 128, 256), 7.0, device="cuda", dtype=torch.bfloat16)   # rows past the end stay untouched
    h = guard[:rows]
    _lib.call("imp_pathnet_fwd", x, w, b, h, rows, 512, 0.0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert bool((guard[rows:] == 7.0).all())
    ref = _ref_fwd(x, w, b)
    err = (h.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, err      # bf16 output rounding
    # tighter: relative Frobenius error
    assert ((h.float() - ref).norm() / ref.norm()).item() < 5e-3


@pytest.mark.parametrize("kdim", [256, 1024])
def test_pathnet_fwd_streaming_variant(kdim):
    """kdim != 512 takes the kernel that streams W1 per tile (the W1-stationary one holds exactly 128 x 512)."""
    from imp_b200 import _lib
    rows = 148 * 128 * 3 + 5
    g = torch.Generator(device="cuda").manual_seed(kdim)
    x = torch.randn(rows, kdim, device="cuda", generator=g).bfloat16()
    w = (torch.randn(256, kdim, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    h = torch.empty(rows, 256, device="cuda", dtype=torch.bfloat16)
    _lib.call("imp_pathnet_fwd", x, w, b, h, rows, kdim, 0.0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = _ref_fwd(x, w, b)
    assert ((h.float() - ref).norm() / ref.norm()).item() < 5e-3


def test_pathnet_fwd_dropout_deterministic():
    from imp_b200 import _lib
    rows = 777
    x = torch.randn(rows, 512, device="cuda").bfloat16()
    w = (torch.randn(256, 512, device="cuda") * 0.05).bfloat16()
    b = torch.zeros(256, device="cuda")
    outs = []
    for _ in range(2):
        h = torch.empty(rows, 256, device="cuda", dtype=torch.bfloat16)
        _lib.call("imp_pathnet_fwd", x, w, b, h, rows, 512, 0.25, 1234, _lib.stream_ptr())
        outs.append(h)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = _ref_fwd(x, w, b)
    pos = ref > 1e-3
    kept = (outs[0].float() != 0) & pos
    frac = kept.sum().item() / pos.sum().item()
    assert abs(frac - 0.75) < 0.01, frac
    ratio = (outs[0].float()[kept] / ref[kept]).mean().item()
    assert abs(ratio - 1 / 0.75) < 0.01, ratio


@pytest.mark.parametrize("rows", [64, 200, 4096, 16384 + 5])
def test_pathnet_dw(rows):
    from imp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, 512, device="cuda", generator=g).bfloat16()
    dz = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
    ws = torch.empty(_lib.query("imp_pathnet_dw_workspace_bytes", 512), device="cuda", dtype=torch.uint8)
    dw = torch.full((256, 512), 7.0, device="cuda")
    _lib.call("imp_pathnet_dw", dz, x, dw, ws, rows, 512, 0, _lib.stream_ptr())
    ref = dz.float().t() @ x.float()
    torch.cuda.synchronize()
    assert ((dw - ref).norm() / ref.norm()).item() < 1e-4
    _lib.call("imp_pathnet_dw", dz, x, dw, ws, rows, 512, 1, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert ((dw - 2 * ref).norm() / ref.norm()).item() < 2e-4
