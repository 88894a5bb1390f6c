"""N1 token tail: csrc/nystrom.cu (the Nystrom attention core, one CTA per slide and head) against the batched torch
form of token_tail.nystrom_short it replaces on the GPU -- same reduced matrices, same iteration
(medmm/modeling/ops/attention.py:105-127, ops/utils.py:116-131) -- in fp32 and against fp64."""
import pytest
import torch

from util_hotpath import record

pytestmark = pytest.mark.gpu


def _run(q, k, v, m, iters, use_kernel, monkeypatch):
    from imp_b200 import token_tail as T
    with monkeypatch.context() as mp:
        if not use_kernel:
            mp.setattr(T, "_core_supported", lambda *a: False)
        q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = T.nystrom_short(q, k, v, m, iters)
        g = torch.autograd.grad((out * torch.linspace(-1, 1, out.numel(), device=out.device, dtype=out.dtype).view_as(out)).sum(), [q, k, v])
    return out.detach(), [t.detach() for t in g]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("n,d,bsz", [(38, 32, 32), (7, 32, 4), (40, 32, 3), (47, 32, 2), (12, 64, 2), (1, 32, 2)])
def test_nystrom_core_matches_the_batched_form(n, d, bsz, monkeypatch):
    from imp_b200 import _lib
    torch.manual_seed(n * 131 + d)
    h, m, iters = 8, 128, 6
    q = torch.randn(bsz, h, n, d, device="cuda") * d ** -0.5
    k = torch.randn(bsz, h, n, d, device="cuda")
    v = torch.randn(bsz, h, n, d, device="cuda")
    before = _lib.launch_count()
    out_k, g_k = _run(q, k, v, m, iters, True, monkeypatch)
    assert _lib.launch_count() - before == 4              # matrix build + core, forward and backward
    out_t, g_t = _run(q, k, v, m, iters, False, monkeypatch)
    out_d, g_d = _run(q.double(), k.double(), v.double(), m, iters, False, monkeypatch)
    # the kernel is as close to fp64 as the fp32 library form is (both are fp32 chains of ~30 products), within 2x + 1e-6
    tag = "nystrom_core n=%d d=%d" % (n, d)
    record("test_nystrom_gpu", tag + " output", _rel(out_k, out_d), 1e-4, "batched torch form in fp64",
           "the fp32 library form it replaces: %.1e" % _rel(out_t, out_d))
    assert _rel(out_k, out_d) <= 2 * _rel(out_t, out_d) + 1e-6, (_rel(out_k, out_d), _rel(out_t, out_d))
    assert _rel(out_k, out_d) < 1e-4
    for a, b, c, name in zip(g_k, g_t, g_d, "qkv"):
        record("test_nystrom_gpu", tag + " grad " + name, _rel(a, c), 1e-3, "batched torch form in fp64",
               "the fp32 library form it replaces: %.1e" % _rel(b, c))
        assert _rel(a, c) <= 2 * _rel(b, c) + 1e-5, (name, _rel(a, c), _rel(b, c))
        assert _rel(a, c) < 1e-3, (name, _rel(a, c))


def test_nystrom_layer_gradients_reach_the_parameters(monkeypatch):
    """The whole NystromAttention layer (to_qkv, residual conv, to_out) with the kernel inside, under bf16 autocast as
    TRAINER.PREC = bf16 runs it: finite, and parameter gradients equal to the batched fp32 form within bf16 tolerance."""
    from imp_b200 import token_tail as T
    torch.manual_seed(3)
    att = T.NystromAttention(dim=256, dim_head=32, heads=8, num_landmarks=128, pinv_iterations=6, residual=True).cuda()
    x = torch.randn(5, 38, 256, device="cuda")
    att(x).square().sum().backward()
    g_kernel = {k: p.grad.clone() for k, p in att.named_parameters()}
    att.zero_grad()
    monkeypatch.setattr(T, "_core_supported", lambda *a: False)
    att(x).square().sum().backward()
    for k, p in att.named_parameters():
        assert _rel(g_kernel[k], p.grad) < 1e-3, (k, _rel(g_kernel[k], p.grad))
    monkeypatch.undo()
    att.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = att(x)
    y.float().square().sum().backward()
    for k, p in att.named_parameters():
        assert torch.isfinite(p.grad).all() and _rel(p.grad, g_kernel[k]) < 5e-2, (k, _rel(p.grad, g_kernel[k]))


def test_nystrom_core_rejects_unsupported_shapes():
    from imp_b200 import _lib, token_tail as T
    assert not T._core_supported(49, 32, 6) and not T._core_supported(20, 48, 6)
    mat = torch.eye(50, device="cuda").repeat(2, 1, 1)
    s = torch.ones(1, device="cuda")
    v = torch.zeros(2, 49, 32, device="cuda")
    with pytest.raises(_lib.ImpError):
        _lib.call("imp_nystrom_core_fwd", mat, s, v, None, 1, 1, 2, 50, 32, 6, torch.empty_like(v), None, _lib.stream_ptr())


def test_nystrom_backward_with_and_without_kept_iterates():
    """imp_nystrom_core_bwd with the iterates the forward kept and with saved = NULL (iteration repeated): same result."""
    from imp_b200 import _lib
    torch.manual_seed(5)
    n_mat, n_dim, d, iters = 16, 39, 32, 6
    mat = torch.softmax(torch.randn(n_mat, n_dim, n_dim, device="cuda"), -1)
    s = (1.0 / (mat.abs().sum(-1).max() * mat.abs().sum(-2).max())).reshape(1)
    v, dy = torch.randn(n_mat, n_dim - 1, d, device="cuda"), torch.randn(n_mat, n_dim - 1, d, device="cuda")
    y = torch.empty_like(v)
    saved = torch.empty(n_mat, _lib.query("imp_nystrom_core_saved_floats", n_dim, iters), device="cuda")
    w = torch.randn(8, 33, device="cuda") * 0.1                       # residual convolution: 8 heads, 33 taps
    _lib.call("imp_nystrom_core_fwd", mat, s, v, w, 8, 33, n_mat, n_dim, d, iters, y, saved, _lib.stream_ptr())
    res = []
    for buf in (saved, None):
        dmat, dv, ds = torch.empty_like(mat), torch.empty_like(v), torch.empty(n_mat, device="cuda")
        dw = torch.empty(n_mat, 33, device="cuda")
        _lib.call("imp_nystrom_core_bwd", mat, s, v, dy, buf, w, 8, 33, n_mat, n_dim, d, iters, dmat, ds, dv, dw, _lib.stream_ptr())
        res.append((dmat, dv, ds, dw))
    torch.cuda.synchronize()
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n", [5, 20, 38])
def test_fused_residual_convolution(n, monkeypatch):
    """nystrom_short with the depth-wise residual convolution inside the kernel against the batched form + F.conv2d:
    outputs, and gradients into q, k, v and the convolution weight (n < 17, = and > the half-width of the 33 taps)."""
    from imp_b200 import token_tail as T
    torch.manual_seed(n)
    b, h, d, m = 4, 8, 32, 128
    q = torch.randn(b, h, n, d, device="cuda", dtype=torch.float64) * d ** -0.5
    k, v = torch.randn(b, h, n, d, device="cuda", dtype=torch.float64), torch.randn(b, h, n, d, device="cuda", dtype=torch.float64)
    w = torch.randn(h, 1, 33, 1, device="cuda", dtype=torch.float64) * 0.2
    res = []
    for use_kernel, dt in ((True, torch.float32), (False, torch.float64)):
        with monkeypatch.context() as mp:
            if not use_kernel:
                mp.setattr(T, "_core_supported", lambda *a: False)
            args = [t.to(dt).clone().requires_grad_(True) for t in (q, k, v, w)]
            out = T.nystrom_short(args[0], args[1], args[2], m, 6, args[3])
            g = torch.autograd.grad((out * torch.linspace(-1, 1, out.numel(), device="cuda", dtype=dt).view_as(out)).sum(), args)
            res.append((out.detach(), g))
    assert _rel(res[0][0], res[1][0]) < 1e-4
    for a, b_, name in zip(res[0][1], res[1][1], "qkvw"):
        assert _rel(a, b_) < 1e-3, (name, _rel(a, b_))
