"""Token-level tail of ``UMEML_GAN``: everything that works on the <= 40 tokens per slide
after the prototype-fusion hot path.  Batched torch, fixed shapes, no host synchronisation.

Reference (medmm/modeling/...), restated -- module and parameter names are the ``state_dict`` contract and are
kept; the arithmetic is re-derived in batched form:
  * ``NystromAttention``            ops/attention.py:46-161 (+ ``moore_penrose_iter_pinv`` ops/utils.py:116-131)
  * ``TransLayer``                  ops/blocks.py:252-268
  * ``Block``                       models/umeml_gan.py:86-96
  * ``BottleneckAttentionBlock``    models/umeml_gan.py:100-229 -- the reference fills a 7 x 7 similarity matrix with
    49 ``.item()`` round trips per slide (:126-129), sorts it on the host and pairs greedily (:174-186), then
    rebuilds the token list in Python (:188-220), twice per forward.  Here the greedy pairing is three masked
    arg-max steps on the device for the whole batch and the re-ordering is two gathers.
  * ``Generator`` / ``Discriminator``   models/umeml_gan.py:25-62
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def iterative_pinv(a: torch.Tensor, iters: int = 6) -> torch.Tensor:
    """Newton-Schulz style pseudo-inverse used by the Nystrom layers (ops/utils.py:116-131).  The initial scale is
    1 / (max column abs-sum x max row abs-sum) taken over the WHOLE tensor (all slides and heads together), as in
    the reference (``torch.max(col) * torch.max(row)`` :121)."""
    absa = a.abs()
    z = a.transpose(-1, -2) / (absa.sum(dim=-1).max() * absa.sum(dim=-2).max())
    eye = torch.eye(a.shape[-1], device=a.device, dtype=a.dtype)
    for _ in range(iters):
        az = a @ z
        z = 0.25 * z @ (13 * eye - az @ (15 * eye - az @ (7 * eye - az)))
    return z


# ------------------------------------------------------------------------------------------
# Nystrom attention when the sequence is shorter than the landmark count (always the case here: 7 - 40 tokens
# against 128 landmarks).  The reference pads the sequence IN FRONT with p = m - n zero tokens (attention.py:79-81);
# with one token per landmark the three similarity matrices coincide, A = softmax(q k^T) (m x m), and because to_qkv
# has no bias the padded tokens have q = k = v = 0.  A then has the block form
#       [ (1/m) J    (1/m) 1 1_n^T ]      J = all-ones (p x p), 1 = ones(p), c_i = 1 / Z_i, D = exp(s) / Z (n x n)
#       [  c 1^T          D        ]
# and every matrix the layer forms from it (products, transposes, a I - X) maps the (n+1)-dimensional subspace spanned
# by u = 1/sqrt(p) and the real-token coordinates into itself, acting there as
#       M(X) = [ alpha p + beta   sqrt(p) b^T ]        M(X Y) = M(X) M(Y),  M(X^T) = M(X)^T,  M(a I - X) = a I - M(X)
#              [ sqrt(p) c            D       ]
# (on the orthogonal complement X is a multiple of the identity).  The values [0; v] and therefore the whole output
# A pinv(A) A [0; v] live in that subspace, so the pseudo-inverse iteration (ops/utils.py:116-131) can run on the
# (n+1) x (n+1) matrices M: the same ~30 batched products as the literal form, on 8 x 8 ... 41 x 41 instead of 128 x 128
# matrices (26 GFLOP per layer call at 32 slides -- 26 ms of a 61 ms whole-model step -- become ~0.3 GFLOP), with
# results identical up to fp32 round-off.
# ------------------------------------------------------------------------------------------
def nystrom_short(q, k, v, m: int, iters: int, conv_w=None):
    """q (scaled), k, v (B,H,n,d) with n < m -> the last n rows of A pinv(A) A [0; v]  (B,H,n,d), plus the depth-wise
    residual convolution of v over the tokens when ``conv_w`` (H,1,taps,1) is given (attention.py:129-131; the zero
    tokens the reference pads in front act like the convolution's own zero padding)."""
    n = q.shape[-2]
    p = float(m - n)
    rp = math.sqrt(p)
    if q.is_cuda and _core_supported(n + 1, v.shape[-1], iters) and q.shape[-1] <= 128:
        # two kernels per direction (csrc/nystrom.cu): the reduced matrix with its row / column maxima, then the core.
        # Only the batch-global scale stays in torch, so that autograd sends its gradient to the arg-max row / column.
        mat, rowmax, colmax = _NystromBuild.apply(q, k, int(m))
        rs = rowmax.max().clamp_min((p + n) / m)                 # padded rows sum to (p + n) / m = 1
        inv_scale = 1.0 / (rs * colmax.max())
        return _NystromCore.apply(mat, inv_scale, v, iters, None if conv_w is None else conv_w.reshape(conv_w.shape[0], -1))
    s = q @ k.transpose(-1, -2)                                  # (B,H,n,n) real-token block of q k^T
    smax = s.max(dim=-1, keepdim=True).values.clamp_min(0.0)     # padded columns have logit 0
    es = torch.exp(s - smax)
    e0 = torch.exp(-smax)                                        # (B,H,n,1)
    zsum = p * e0 + es.sum(-1, keepdim=True)
    cvec, dmat = e0 / zsum, es / zsum
    # initial scale 1 / (max row-abs-sum * max column-abs-sum) of the FULL m x m matrix, over the whole batch (utils.py:119-121)
    rs = (p * cvec + dmat.sum(-1, keepdim=True)).max().clamp_min((p + n) / m)      # padded rows sum to (p + n) / m = 1
    cs = torch.maximum((p / m + cvec.sum(-2)).max(), (p / m + dmat.sum(-2)).max())
    top = torch.cat([s.new_full(s.shape[:-2] + (1, 1), p / m), s.new_full(s.shape[:-2] + (1, n), rp / m)], dim=-1)
    mat = torch.cat([top, torch.cat([rp * cvec, dmat], dim=-1)], dim=-2)          # M(A), (B,H,n+1,n+1)
    eye = torch.eye(n + 1, device=s.device, dtype=s.dtype)
    z = mat.transpose(-1, -2) / (rs * cs)
    for _ in range(iters):
        az = mat @ z
        z = 0.25 * z @ (13 * eye - az @ (15 * eye - az @ (7 * eye - az)))
    v1 = F.pad(v, (0, 0, 1, 0))                                  # coordinates of [0; v]: zero along u
    out = (mat @ (z @ (mat @ v1)))[..., 1:, :]
    if conv_w is not None:
        out = out + F.conv2d(v, conv_w, padding=(conv_w.shape[2] // 2, 0), groups=conv_w.shape[0])
    return out


def _core_supported(n_dim: int, head_dim: int, iters: int) -> bool:
    """Shapes csrc/nystrom.cu holds in shared memory; anything else (more than 47 tokens: P = 64 prototypes) keeps
    the batched form above."""
    return n_dim <= 48 and head_dim in (32, 64) and 0 <= iters <= 8


class _NystromBuild(torch.autograd.Function):
    """(q scaled, k) (B,H,n,d) -> the reduced matrix M (B,H,n+1,n+1) and, per matrix, the largest row and column sum of
    the full m x m soft-max matrix (differentiable: their gradients go to the arg-max row / column)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, m):
        from . import _lib
        q, k = q.contiguous(), k.contiguous()
        b, h, n, d = q.shape
        mat = torch.empty(b, h, n + 1, n + 1, device=q.device, dtype=torch.float32)
        rowmax = torch.empty(b, h, device=q.device, dtype=torch.float32)
        colmax = torch.empty(b, h, device=q.device, dtype=torch.float32)
        _lib.call("imp_nystrom_build_fwd", q, k, b * h, n, d, int(m), mat, rowmax, colmax, _lib.stream_ptr())
        ctx.save_for_backward(q, k)
        ctx.m = int(m)
        return mat, rowmax, colmax

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dmat, drow, dcol):
        from . import _lib
        q, k = ctx.saved_tensors
        b, h, n, d = q.shape
        dmat = torch.zeros(b, h, n + 1, n + 1, device=q.device) if dmat is None else dmat.contiguous().float()
        drow = torch.zeros(b, h, device=q.device) if drow is None else drow.contiguous().float()
        dcol = torch.zeros(b, h, device=q.device) if dcol is None else dcol.contiguous().float()
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        _lib.call("imp_nystrom_build_bwd", q, k, dmat, drow, dcol, b * h, n, d, ctx.m, dq, dk, _lib.stream_ptr())
        return dq, dk, None


class _NystromCore(torch.autograd.Function):
    """y = rows 1.. of M (pinv_iter(M) (M [0; v])) with the iteration's initial scale as a tensor argument, so that the
    gradient the reference sends through ``torch.max`` of the row / column sums (ops/utils.py:119-121) is kept."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, mat, inv_scale, v, iters, conv_w=None):
        from . import _lib
        mat, v = mat.contiguous(), v.contiguous()
        heads, taps = (conv_w.shape[0], conv_w.shape[1]) if conv_w is not None else (1, 1)
        if conv_w is not None:
            conv_w = conv_w.contiguous()
            if v.dim() != 4 or v.shape[1] != heads:
                raise ValueError("v must be (slides, heads, tokens, head_dim) with %d heads" % heads)
        inv_scale = inv_scale.reshape(1).contiguous()
        n_dim, d = mat.shape[-1], v.shape[-1]
        n_mat = mat.numel() // (n_dim * n_dim)
        y = torch.empty_like(v)
        saved = None
        if any(ctx.needs_input_grad):                         # the iterates Z_k, so that the backward does not repeat the iteration
            saved = torch.empty(n_mat, _lib.query("imp_nystrom_core_saved_floats", n_dim, int(iters)), device=mat.device,
                                dtype=torch.float32)
        _lib.call("imp_nystrom_core_fwd", mat, inv_scale, v, conv_w, heads, taps, n_mat, n_dim, d, int(iters), y, saved,
                  _lib.stream_ptr())
        ctx.save_for_backward(mat, inv_scale, v, saved, conv_w)
        ctx.iters = int(iters)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        from . import _lib
        mat, inv_scale, v, saved, conv_w = ctx.saved_tensors
        n_dim, d = mat.shape[-1], v.shape[-1]
        n_mat = mat.numel() // (n_dim * n_dim)
        dy = dy.contiguous().float()
        dmat, dv = torch.empty_like(mat), torch.empty_like(v)
        dscale = torch.empty(n_mat, device=mat.device, dtype=torch.float32)
        heads, taps = (conv_w.shape[0], conv_w.shape[1]) if conv_w is not None else (1, 1)
        dconv = torch.empty(n_mat, taps, device=mat.device, dtype=torch.float32) if conv_w is not None else None
        _lib.call("imp_nystrom_core_bwd", mat, inv_scale, v, dy, saved, conv_w, heads, taps, n_mat, n_dim, d, ctx.iters, dmat,
                  dscale, dv, dconv, _lib.stream_ptr())
        return (dmat, dscale.sum().reshape(()), dv, None,
                None if conv_w is None else dconv.view(-1, heads, taps).sum(0))


class NystromAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6, residual=True,
                 residual_conv_kernel=33, eps=1e-8, dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        self.eps, self.heads, self.num_landmarks, self.pinv_iterations = eps, heads, num_landmarks, pinv_iterations
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            self.res_conv = nn.Conv2d(heads, heads, (residual_conv_kernel, 1), padding=(residual_conv_kernel // 2, 0),
                                      groups=heads, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, n, _ = x.shape
        h, m = self.heads, self.num_landmarks
        if n < m:                                               # every call of this model: reduced block algebra
            qkv = self.to_qkv(x).view(b, n, 3, h, -1).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
            out = nystrom_short(q, k, v, m, self.pinv_iterations, self.res_conv.weight if self.residual else None)
            return self.to_out(out.transpose(1, 2).reshape(b, n, -1))
        return self.forward_dense(x)

    def forward_dense(self, x: torch.Tensor) -> torch.Tensor:
        """The literal form (any n): pad, landmarks by group means, three softmaxes, dense pseudo-inverse."""
        b, n, _ = x.shape
        h, m = self.heads, self.num_landmarks
        pad = (m - n % m) % m                                   # zero tokens in FRONT (attention.py:79-81)
        if pad:
            x = F.pad(x, (0, 0, pad, 0))
        npad = n + pad
        qkv = self.to_qkv(x).view(b, npad, 3, h, -1).permute(2, 0, 3, 1, 4)     # (3, b, h, npad, d)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        grp = math.ceil(n / m)                                  # tokens per landmark (:97-101)
        ql = q.reshape(b, h, npad // grp, grp, -1).sum(dim=3) / grp
        kl = k.reshape(b, h, npad // grp, grp, -1).sum(dim=3) / grp
        a1 = torch.softmax(q @ kl.transpose(-1, -2), dim=-1)
        a2 = torch.softmax(ql @ kl.transpose(-1, -2), dim=-1)
        a3 = torch.softmax(ql @ k.transpose(-1, -2), dim=-1)
        out = (a1 @ iterative_pinv(a2, self.pinv_iterations)) @ (a3 @ v)
        if self.residual:
            out = out + self.res_conv(v)
        out = self.to_out(out.transpose(1, 2).reshape(b, npad, -1))
        return out[:, -n:]


class TransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                     residual=True, dropout=0.1)

    def forward(self, x):
        return x + self.attn(self.norm(x))


class Block(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.attn = TransLayer(dim=dim)

    def forward(self, x):
        return self.attn(x)


def greedy_pairs(sim: torch.Tensor, k: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """sim (B, Lp, Lo) -> (i_p (B,k), i_o (B,k)): the k best pairs with distinct rows and columns, best first.
    Equivalent to the reference's scan of the descending-sorted entries that skips used rows/columns
    (umeml_gan.py:174-186): the next accepted entry is always the arg-max over the rows and columns still free."""
    b, lp, lo = sim.shape
    work = sim.clone()
    ips, ios = [], []
    neg = torch.finfo(sim.dtype).min
    for _ in range(k):
        flat = work.reshape(b, -1).argmax(dim=1)
        ip, io = flat // lo, flat % lo
        ips.append(ip); ios.append(io)
        work = work.masked_fill(F.one_hot(ip, lp).bool()[:, :, None] | F.one_hot(io, lo).bool()[:, None, :], neg)
    return torch.stack(ips, dim=1), torch.stack(ios, dim=1)


def _remaining(tokens: torch.Tensor, picked: torch.Tensor) -> torch.Tensor:
    """tokens (B,L,D) without the rows ``picked`` (B,k), original order kept (umeml_gan.py:200-217)."""
    b, l, d = tokens.shape
    taken = F.one_hot(picked, l).sum(dim=1)                                      # (B,L) 0/1
    order = torch.argsort(taken * l + torch.arange(l, device=tokens.device)[None, :], dim=1)
    keep = order[:, : l - picked.shape[1]]
    return torch.gather(tokens, 1, keep[:, :, None].expand(-1, -1, d))


class BottleneckAttentionBlock(nn.Module):
    def __init__(self, dim: int = 256, n_reg: int = 2):
        super().__init__()
        self.bottle_tokens = nn.Parameter(torch.empty(1, n_reg, dim).uniform_())
        self.encoders = nn.ModuleList([Block(dim=dim) for _ in range(2)])
        self.linear_p = nn.Linear(dim, dim)          # the reference attaches these from UMEML_GAN.__init__ (:307-308)
        self.linear_o = nn.Linear(dim, dim)

    def forward(self, x_path: torch.Tensor, x_omic: torch.Tensor, patient_id=None):
        b, path_len, d = x_path.shape
        token_len = self.bottle_tokens.shape[1]
        pn = x_path / x_path.norm(dim=-1, keepdim=True).clamp_min(1e-8)          # F.cosine_similarity, eps 1e-8 (:129)
        on = x_omic / x_omic.norm(dim=-1, keepdim=True).clamp_min(1e-8)
        ip, io = greedy_pairs((pn @ on.transpose(1, 2)).detach(), 3)
        sel_p = torch.gather(x_path, 1, ip[:, :, None].expand(-1, -1, d))
        sel_o = torch.gather(x_omic, 1, io[:, :, None].expand(-1, -1, d))
        ks = self.linear_p(sel_p) + self.linear_o(sel_o)                         # :194, in pairing order
        x = torch.cat([ks, _remaining(x_path, ip), self.bottle_tokens.expand(b, -1, -1), _remaining(x_omic, io)], dim=1)
        for blk in self.encoders:
            x = blk(x)
        # the reference slices with the ORIGINAL lengths although three tokens per modality were merged (:227-228)
        t_path, f_path = x[:, :1], x[:, 1:path_len]
        t_omic, f_omic = x[:, path_len + token_len: path_len + token_len + 1], x[:, path_len + token_len + 1:]
        return t_path, f_path, t_omic, f_omic


class Generator(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = tuple(input_dim), tuple(output_dim)
        self.net = nn.Sequential(nn.Linear(input_dim[0] * input_dim[1], 1024), nn.ReLU(),
                                 nn.Linear(1024, output_dim[0] * output_dim[1]), nn.Softplus())

    def forward(self, x):
        if tuple(x.shape[-2:]) != self.input_dim:
            raise ValueError(f"Expected input shape (-2 dimensions): {self.input_dim}, but got: {tuple(x.shape[-2:])}")
        return self.net(x.reshape(x.shape[0], -1)).view(x.shape[0], *self.output_dim)


class Discriminator(nn.Module):
    def __init__(self, input_shape):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(input_shape[0] * input_shape[1], 256), nn.ReLU(), nn.Linear(256, 1), nn.Sigmoid())

    def forward(self, x):
        return self.layers(x.reshape(x.shape[0], -1))


def transform_importance(x: torch.Tensor) -> torch.Tensor:
    """per-sample min-max to [0.5, 1] (umeml_gan.py:689-694)"""
    lo, hi = x.min(dim=1, keepdim=True)[0], x.max(dim=1, keepdim=True)[0]
    return 0.5 + 0.5 * (x - lo) / (hi - lo + 1e-8)


def transform_importance_to_half_one_point_five(x: torch.Tensor) -> torch.Tensor:
    """per-sample min-max to [0.5, 1.5] (umeml_gan.py:696-702)"""
    lo, hi = x.min(dim=1, keepdim=True)[0], x.max(dim=1, keepdim=True)[0]
    return 0.5 + (x - lo) / (hi - lo + 1e-8)
