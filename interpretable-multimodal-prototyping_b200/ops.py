"""Host-side mirror of the reference operator surface for the hot path.

Mirrors (same names, argument meaning and layouts):
  * ``MultiheadAttention``        medmm/modeling/ops/blocks.py:346-526  (single head, key is value)
  * ``PathProtoGenerator``        medmm/modeling/models/umeml_gan.py:65-80
  * ``compute_modularity``        medmm/modeling/ops/utils.py:205-228   (see modularity.py)
and adds the fused entry the model uses: ``proto_fusion`` = path_net + the stacked prototype
blocks over a packed varlen batch of bags, with one recompute-free backward pass over h.

All N-scaling work runs in the CUDA kernels behind ``kernels``; the P x 256 token algebra
(q/k fold, out-proj, LayerNorm) is a handful of tiny cuBLAS calls differentiated by autograd.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels, parallel

D = kernels.D
SENTINEL = -10000.0          # data/data_manager.py:387, umeml_gan.py:404


# ------------------------------------------------------------------------------------------
# token algebra (exact refactor of attention.py:355-533 for heads = 1, key = value)
# ------------------------------------------------------------------------------------------
def fold_query(c: torch.Tensor, in_w: torch.Tensor, in_b: Optional[torch.Tensor]) -> torch.Tensor:
    """q~ = ((c Wq^T + bq) * D^-1/2) Wk.  S_pn = q~_p . h_n + q_p . b_k; the second term is
    constant over patches and cancels in the softmax (attention.py:368,382,432,509,527)."""
    d = c.shape[-1]
    q = F.linear(c, in_w[:d], None if in_b is None else in_b[:d]) * (float(d) ** -0.5)
    return q @ in_w[d:2 * d]


def block_tail(c_in: torch.Tensor, pooled: torch.Tensor, in_w, in_b, out_w, out_b, ln_w, ln_b) -> torch.Tensor:
    """c + LayerNorm(((a h) Wv^T + bv) Wo^T + bo)   (attention.py:530-533, umeml_gan.py:79).
    sum_n a_pn = 1, so the value bias passes through the pooling unchanged."""
    d = c_in.shape[-1]
    v = F.linear(pooled, in_w[2 * d:], None if in_b is None else in_b[2 * d:])
    o = F.linear(v, out_w, out_b)
    return c_in + F.layer_norm(o, (d,), ln_w, ln_b, 1e-5)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


def _to_bf16(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.bfloat16:
        return t.contiguous()
    return kernels.cast_bf16(t.contiguous().float())


def _cu_from_lengths(lengths: Sequence[int], device) -> torch.Tensor:
    cu = [0]
    for n in lengths:
        cu.append(cu[-1] + int(n))
    return torch.tensor(cu, dtype=torch.int32, device=device)


# ------------------------------------------------------------------------------------------
# generic pooling op:  pooled = softmax_n(qt h^T) h   with gradients to qt and h
# ------------------------------------------------------------------------------------------
class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, cu, max_len, qt):
        pooled, lse = kernels.pool_fwd(h, cu, max_len, qt.contiguous())
        ctx.save_for_backward(h, cu, qt, pooled, lse)
        ctx.max_len = max_len
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        h, cu, qt, pooled, lse = ctx.saved_tensors
        dpooled = dpooled.contiguous()
        delta = (_bf16_round(dpooled) * pooled).sum(-1).contiguous()
        dq, dh = kernels.pool_bwd(h, cu, ctx.max_len, [qt.contiguous()], [dpooled], [lse], [delta], 0,
                                  want_dz=ctx.needs_input_grad[0], relu_mask=False)
        if qt.shape[0] == 1 and dq.shape[0] != 1:
            dq = dq.sum(0, keepdim=True)
        return dh, None, None, dq.view_as(qt)


def softmax_pool(h: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, qt: torch.Tensor) -> torch.Tensor:
    """h (R,256) bf16 packed bags, qt (B|1,P,256) fp32 -> pooled (B,P,256) fp32."""
    return _PoolFn.apply(h, cu_seqlens, int(max_len), qt)


# ------------------------------------------------------------------------------------------
# fused hot path: path_net + stacked prototype blocks (umeml_gan.py:410,425-434)
# ------------------------------------------------------------------------------------------
def _block_forward(c, pooled_fn, in_w, in_b, out_w, out_b, ln_w, ln_b):
    """One prototype block on tokens c (B|1,P,D) with every intermediate the hand-derived backward needs.
    ``pooled_fn(q~) -> (pooled, lse)`` is the pooling kernel (plus the cross-rank merge in giant-bag mode)."""
    d = c.shape[-1]
    q = F.linear(c, in_w[:d], None if in_b is None else in_b[:d]) * (float(d) ** -0.5)      # attention.py:368,432
    qt = (q @ in_w[d:2 * d]).contiguous()                                                     # folded key projection
    pooled, lse = pooled_fn(qt)
    v = F.linear(pooled, in_w[2 * d:], None if in_b is None else in_b[2 * d:])               # :382 (value half), :530
    o = F.linear(v, out_w, out_b)                                                             # :533
    y, mean, rstd = torch.native_layer_norm(o, (d,), ln_w, ln_b, 1e-5)                        # umeml_gan.py:79
    return c + y, dict(c=c, q=q, qt=qt, pooled=pooled, lse=lse, v=v, o=o, mean=mean, rstd=rstd)


def _block_tail_backward(g_c, sv, in_w, out_w, ln_w, ln_b):
    """Cotangent g_c of the block output -> (dpooled, d c_in through the residual, parameter gradients of the tail)."""
    d = g_c.shape[-1]
    do, dln_w, dln_b = torch.ops.aten.native_layer_norm_backward(g_c, sv["o"], [d], sv["mean"], sv["rstd"], ln_w, ln_b, [True, True, True])
    do2, v2 = do.reshape(-1, d), sv["v"].reshape(-1, d)
    dv = do2 @ out_w
    grads = {"out_w": do2.t() @ v2, "out_b": do2.sum(0), "ln_w": dln_w, "ln_b": dln_b,
             "wv": dv.t() @ sv["pooled"].reshape(-1, d), "bv": dv.sum(0)}
    dpooled = (dv @ in_w[2 * d:]).view_as(sv["pooled"])
    return dpooled, grads


def _fold_backward(dqt, sv, in_w):
    """Cotangent of the folded queries q~ (same batch shape as c_in) -> (d c_in, dWq, dbq, dWk)."""
    d = dqt.shape[-1]
    scale = float(d) ** -0.5
    dqt2, q2, c2 = dqt.reshape(-1, d), sv["q"].reshape(-1, d), sv["c"].reshape(-1, d)
    dpre = (dqt2 @ in_w[d:2 * d].t()) * scale
    return (dpre @ in_w[:d]).view_as(sv["c"]), dpre.t() @ c2, dpre.sum(0), q2.t() @ dqt2


class _ProtoFusionFn(torch.autograd.Function):
    """inputs: x (R,512) bf16 packed, cu (B+1) int32, max_len, p_drop, seed, p_proto (1|B,P,256),
    w1, b1, then 6 tensors per block (in_proj_weight, in_proj_bias, out_proj.weight,
    out_proj.bias, norm1.weight, norm1.bias).  Returns (c (B,P,256) fp32, h (R,256) bf16).

    The P x 256 token algebra between the kernels (query fold, value / output projection, LayerNorm, residual) is
    evaluated with its intermediates kept, and differentiated by hand in ``backward``: about a dozen batched GEMMs per
    block instead of the ~100 small kernels per block an autograd graph over the same expressions launches."""

    @staticmethod
    def forward(ctx, x, cu, max_len, p_drop, seed, shard_group, p_proto, w1, b1, *blk):
        nblk = len(blk) // 6
        if nblk not in (1, 2):
            raise ValueError("proto_fusion supports 1 or 2 stacked blocks (reference: 2, umeml_gan.py:289)")
        nb = cu.numel() - 1
        h = kernels.pathnet_fwd(x, kernels.cast_bf16(w1.detach().contiguous()), b1.detach().contiguous(),
                                p_drop, seed)

        def pooled_fn(qt):
            pooled, lse = kernels.pool_fwd(h, cu, max_len, qt)
            if shard_group is not None:          # giant bag sharded over ranks: LSE merge of the partial states
                pooled, lse = parallel.merge_shards(pooled, lse, shard_group)
            return pooled, lse

        c = p_proto.detach()
        saved = []
        with torch.no_grad():
            for k in range(nblk):
                c, sv = _block_forward(c, pooled_fn, *blk[6 * k:6 * k + 6])
                saved.append(sv)
        if c.shape[0] != nb:
            c = c.expand(nb, -1, -1)
        ctx.save_for_backward(x, h, cu, w1, b1, *blk)
        ctx.block_saved = saved
        ctx.nblk, ctx.max_len, ctx.p_drop, ctx.shard_group = nblk, max_len, p_drop, shard_group
        ctx.mark_non_differentiable(h)
        return c.contiguous(), h

    @staticmethod
    def backward(ctx, dc, _dh_unused):
        nblk = ctx.nblk
        t = ctx.saved_tensors
        x, h, cu, w1, b1 = t[:5]
        blk = t[5:5 + 6 * nblk]
        saved = ctx.block_saved
        keep_scale = 1.0
        if ctx.p_drop > 0:   # the kernel keeps an element iff its 16 hash bits >= round(65536 p): match its scale
            thr = int(ctx.p_drop * 65536.0 + 0.5)
            keep_scale = 65536.0 / (65536.0 - thr) if thr else 1.0
        d = D
        g_c = dc.contiguous()                 # cotangent of the current block's output, (B,P,D)
        qts = [sv["qt"] for sv in saved]
        lses = [sv["lse"] for sv in saved]
        dpool: List[Optional[torch.Tensor]] = [None] * nblk
        delta: List[Optional[torch.Tensor]] = [None] * nblk
        par: List[Optional[dict]] = [None] * nblk
        db1 = torch.empty(D, device=x.device, dtype=torch.float32)
        dz = None
        d_proto = None
        for k in reversed(range(nblk)):
            in_w, in_b, out_w, out_b, ln_w, ln_b = blk[6 * k:6 * k + 6]
            sv = saved[k]
            if g_c.shape[0] != sv["pooled"].shape[0]:
                g_c = g_c.expand(sv["pooled"].shape[0], -1, -1)
            dp, grads = _block_tail_backward(g_c.contiguous(), sv, in_w, out_w, ln_w, ln_b)
            dpool[k] = dp.contiguous()
            delta[k] = (_bf16_round(dpool[k]) * sv["pooled"]).sum(-1).contiguous()
            if k > 0:
                dq, _ = kernels.pool_bwd(h, cu, ctx.max_len, [qts[k]], [dpool[k]], [lses[k]], [delta[k]], 0, want_dz=False)
            else:
                dq, dz = kernels.pool_bwd(h, cu, ctx.max_len, qts, dpool, lses, delta, 0, want_dz=True,
                                          relu_mask=True, keep_scale=keep_scale, db1=db1)
            if ctx.shard_group is not None:       # every rank saw only its patches: sum the partial dq~
                parallel.allreduce_sum_(dq, ctx.shard_group)
            shared = sv["c"].shape[0] == 1 and dq.shape[0] != 1
            if shared:
                dq = dq.sum(0, keepdim=True)
            dc_in, dwq, dbq, dwk = _fold_backward(dq, sv, in_w)
            grads.update(wq=dwq, bq=dbq, wk=dwk)
            par[k] = grads
            # cotangent of this block's input tokens: residual path + query path
            g_prev = g_c.sum(0, keepdim=True) if shared else g_c
            g_c = g_prev + dc_in
        d_proto = g_c
        dw1 = kernels.pathnet_dw(dz, x)
        if ctx.shard_group is not None:
            parallel.allreduce_sum_(dw1, ctx.shard_group)
            parallel.allreduce_sum_(db1, ctx.shard_group)
        need = ctx.needs_input_grad
        res = [None, None, None, None, None, None,
               d_proto if need[6] else None,
               dw1.to(w1.dtype) if need[7] else None,
               db1.to(b1.dtype) if need[8] else None]
        for k in range(nblk):
            g = par[k]
            in_b = blk[6 * k + 1]
            res += [torch.cat([g["wq"], g["wk"], g["wv"]], dim=0) if need[9 + 6 * k] else None,
                    (torch.cat([g["bq"], torch.zeros_like(g["bq"]), g["bv"]]) if (in_b is not None and need[10 + 6 * k]) else None),
                    g["out_w"] if need[11 + 6 * k] else None,
                    g["out_b"] if need[12 + 6 * k] else None,
                    g["ln_w"] if need[13 + 6 * k] else None,
                    g["ln_b"] if need[14 + 6 * k] else None]
        return tuple(res)


def block_params(blk: "PathProtoGenerator"):
    a = blk.cross_attn
    return (a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, blk.norm1.weight, blk.norm1.bias)


def proto_fusion(x_packed: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, p_proto: torch.Tensor,
                 w1: torch.Tensor, b1: torch.Tensor, blocks: Sequence[Sequence[torch.Tensor]],
                 p_drop: float = 0.0, seed: int = 0, shard_group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (prototype tokens (B,P,256) fp32, h (R,256) bf16 for the modularity term).
    ``shard_group``: a torch.distributed group whose ranks each hold a contiguous patch shard of the
    SAME bags (giant-bag mode): partial softmax states are merged across the group."""
    flat = [t for b in blocks for t in b]
    return _ProtoFusionFn.apply(x_packed, cu_seqlens, int(max_len), float(p_drop), int(seed), shard_group, p_proto,
                                w1, b1, *flat)


# ------------------------------------------------------------------------------------------
# A0: reference batch layout -> packed bf16 rows
# ------------------------------------------------------------------------------------------
def strip_and_pack(img: torch.Tensor, lengths: Optional[Sequence[int]] = None):
    """img (B,Npad,512) fp32 with -10000 row padding (data_manager.py:356-367) -> (x (R,512) bf16,
    cu_seqlens (B+1) int32 on device, max_len).  With ``lengths`` (host ints) the packed buffer is
    exactly sized; otherwise lengths come from the device sentinel scan, no host sync, and the
    buffer holds B*Npad rows of which the first cu[B] are meaningful."""
    b, npad, d = img.shape
    img = img.contiguous().float()
    if lengths is not None:
        cu = _cu_from_lengths(lengths, img.device)
        total = int(sum(int(n) for n in lengths))
        x = torch.empty(total, d, device=img.device, dtype=torch.bfloat16)
        max_len = max(int(n) for n in lengths) if len(lengths) else 0
    else:
        _, cu = kernels.bag_lengths(img, SENTINEL)
        x = torch.zeros(b * npad, d, device=img.device, dtype=torch.bfloat16)
        max_len = npad
    kernels.pack_bags(img, cu, x)
    return x, cu, max_len


# ------------------------------------------------------------------------------------------
# modules with the reference's parameter names
# ------------------------------------------------------------------------------------------
class MultiheadAttention(nn.Module):
    """Single-head cross attention with key is value (the only configuration on the hot path;
    blocks.py:346-526 / attention.py:236-547).  Layout (L,B,E) like the reference; returns
    (attn_output (L,B,E), raw pre-softmax logits (B,1,L,S) or None when need_raw=False)."""

    def __init__(self, embed_dim: int, num_heads: int = 1, dropout: float = 0.0, bias: bool = True):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError("the IMP hot path uses num_heads=1 (umeml_gan.py:72)")
        if dropout != 0.0:
            raise NotImplementedError("attention dropout is 0 on the hot path (blocks.py:378)")
        if embed_dim != D:
            raise NotImplementedError("kernels are built for embed_dim=256 (MODEL.HIDDEN_DIM)")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, 1, embed_dim
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim)) if bias else None
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self._reset_parameters()

    def _reset_parameters(self):           # blocks.py:418-428
        nn.init.xavier_uniform_(self.in_proj_weight)
        if self.in_proj_bias is not None:
            nn.init.constant_(self.in_proj_bias, 0.0)
            nn.init.constant_(self.out_proj.bias, 0.0)

    def pooled_tokens(self, c: torch.Tensor, h: torch.Tensor, cu: torch.Tensor, max_len: int) -> torch.Tensor:
        """c (B|1,P,E), packed h -> attention output before the residual/LayerNorm, (B,P,E)."""
        d = self.embed_dim
        qt = fold_query(c, self.in_proj_weight, self.in_proj_bias)
        pooled = softmax_pool(h, cu, max_len, qt)
        bias = self.in_proj_bias
        v = F.linear(pooled, self.in_proj_weight[2 * d:], None if bias is None else bias[2 * d:])
        return self.out_proj(v)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, need_raw=True, attn_mask=None):
        if key is not value and not torch.equal(key, value):
            raise NotImplementedError("hot path: key and value are the same patch tokens (umeml_gan.py:77)")
        if key_padding_mask is not None or attn_mask is not None:
            raise NotImplementedError("masks are not used on the hot path")
        L, B, E = query.shape
        S = key.shape[0]
        if key.shape[1] != B:
            raise ValueError("batch mismatch between query and key")
        hk = key.transpose(0, 1).reshape(B * S, E)                    # bags back to back
        h = _CastBf16.apply(hk) if hk.requires_grad else _to_bf16(hk)    # keep the graph through the cast
        cu = torch.arange(0, (B + 1) * S, S, dtype=torch.int32, device=query.device)
        out = self.pooled_tokens(query.transpose(0, 1), h, cu, S).transpose(0, 1)
        raw = None
        if need_raw:
            d = E
            q = F.linear(query.transpose(0, 1), self.in_proj_weight[:d], self.in_proj_bias[:d]) * (float(d) ** -0.5)
            k = F.linear(key.transpose(0, 1).float(), self.in_proj_weight[d:2 * d], self.in_proj_bias[d:2 * d])
            raw = torch.bmm(q, k.transpose(1, 2)).view(B, 1, L, S)   # attention.py:509,535-538
        return out, raw


class _CastBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        ctx.dtype = t.dtype
        return _to_bf16(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype)


class PathProtoGenerator(nn.Module):
    """c <- c + LayerNorm(CrossAttn(c, x, x))   (umeml_gan.py:65-80).  x (B,N,D), c (B|1,P,D)."""

    def __init__(self, dim: int, drop_path: float = 0.0):
        super().__init__()
        if drop_path > 0.0:
            raise NotImplementedError("drop_path is 0 in the reference model (umeml_gan.py:289)")
        self.cross_attn = MultiheadAttention(embed_dim=dim, num_heads=1)
        self.drop_path1 = nn.Identity()
        self.norm1 = nn.LayerNorm(dim)

    def forward(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        _c, _ = self.cross_attn(c.transpose(1, 0), x.transpose(1, 0), x.transpose(1, 0), need_raw=False)
        return c + self.drop_path1(self.norm1(_c.transpose(1, 0)))


def reset_prototypes(n_proto: int, dim: int = D, generator: Optional[torch.Generator] = None, device=None):
    """p_proto ~ U(-1/P, 1/P), shape (1,P,D): umeml_gan.py:23,310-315 (plain tensor, not a Parameter)."""
    t = torch.empty(1, n_proto, dim, device=device)
    return t.uniform_(-1.0 / n_proto, 1.0 / n_proto, generator=generator)


__all__ = ["MultiheadAttention", "PathProtoGenerator", "fold_query", "block_tail", "softmax_pool",
           "proto_fusion", "block_params", "strip_and_pack", "reset_prototypes", "SENTINEL"]
