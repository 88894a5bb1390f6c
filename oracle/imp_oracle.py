"""CPU restatement of IMP's prototype-fusion hot path.

TEST INFRASTRUCTURE ONLY.  Importers allowed: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package
(``interpretable-multimodal-prototyping_b200/``) never imports this file and has no CPU path.

Every function restates the arithmetic of the reference file:line it cites, in plain
torch-on-CPU (fp32 by default, fp64 on request).  Nothing is copied from the reference: the
functions are re-derived so that they (i) accept any number of prototypes P (the shipped
model only runs at P = 6) and (ii) never materialise the P x N x N tensor the
reference builds (``ops/utils.py:220``), so they scale to 16k / 120k patch bags.

Pinning: the reference ships no tests / golden vectors.  The oracle is pinned by
executing the UNMODIFIED reference code under ``oracle/ref_harness.py`` in the build container
(``tests/test_oracle_vs_reference.py``) and through the fixtures that run produced
(``tests/golden/*.npz`` made by ``tests/golden/make_golden.py``).  The k-means row (A9) has no
reference implementation at all: **parity unpinned** for A9 (only the distance formula of
``medmm/metrics/distance.py:46-61`` is reference-derived).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SENTINEL = -10000.0                       # data/data_manager.py:387 (pad value), umeml_gan.py:404
GROUP_SIZES = [82, 330, 513, 440, 1538, 451]   # umeml_gan.py:274


# --------------------------------------------------------------------------------------------
# A0  sentinel strip  (umeml_gan.py:401-410)
# --------------------------------------------------------------------------------------------
def bag_length(x_padded: torch.Tensor) -> int:
    """Row index of the first element equal to -10000 in row-major order (umeml_gan.py:404-409).

    The reference leaves ``x_path_this`` stale/undefined when no sentinel exists; here a bag
    without sentinel is taken at its full length (documented deviation).
    """
    hit = (x_padded == SENTINEL).any(dim=1)
    idx = torch.nonzero(hit)
    return int(idx[0, 0]) if idx.numel() else int(x_padded.shape[0])


def strip_bags(img: torch.Tensor) -> List[torch.Tensor]:
    """(B, Npad, 512) -> list of (N_i, 512) views."""
    return [img[i, : bag_length(img[i])] for i in range(img.shape[0])]


# --------------------------------------------------------------------------------------------
# A1  path_net  (umeml_gan.py:266-268, 410)
# --------------------------------------------------------------------------------------------
def path_net(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
             keep_mask: Optional[torch.Tensor] = None, p_drop: float = 0.0) -> torch.Tensor:
    """h = Dropout(ReLU(x W1^T + b1)).  ``keep_mask`` (N,256) of {0,1} injects the dropout
    mask (the reference RNG stream cannot be reproduced); kept values are scaled 1/(1-p)."""
    h = torch.relu(x @ w1.t() + b1)
    if keep_mask is not None:
        h = h * keep_mask.to(h.dtype) / (1.0 - p_drop)
    return h


# --------------------------------------------------------------------------------------------
# A3  single-head cross attention  (ops/attention.py:345-533, encoder-decoder branch)
# --------------------------------------------------------------------------------------------
def cross_attention(c: torch.Tensor, h: torch.Tensor, in_proj_weight: torch.Tensor,
                    in_proj_bias: torch.Tensor, out_w: torch.Tensor, out_b: torch.Tensor,
                    return_raw: bool = False):
    """c (P,D) queries, h (N,D) keys = values; heads = 1, scaling = D^-1/2 (attention.py:352).

    q = (c Wq^T + bq) * scaling (:368,:432); [k,v] = h Wkv^T + bkv (:382); S = q k^T (:509);
    A = softmax over patches (:527); o = (A v) Wo^T + bo (:530-533).  Returns (o, S) when
    ``return_raw`` (the module hands back the raw pre-softmax logits, :535-538).
    """
    d = c.shape[-1]
    wq, wk, wv = in_proj_weight[:d], in_proj_weight[d:2 * d], in_proj_weight[2 * d:]
    bq, bk, bv = in_proj_bias[:d], in_proj_bias[d:2 * d], in_proj_bias[2 * d:]
    q = (c @ wq.t() + bq) * (float(d) ** -0.5)
    k = h @ wk.t() + bk
    v = h @ wv.t() + bv
    s = q @ k.t()
    a = torch.softmax(s, dim=-1)
    o = (a @ v) @ out_w.t() + out_b
    return (o, s) if return_raw else o


def layer_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


# --------------------------------------------------------------------------------------------
# A2  PathProtoGenerator  (umeml_gan.py:65-80): c <- c + LayerNorm(MHA(c, h, h))
# --------------------------------------------------------------------------------------------
def proto_block(h: torch.Tensor, c: torch.Tensor, blk: Dict[str, torch.Tensor]) -> torch.Tensor:
    o = cross_attention(c, h, blk["in_proj_weight"], blk["in_proj_bias"],
                        blk["out_proj.weight"], blk["out_proj.bias"])
    return c + layer_norm(o, blk["norm1.weight"], blk["norm1.bias"])


def prototype_pool(x: torch.Tensor, p_proto: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
                   blocks: Sequence[Dict[str, torch.Tensor]],
                   keep_mask: Optional[torch.Tensor] = None, p_drop: float = 0.0):
    """One bag through path_net and the stacked PathProtoGenerator blocks
    (umeml_gan.py:410, 425-434).  Returns (c_final (P,D), h (N,D))."""
    h = path_net(x, w1, b1, keep_mask, p_drop)
    c = p_proto
    for blk in blocks:
        c = proto_block(h, c, blk)
    return c, h


# --------------------------------------------------------------------------------------------
# A4  cluster assignment  (ops/utils.py:178-181 as called from :214 with c.T of a 3-D tensor)
# --------------------------------------------------------------------------------------------
def cluster_assignment(x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """x (N,D), c (P,D) -> C (N,P) = relu(x_hat . c_hat).

    x_hat: patches L2-normalised over features (eps 1e-12).  c_hat[p,d] = c[p,d]/||c[:,d]||_2,
    i.e. normalised ACROSS PROTOTYPES per feature: the caller passes ``c.T`` of a (1,P,D)
    tensor = (D,P,1) and ``F.normalize(dim=1)`` then runs over P (ops/utils.py:180,214).
    """
    xh = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    ch = c / c.norm(dim=0, keepdim=True).clamp_min(1e-12)
    return torch.relu(xh @ ch.t())


# --------------------------------------------------------------------------------------------
# A5 + A6  modularity  (ops/utils.py:188-228), restated without N x N / P x N x N tensors
# --------------------------------------------------------------------------------------------
def modularity_literal(c: torch.Tensor, x: torch.Tensor, temp: float = 0.1) -> torch.Tensor:
    """Small-N literal form (materialises N x N and P x N x N); differentiable in c.
    Used only to validate ``modularity`` below and against the reference at N <= 4096."""
    x = x.detach()                                             # utils.py:208
    xh = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)      # :193
    a = torch.relu(xh @ xh.t())                                # :194
    a = a - torch.diag(torch.diagonal(a))                      # :198 zero the diagonal
    d = a.sum(dim=1, keepdim=True)                             # :199
    e = a.sum()                                                # :200
    w = a - (d @ d.t()) / e                                    # :201
    cm = cluster_assignment(x, c)                              # :214  (N,P)
    em = torch.tanh(cm.t()[:, :, None] * cm.t()[:, None, :] / temp)   # :220  (P,N,N)
    delta = em.max(dim=0).values                               # :221
    q = (w / e) @ delta                                        # :222
    return -torch.diagonal(q).sum() * 100.0                    # :225-228


def modularity(c: torch.Tensor, x: torch.Tensor, temp: float = 0.1, chunk: int = 512,
               want_grad: bool = True, acc_dtype: torch.dtype = torch.float64, gram_bf16: bool = False):
    """Chunked modularity loss and its gradient wrt c.

    loss = -100 * [ sum_ij A_ij delta_ij / e  -  sum_ij d_i d_j delta_ij / e^2 ],
    delta_ij = tanh(max_p C_ip C_jp / temp)   (tanh is monotone and C >= 0, so the max over
    prototypes commutes with tanh; equal to utils.py:220-228 because delta is symmetric).
    Gradient: dL/dC_ip = sum_j 2 g_ij (1-delta_ij^2)/temp * [p = argmax] * C_jp with
    g_ij = -100 (A_ij/e - d_i d_j/e^2); then back through relu / the two normalisations by
    autograd on the (N,P) assignment matrix.  Returns (loss, dc or None).
    """
    x = x.detach()
    n = x.shape[0]
    dev = x.device                       # runs wherever its inputs live (tests use fp64 on the GPU for 16k / 120k bags)
    cw = c.detach().clone().requires_grad_(want_grad)
    cm = cluster_assignment(x, cw)                                  # (N,P) with graph to cw
    cmd = cm.detach()
    xh = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    if gram_bf16:                        # the device keeps the Gram operand x_hat in bf16: same rounding point here
        xh = xh.to(torch.bfloat16).to(x.dtype)
    # pass 1: degrees
    d = torch.zeros(n, dtype=acc_dtype, device=dev)
    for i0 in range(0, n, chunk):
        a = torch.relu(xh[i0:i0 + chunk] @ xh.t())
        r = torch.arange(i0, min(i0 + chunk, n), device=dev)
        a[r - i0, r] = 0
        d[i0:i0 + chunk] = a.sum(dim=1, dtype=acc_dtype)
    e = d.sum()
    # pass 2: the two traces and dL/dC
    s1 = torch.zeros((), dtype=acc_dtype, device=dev)
    s2 = torch.zeros((), dtype=acc_dtype, device=dev)
    dcm = torch.zeros(cmd.shape, dtype=acc_dtype, device=dev)
    for i0 in range(0, n, chunk):
        ci = cmd[i0:i0 + chunk]                                      # (m,P)
        a = torch.relu(xh[i0:i0 + chunk] @ xh.t())
        r = torch.arange(i0, min(i0 + chunk, n), device=dev)
        a[r - i0, r] = 0
        u = torch.full((ci.shape[0], n), -1.0, dtype=cmd.dtype, device=dev)
        arg = torch.zeros((ci.shape[0], n), dtype=torch.long, device=dev)
        for p in range(cmd.shape[1]):                                # first max wins (torch.max)
            v = ci[:, p:p + 1] * cmd[None, :, p]
            better = v > u
            u = torch.where(better, v, u)
            arg = torch.where(better, torch.full_like(arg, p), arg)
        delta = torch.tanh(u / temp).to(acc_dtype)
        dd = d[i0:i0 + chunk, None] * d[None, :]
        s1 += (a.to(acc_dtype) * delta).sum()
        s2 += (dd * delta).sum()
        if want_grad:
            g = -100.0 * (a.to(acc_dtype) / e - dd / (e * e))
            wgt = 2.0 * g * (1.0 - delta * delta) / temp            # (m,N)
            cols = torch.arange(n, device=dev)[None, :].expand(ci.shape[0], -1)
            picked = cmd.to(acc_dtype)[cols, arg]                    # C_j,p*  (m,N)
            dcm[i0:i0 + chunk].scatter_add_(1, arg.reshape(ci.shape[0], -1), wgt * picked)
    loss = -100.0 * (s1 / e - s2 / (e * e))
    dc = None
    if want_grad:
        (dc,) = torch.autograd.grad(cm, cw, dcm.to(cm.dtype))
    return loss.to(c.dtype), dc


# --------------------------------------------------------------------------------------------
# A7  per-pathway omic encoders  (umeml_gan.py:274-283, 413-419)
# --------------------------------------------------------------------------------------------
def omic_encode(x_omic: torch.Tensor, group_index: Sequence[Sequence[int]],
                weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                keep_mask: Optional[torch.Tensor] = None, p_drop: float = 0.0) -> torch.Tensor:
    """x_omic (B,G) -> (B,K,256): o_k = Dropout(ReLU(x[:, idx_k] W_k^T + b_k)), concatenated."""
    outs = []
    for idx, w, b in zip(group_index, weights, biases):
        xi = x_omic[:, torch.as_tensor(list(idx), dtype=torch.long)]
        outs.append(torch.relu(xi @ w.t() + b))
    o = torch.stack(outs, dim=1)
    if keep_mask is not None:
        o = o * keep_mask.to(o.dtype) / (1.0 - p_drop)
    return o


# --------------------------------------------------------------------------------------------
# A8  missing-omics handling  (umeml_gan.py:380-392, 500-511)
# --------------------------------------------------------------------------------------------
def impute_missing_genes(x_omic: torch.Tensor, insample_without_omic: Optional[torch.Tensor],
                         omic_means: torch.Tensor) -> torch.Tensor:
    """x = where(mask, means, x) when any gene is masked (umeml_gan.py:391-392)."""
    if insample_without_omic is None or int(insample_without_omic.sum()) == 0:
        return x_omic
    return torch.where(insample_without_omic.bool(), omic_means[None, :].expand_as(x_omic), x_omic)


def generator_p2o(h_path: torch.Tensor, gw0, gb0, gw1, gb1) -> torch.Tensor:
    """Generator (umeml_gan.py:25-45): Linear -> ReLU -> Linear -> Softplus on flattened tokens."""
    b, t, d = h_path.shape
    z = torch.relu(h_path.reshape(b, -1) @ gw0.t() + gb0)
    return F.softplus(z @ gw1.t() + gb1).reshape(b, t, d)


def blend_missing_omics(h_omic: Optional[torch.Tensor], h_omic_gen: torch.Tensor,
                        without_omic: Optional[torch.Tensor],
                        insample_without_omic: Optional[torch.Tensor]) -> torch.Tensor:
    """Sample-level replace then feature-level blend (umeml_gan.py:500-511).
    r = mask.sum()/mask.numel() over the WHOLE batch (:509)."""
    if h_omic is None:                                                       # :506-507
        return h_omic_gen
    if without_omic is not None and int(without_omic.sum()) > 0:             # :503-505
        h_omic = torch.where((without_omic == 1).view(-1, 1, 1), h_omic_gen, h_omic)
    if insample_without_omic is not None and int(insample_without_omic.sum()) > 0:   # :508-511
        r = insample_without_omic.sum().to(h_omic.dtype) / insample_without_omic.numel()
        h_omic = (1 - r) * h_omic + r * h_omic_gen
    return h_omic


# --------------------------------------------------------------------------------------------
# A9  k-means assignment (new; distance of medmm/metrics/distance.py:46-61 + argmin)
# --------------------------------------------------------------------------------------------
def kmeans_assign(x: torch.Tensor, mu: torch.Tensor, chunk: int = 1 << 16) -> torch.Tensor:
    """argmin_k ||x_n||^2 + ||mu_k||^2 - 2 x_n.mu_k, first index on ties.  PARITY UNPINNED."""
    out = torch.empty(x.shape[0], dtype=torch.int32)
    m2 = (mu * mu).sum(dim=1)
    for i0 in range(0, x.shape[0], chunk):
        xi = x[i0:i0 + chunk]
        dist = (xi * xi).sum(dim=1, keepdim=True) + m2[None, :] - 2.0 * (xi @ mu.t())
        out[i0:i0 + chunk] = dist.argmin(dim=1).to(torch.int32)
    return out


def kmeans_update(x: torch.Tensor, assign: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lloyd update: per-centroid sums and counts (free design choice)."""
    sums = torch.zeros(k, x.shape[1], dtype=torch.float64)
    sums.index_add_(0, assign.long(), x.double())
    counts = torch.bincount(assign.long(), minlength=k)
    return sums, counts


# --------------------------------------------------------------------------------------------
# multi-GPU: log-sum-exp merge of partial softmax-pooling states
# --------------------------------------------------------------------------------------------
def pool_partial(h: torch.Tensor, qt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Partial state of softmax pooling over a shard of patches: (m (P), l (P), acc (P,D)) with
    scores S = qt h^T, m = max_n S, l = sum_n exp(S-m), acc = sum_n exp(S-m) h_n."""
    s = qt @ h.t()
    m = s.max(dim=1).values
    w = torch.exp(s - m[:, None])
    return m, w.sum(dim=1), w @ h


def lse_merge(parts: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]):
    """Merge partial states -> (pooled (P,D), lse (P))."""
    m = torch.stack([p[0] for p in parts]).max(dim=0).values
    l = sum(p[1] * torch.exp(p[0] - m) for p in parts)
    acc = sum(p[2] * torch.exp(p[0] - m)[:, None] for p in parts)
    return acc / l[:, None], m + torch.log(l)


def folded_query(c: torch.Tensor, in_proj_weight: torch.Tensor, in_proj_bias: torch.Tensor) -> torch.Tensor:
    """q~ = ((c Wq^T + bq)/sqrt(D)) Wk  (P,D): S_pn = h_n . q~_p + const_p, and the constant
    q_p . b_k cancels in the softmax over patches (verified exact)."""
    d = c.shape[-1]
    q = (c @ in_proj_weight[:d].t() + in_proj_bias[:d]) * (float(d) ** -0.5)
    return q @ in_proj_weight[d:2 * d]


# --------------------------------------------------------------------------------------------
# the whole hot path for a batch of bags: loss pieces + gradients (used as the parity checker
# and as bench.py's cpu_baseline "port")
# --------------------------------------------------------------------------------------------
def hot_path_step(bags: Sequence[torch.Tensor], params: Dict[str, torch.Tensor],
                  p_proto: torch.Tensor, with_modularity: bool = True,
                  grad_seed: Optional[torch.Tensor] = None, chunk: int = 512):
    """Forward + backward of path_net -> 2 x PathProtoGenerator (-> modularity) over ``bags``.

    params: 'path_net.0.weight', 'path_net.0.bias', and for b in {0,1}
    'proto_g_blocks.b.cross_attn.in_proj_weight' ... 'proto_g_blocks.b.norm1.bias'
    (the reference state_dict names).
    The scalar that is differentiated is  sum(c_final * grad_seed) + mean_j modularity_j,
    which exercises every gradient path of the reference training step that crosses the hot
    path (downstream token ops are represented by the fixed cotangent ``grad_seed``).
    Returns dict(c=(B,P,D), modularity=scalar, grads={name: tensor}).
    """
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    blocks = []
    for b in range(2):
        pre = "proto_g_blocks.%d." % b
        blocks.append({
            "in_proj_weight": leaves[pre + "cross_attn.in_proj_weight"],
            "in_proj_bias": leaves[pre + "cross_attn.in_proj_bias"],
            "out_proj.weight": leaves[pre + "cross_attn.out_proj.weight"],
            "out_proj.bias": leaves[pre + "cross_attn.out_proj.bias"],
            "norm1.weight": leaves[pre + "norm1.weight"],
            "norm1.bias": leaves[pre + "norm1.bias"],
        })
    cs, mods, total = [], [], 0.0
    for j, x in enumerate(bags):
        c, h = prototype_pool(x, p_proto, leaves["path_net.0.weight"], leaves["path_net.0.bias"], blocks)
        cs.append(c)
        if grad_seed is not None:
            total = total + (c * grad_seed[j]).sum()
        if with_modularity:
            loss_j, dc_j = modularity(c, h, chunk=chunk)
            mods.append(loss_j)
            total = total + (c * dc_j.detach()).sum() / len(bags)      # inject d(mean mod)/dc
    if isinstance(total, torch.Tensor):
        total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return {"c": torch.stack([c.detach() for c in cs]),
            "modularity": (torch.stack(mods).mean() if mods else torch.zeros(())),
            "grads": grads}


def survival_risk(logits: torch.Tensor) -> torch.Tensor:
    """risk = -sum_k prod_{j<=k} (1 - sigmoid(logit_j))  (evaluation/evaluator.py:369-382)."""
    hazards = torch.sigmoid(logits)
    return -torch.cumprod(1 - hazards, dim=1).sum(dim=1)
