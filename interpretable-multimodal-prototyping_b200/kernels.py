"""Thin tensor-level wrappers over the C-ABI (one function per ``imp_*`` entry point).

These only validate shapes/dtypes, allocate outputs and workspaces through torch's caching
allocator and pass raw device pointers plus the current CUDA stream.  No arithmetic happens
in Python and there is no fallback path: every function requires CUDA tensors.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

D = 256           # MODEL.HIDDEN_DIM (umeml_gan.py:247)
_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.ImpError("%s must be a CUDA tensor: the IMP hot path has no CPU implementation" % name)
    if t.dtype != dtype:
        raise TypeError("%s: expected %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _ptr_array(tensors: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def pathnet_fwd(x: torch.Tensor, w1_bf16: torch.Tensor, b1: torch.Tensor, p_drop: float = 0.0,
                seed: int = 0) -> torch.Tensor:
    """h = Dropout(ReLU(x W1^T + b1)) -> (rows,256) bf16   [umeml_gan.py:266-268,410]."""
    _chk(x, torch.bfloat16, "x"); _chk(w1_bf16, torch.bfloat16, "w1"); _chk(b1, torch.float32, "b1")
    rows, kin = x.shape
    h = torch.empty(rows, D, device=x.device, dtype=torch.bfloat16)
    if rows:
        _lib.call("imp_pathnet_fwd", x, w1_bf16, b1, h, rows, kin, float(p_drop), int(seed) & 0x7FFFFFFF,
                  _lib.stream_ptr())
    return h


def pathnet_dw(dz: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
               accumulate: bool = False) -> torch.Tensor:
    """dW1 (256,in) fp32 (+)= dz^T x."""
    _chk(dz, torch.bfloat16, "dz"); _chk(x, torch.bfloat16, "x")
    rows, kin = x.shape
    if out is None:
        out = torch.empty(D, kin, device=x.device, dtype=torch.float32)
        accumulate = False
    ws = _workspace(_lib.query("imp_pathnet_dw_workspace_bytes", kin), x.device)
    _lib.call("imp_pathnet_dw", dz, x, out, ws, rows, kin, bool(accumulate), _lib.stream_ptr())
    return out


def pool_fwd(h: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, qt: torch.Tensor
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Softmax pooling of each bag's patches into P tokens.  qt (B,P,256) or (1,P,256)/(P,256)
    shared.  Returns pooled (B,P,256) fp32, lse (B,P) fp32   [attention.py:509-530]."""
    _chk(h, torch.bfloat16, "h"); _chk(cu_seqlens, torch.int32, "cu_seqlens"); _chk(qt, torch.float32, "qt")
    nb = cu_seqlens.numel() - 1
    if qt.dim() == 2:
        qt = qt.unsqueeze(0)
    p = qt.shape[1]
    stride = 0 if qt.shape[0] == 1 and nb != 1 else p * D
    if qt.shape[0] not in (1, nb):
        raise ValueError("qt batch %d does not match %d bags" % (qt.shape[0], nb))
    if h.shape[0] == 0 or max_len <= 0:          # an empty shard of a giant bag: no mass
        return (torch.zeros(nb, p, D, device=h.device, dtype=torch.float32),
                torch.full((nb, p), float("-inf"), device=h.device, dtype=torch.float32))
    pooled = torch.empty(nb, p, D, device=h.device, dtype=torch.float32)
    lse = torch.empty(nb, p, device=h.device, dtype=torch.float32)
    ws = _workspace(_lib.query("imp_pool_fwd_workspace_bytes", nb, int(max_len), p), h.device)
    _lib.call("imp_pool_fwd", h, h.shape[0], cu_seqlens, nb, int(max_len), qt, ctypes.c_longlong(stride), p,
              ws, pooled, lse, _lib.stream_ptr())
    return pooled, lse


def pool_bwd(h: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, qt: Sequence[torch.Tensor],
             dpooled: Sequence[torch.Tensor], lse: Sequence[torch.Tensor], delta: Sequence[torch.Tensor],
             dq_block: int, want_dz: bool, relu_mask: bool = True, keep_scale: float = 1.0,
             db1: Optional[torch.Tensor] = None, db_accumulate: bool = False):
    """Backward of 1 or 2 stacked pooling blocks sharing h.  Returns (dq (B,P,256), dz or None)."""
    nblk = len(qt)
    nb = cu_seqlens.numel() - 1
    _chk(h, torch.bfloat16, "h"); _chk(cu_seqlens, torch.int32, "cu_seqlens")
    qs, strides = [], []
    for k in range(nblk):
        q = qt[k] if qt[k].dim() == 3 else qt[k].unsqueeze(0)
        _chk(q, torch.float32, "qt"); _chk(dpooled[k], torch.float32, "dpooled")
        _chk(lse[k], torch.float32, "lse"); _chk(delta[k], torch.float32, "delta")
        qs.append(q)
        strides.append(0 if q.shape[0] == 1 and nb != 1 else q.shape[1] * D)
    p = qs[0].shape[1]
    if h.shape[0] == 0 or max_len <= 0:
        if want_dz and db1 is not None and not db_accumulate:
            db1.zero_()
        return torch.zeros(nb, p, D, device=h.device, dtype=torch.float32), (torch.empty_like(h) if want_dz else None)
    dq = torch.empty(nb, p, D, device=h.device, dtype=torch.float32)
    dz = torch.empty_like(h) if want_dz else None
    ws = _workspace(_lib.query("imp_pool_bwd_workspace_bytes", nb, int(max_len), p), h.device)
    _lib.call("imp_pool_bwd", h, h.shape[0], cu_seqlens, nb, int(max_len), nblk, _ptr_array(qs),
              (ctypes.c_longlong * nblk)(*strides), _ptr_array(list(dpooled)), _ptr_array(list(lse)),
              _ptr_array(list(delta)), p, int(dq_block), bool(relu_mask), float(keep_scale), ws, dq, dz,
              db1 if want_dz else None, bool(db_accumulate), _lib.stream_ptr())
    return dq, dz


def cast_bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even) on the device."""
    _chk(t, torch.float32, "t")
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    n = t.numel()
    if n % 4:
        raise ValueError("cast_bf16: element count must be a multiple of 4")
    if n:
        _lib.call("imp_cast_bf16", t, out, ctypes.c_size_t(n), _lib.stream_ptr())
    return out


def bag_lengths(img: torch.Tensor, sentinel: float = -10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """img (B,Npad,D) fp32 -> (lengths (B) int32, cu_seqlens (B+1) int32), all on the device
    [umeml_gan.py:401-410]."""
    _chk(img, torch.float32, "img")
    b, npad, d = img.shape
    lengths = torch.empty(b, device=img.device, dtype=torch.int32)
    cu = torch.empty(b + 1, device=img.device, dtype=torch.int32)
    _lib.call("imp_bag_lengths", img, b, npad, d, float(sentinel), lengths, cu, _lib.stream_ptr())
    return lengths, cu


def pack_bags(img: torch.Tensor, cu_seqlens: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """valid rows of img (B,Npad,D) fp32 -> out (>= cu[B], D) bf16 packed."""
    _chk(img, torch.float32, "img"); _chk(cu_seqlens, torch.int32, "cu_seqlens"); _chk(out, torch.bfloat16, "out")
    b, npad, d = img.shape
    _lib.call("imp_pack_bags", img, b, npad, d, cu_seqlens, out, _lib.stream_ptr())
    return out


def modularity(h: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, chat: torch.Tensor, n_tok1: int,
               n_tok2: int = 0, temp: float = 0.1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Modularity loss per (bag, token group) and its gradient wrt the normalised tokens
    [ops/utils.py:178-228].  chat (B, n_tok1+n_tok2, 256) fp32 -> (loss (B,2), dchat like chat)."""
    _chk(h, torch.bfloat16, "h"); _chk(cu_seqlens, torch.int32, "cu_seqlens"); _chk(chat, torch.float32, "chat")
    nb = cu_seqlens.numel() - 1
    if chat.shape != (nb, n_tok1 + n_tok2, D):
        raise ValueError("chat shape %s != (%d,%d,%d)" % (tuple(chat.shape), nb, n_tok1 + n_tok2, D))
    loss = torch.empty(nb, 2, device=h.device, dtype=torch.float32)
    dchat = torch.empty_like(chat)
    nbytes = _lib.query("imp_modularity_workspace_bytes", h.shape[0], nb, n_tok1, n_tok2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=h.device)       # 256-byte aligned by the allocator
    _lib.call("imp_modularity", h, h.shape[0], cu_seqlens, nb, int(max_len), chat, int(n_tok1), int(n_tok2),
              float(temp), ws, loss, dchat, _lib.stream_ptr())
    return loss, dchat


_shard_cache = {}
_cu_cache = {}


def _one_bag_cu(total_rows: int, device) -> torch.Tensor:
    """cu_seqlens [0, total_rows] of a single bag, created once per (size, device): a host->device copy per step would
    also be illegal inside a CUDA-graph capture."""
    key = (total_rows, device)
    cu = _cu_cache.get(key)
    if cu is None:
        cu = _cu_cache[key] = torch.tensor([0, total_rows], dtype=torch.int32, device=device)
    return cu


def _cached(name: str, nbytes: int, device) -> torch.Tensor:
    """Reusable byte buffer for the sharded path (the 80 MB modularity workspace, the exchange staging)."""
    key = (name, device, torch.cuda.current_stream(device).cuda_stream)
    buf = _shard_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _shard_cache[key] = buf
    return buf


def _exchange_sections(ws: torch.Tensor, base: int, unit: int, windows, per_units: int, rank: int, name: str, group):
    """Every rank owns the byte range [base + lo*unit, base + hi*unit) of ``ws`` (its window ``(lo, hi)`` of
    ``windows``, in units); after the call every rank holds all of them.  One ``all_gather_into_tensor`` of equal
    ``per_units * unit`` chunks through a staging buffer (the last windows may be shorter or empty, so the padded
    chunks cannot be gathered in place), then one device copy per remote window."""
    import torch.distributed as dist
    world = len(windows)
    chunk = per_units * unit
    if chunk == 0:
        return
    stage = _cached(name, (world + 1) * chunk, ws.device)
    send, recv = stage[:chunk], stage[chunk:(world + 1) * chunk]
    lo, hi = windows[rank]
    if hi > lo:
        send[:(hi - lo) * unit].copy_(ws[base + lo * unit: base + hi * unit])
    dist.all_gather_into_tensor(recv, send, group=group)
    for r, (a, b) in enumerate(windows):
        if r != rank and b > a:
            ws[base + a * unit: base + b * unit].copy_(recv[r * chunk: r * chunk + (b - a) * unit])


def modularity_sharded(h_local: torch.Tensor, row_offset: int, total_rows: int, chat: torch.Tensor, n_tok1: int,
                       n_tok2: int, temp: float, group) -> Tuple[torch.Tensor, torch.Tensor]:
    """One bag of ``total_rows`` patches sharded by ``parallel.shard_bounds(total_rows, world)``: this rank holds the
    rows [row_offset, row_offset + h_local.shape[0]).  Returns the GLOBAL (loss (1,2), dchat (1,Pt,256)), identical on
    every rank of ``group``: prepare (local rows) -> exchange of x_hat / fixed-point assignments (one all-gather
    each; the windows are known on the host, no device->host sync) and column sums / sign flag (all-reduce) ->
    sweep of the local rows against all columns -> sum of the partial results."""
    import torch.distributed as dist
    from . import parallel
    _chk(h_local, torch.bfloat16, "h_local"); _chk(chat, torch.float32, "chat")
    dev = chat.device
    local_rows = int(h_local.shape[0])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    windows = parallel.shard_bounds(int(total_rows), world, 64)
    # validated BEFORE the first collective: a rank that raised inside them would leave the others hanging
    if (int(row_offset), int(row_offset) + local_rows) != windows[rank] and not (local_rows == 0 and windows[rank][0] == windows[rank][1]):
        raise ValueError("rank %d holds rows [%d,%d) but shard_bounds(%d, %d) assigns %s"
                         % (rank, row_offset, row_offset + local_rows, total_rows, world, windows[rank]))
    per_rows = windows[0][1] - windows[0][0]                      # rows of a full shard: a multiple of 64
    nbytes = _lib.query("imp_modularity_workspace_bytes", total_rows, 1, n_tok1, n_tok2)
    ws = _cached("modularity_ws", nbytes, dev)
    cu = _one_bag_cu(int(total_rows), dev)
    offs, sizes = (ctypes.c_size_t * 4)(), (ctypes.c_size_t * 4)()
    _lib.call("imp_modularity_sections", total_rows, 1, n_tok1, n_tok2, offs, sizes)
    _lib.call("imp_modularity_prepare", h_local, local_rows, int(row_offset), int(total_rows), cu, 1, chat, int(n_tok1),
              int(n_tok2), ws, _lib.stream_ptr())
    ntile_total = (total_rows + 128 + 64 + 63) // 64
    tile_bytes = int(sizes[1]) // ntile_total
    _exchange_sections(ws, int(offs[0]), 512, windows, per_rows, rank, "xh_stage", group)          # x_hat rows, 512 B each
    tile_windows = [(lo // 64, (hi + 63) // 64) for lo, hi in windows]
    _exchange_sections(ws, int(offs[1]), tile_bytes, tile_windows, per_rows // 64, rank, "lf_stage", group)
    colsum = ws[int(offs[2]): int(offs[2]) + int(sizes[2])].view(torch.float32)
    flag = ws[int(offs[3]): int(offs[3]) + 4].view(torch.int32)
    dist.all_reduce(colsum, group=group)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    loss = torch.empty(1, 2, device=dev, dtype=torch.float32)
    dchat = torch.empty_like(chat)
    _lib.call("imp_modularity_execute", h_local, local_rows, int(row_offset), int(total_rows), cu, 1, int(total_rows),
              int(n_tok1), int(n_tok2), float(temp), ws, loss, dchat, _lib.stream_ptr())
    packed = torch.cat([loss.reshape(-1), dchat.reshape(-1)])
    dist.all_reduce(packed, group=group)
    return packed[:2].view(1, 2), packed[2:].view_as(chat)


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def omic_encode_fwd(x_omic: torch.Tensor, gene_index: torch.Tensor, group_offsets: Sequence[int],
                    weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                    insample_mask: Optional[torch.Tensor] = None, omic_means: Optional[torch.Tensor] = None,
                    p_drop: float = 0.0, seed: int = 0) -> torch.Tensor:
    """(B,G) fp32 -> (B,K,256) fp32   [umeml_gan.py:274-283,391-392,413-419]."""
    _chk(x_omic, torch.float32, "x_omic"); _chk(gene_index, torch.int32, "gene_index")
    if insample_mask is not None:
        _chk(insample_mask, torch.int32, "insample_mask"); _chk(omic_means, torch.float32, "omic_means")
    for w, b in zip(weights, biases):
        _chk(w, torch.float32, "weight"); _chk(b, torch.float32, "bias")
    bsz, g = x_omic.shape
    k = len(weights)
    out = torch.empty(bsz, k, D, device=x_omic.device, dtype=torch.float32)
    _lib.call("imp_omic_encode_fwd", x_omic, insample_mask, omic_means if insample_mask is not None else None,
              gene_index, _int_array(group_offsets), k, _ptr_array(list(weights)), _ptr_array(list(biases)), bsz, g,
              float(p_drop), int(seed) & 0x7FFFFFFF, out, _lib.stream_ptr())
    return out


def omic_encode_bwd(x_omic, gene_index, group_offsets, out, dout, weights_like: Sequence[torch.Tensor],
                    insample_mask=None, omic_means=None, p_drop: float = 0.0):
    """-> (list of dW_k (256,G_k), list of db_k (256))."""
    _chk(out, torch.float32, "out"); _chk(dout, torch.float32, "dout")
    bsz, g = x_omic.shape
    k = len(weights_like)
    dws = [torch.empty_like(w) for w in weights_like]
    dbs = [torch.empty(D, device=x_omic.device, dtype=torch.float32) for _ in range(k)]
    _lib.call("imp_omic_encode_bwd", x_omic, insample_mask, omic_means if insample_mask is not None else None,
              gene_index, _int_array(group_offsets), k, bsz, g, float(p_drop), out, dout, _ptr_array(dws),
              _ptr_array(dbs), False, _lib.stream_ptr())
    return dws, dbs


def omic_blend(h_omic: torch.Tensor, h_gen: torch.Tensor, without_omic: Optional[torch.Tensor] = None,
               insample_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sample-level replacement + batch-ratio blend of the omic tokens [umeml_gan.py:500-511].
    Returns (blended tokens, r (1) fp32)."""
    _chk(h_omic, torch.float32, "h_omic"); _chk(h_gen, torch.float32, "h_gen")
    if without_omic is not None:
        _chk(without_omic, torch.int32, "without_omic")
    if insample_mask is not None:
        _chk(insample_mask, torch.int32, "insample_mask")
    bsz = h_omic.shape[0]
    per = h_omic.numel() // bsz
    out = torch.empty_like(h_omic)
    scratch = torch.empty(2, device=h_omic.device, dtype=torch.float32)
    _lib.call("imp_omic_blend", h_omic, h_gen, without_omic, insample_mask,
              ctypes.c_longlong(insample_mask.numel() if insample_mask is not None else 0), bsz, per,
              scratch[0:1], out, scratch[1:2], _lib.stream_ptr())
    return out, scratch[1:2]


def kmeans_assign(x: torch.Tensor, centroids: torch.Tensor, want_dist: bool = False):
    """argmin_k ||x_n - mu_k||^2 (first index on ties) -> int32 (N)   [metrics/distance.py:46-61 + argmin]."""
    _chk(x, torch.float32, "x"); _chk(centroids, torch.float32, "centroids")
    n, d = x.shape
    k = centroids.shape[0]
    assign = torch.empty(n, device=x.device, dtype=torch.int32)
    dist = torch.empty(n, device=x.device, dtype=torch.float32) if want_dist else None
    _lib.call("imp_kmeans_assign", x, centroids, n, d, k, assign, dist, _lib.stream_ptr())
    return (assign, dist) if want_dist else assign


def kmeans_update(x: torch.Tensor, assign: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-centroid sums (k,D) fp32 and counts (k) int32 of the assigned rows."""
    _chk(x, torch.float32, "x"); _chk(assign, torch.int32, "assign")
    n, d = x.shape
    sums = torch.zeros(k, d, device=x.device, dtype=torch.float32)
    counts = torch.zeros(k, device=x.device, dtype=torch.int32)
    _lib.call("imp_kmeans_update", x, assign, n, d, int(k), sums, counts, _lib.stream_ptr())
    return sums, counts


def lse_merge(part_pooled: torch.Tensor, part_lse: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """part_pooled (B,S,P,256), part_lse (B,S,P): per-shard pooling results -> whole-bag (pooled, lse)."""
    _chk(part_pooled, torch.float32, "part_pooled"); _chk(part_lse, torch.float32, "part_lse")
    b, s, p, _ = part_pooled.shape
    pooled = torch.empty(b, p, D, device=part_pooled.device, dtype=torch.float32)
    lse = torch.empty(b, p, device=part_pooled.device, dtype=torch.float32)
    scratch = torch.empty(b * s * 2 * p, device=part_pooled.device, dtype=torch.float32)
    _lib.call("imp_lse_merge", part_pooled, part_lse, b, s, p, pooled, lse, scratch, _lib.stream_ptr())
    return pooled, lse
