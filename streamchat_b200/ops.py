"""Thin tensor-level wrappers over the C-ABI (device pointers in, device pointers out).

PyTorch is used for device memory and streams only; all arithmetic happens in
libstreamchat_b200.so.  Every function raises if its tensors are not CUDA tensors — there is
no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

_workspaces = {}


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.SCB_F32
    if t.dtype == torch.float16:
        return L.SCB_F16
    raise TypeError(f"unsupported dtype {t.dtype} (float32 / float16 only)")


def _require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise L.ScbError("streamchat_b200 ops need CUDA tensors: there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise L.ScbError(f"tensors on different devices: {dev} vs {t.device}")
    assert dev is not None
    return dev


def _workspace(dev: torch.device, name: str, nbytes: int) -> torch.Tensor:
    """Per-device scratch buffer, allocated ZEROED (kernels with self-resetting tickets rely on it) and grown on demand."""
    key = (dev.index, name)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def _stream(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------
def cosine_topk(index: torch.Tensor, queries: torch.Tensor, k: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k cosine similarity of each query against every row of `index`.

    index [N, D] (fp32/fp16, rows not normalised, last dim contiguous), queries [Q, D] or [D].
    Returns (scores [Q, k] fp32, indices [Q, k] int64) ordered by (score desc, row asc); slots
    beyond N hold (-inf, -1).  Math of sentence_transformers cos_sim as used at reference
    utiles.py:732,764.
    """
    dev = _require_cuda(index, queries)
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    if index.dim() != 2 or queries.dim() != 2 or index.shape[1] != queries.shape[1]:
        raise ValueError(f"shape mismatch: index {tuple(index.shape)} queries {tuple(queries.shape)}")
    if index.stride(1) != 1 and index.shape[0] > 0:
        index = index.contiguous()
    if queries.stride(1) != 1:
        queries = queries.contiguous()
    n, d = index.shape
    q = queries.shape[0]
    lib = L.load()
    out_i = torch.empty((q, k), dtype=torch.int64, device=dev)
    out_s = torch.empty((q, k), dtype=torch.float32, device=dev)
    nbytes = lib.scb_cosine_topk_workspace_bytes(n, q, k)
    if nbytes == 0:
        raise ValueError(f"invalid cosine_topk configuration: n_queries={q} k={k}")
    ws = _workspace(dev, "topk", nbytes)
    a = L.CosineTopkArgs(
        index=_ptr(index) if n > 0 else None, queries=_ptr(queries), n_rows=n, dim=d, n_queries=q,
        ld_index=index.stride(0) if n > 0 else d, ld_query=queries.stride(0) if q > 1 else d,
        index_dtype=_dtype_code(index), query_dtype=_dtype_code(queries), k=k,
        out_indices=_ptr(out_i), out_scores=_ptr(out_s),
        workspace=_ptr(ws), workspace_bytes=ws.numel())
    with torch.cuda.device(dev):
        L.check(lib.scb_cosine_topk(C.byref(a), _stream(dev)), "scb_cosine_topk")
    return out_s, out_i


# ------------------------------------------------------------------------------------------
def kmeans_step(points: torch.Tensor, centroids: torch.Tensor, weights: Optional[torch.Tensor] = None,
                dist_eps: float = 0.0, do_update: bool = True, want_dist: bool = False):
    """One Lloyd iteration for a batch of centroid sets over the same points.

    points [N, D] fp16/fp32 contiguous; centroids [B, K, D] fp32 contiguous.
    Returns dict(labels [B,N] int32, dist [B,N] | None, centroids [B,K,D] fp32, wsum [B,K],
    shift [B,K]) — the last three only when do_update.
    """
    dev = _require_cuda(points, centroids, weights)
    if points.dim() != 2 or centroids.dim() != 3 or centroids.shape[2] != points.shape[1]:
        raise ValueError(f"shape mismatch: points {tuple(points.shape)} centroids {tuple(centroids.shape)}")
    if centroids.dtype != torch.float32:
        raise TypeError("centroids must be float32")
    points = points.contiguous()
    centroids = centroids.contiguous()
    n, d = points.shape
    b, k, _ = centroids.shape
    if weights is not None:
        weights = weights.to(torch.float32).contiguous()
        if weights.shape != (n,):
            raise ValueError("weights must have shape [N]")
    lib = L.load()
    labels = torch.empty((b, n), dtype=torch.int32, device=dev)
    dist = torch.empty((b, n), dtype=torch.float32, device=dev) if want_dist else None
    new_c = torch.empty_like(centroids) if do_update else None
    wsum = torch.empty((b, k), dtype=torch.float32, device=dev) if do_update else None
    shift = torch.empty((b, k), dtype=torch.float32, device=dev) if do_update else None
    nbytes = lib.scb_kmeans_step_workspace_bytes(n, k, b, d)
    if nbytes == 0:
        raise ValueError("invalid kmeans_step configuration")
    ws = _workspace(dev, "kmeans", nbytes)
    a = L.KmeansStepArgs(
        points=_ptr(points), weights=_ptr(weights), centroids=_ptr(centroids),
        n_points=n, n_clusters=k, n_batch=b, points_dtype=_dtype_code(points), dim=d,
        dist_eps=dist_eps, do_update=int(do_update),
        out_labels=_ptr(labels), out_dist=_ptr(dist), out_centroids=_ptr(new_c),
        out_wsum=_ptr(wsum), out_shift=_ptr(shift),
        workspace=_ptr(ws), workspace_bytes=ws.numel())
    with torch.cuda.device(dev):
        L.check(lib.scb_kmeans_step(C.byref(a), _stream(dev)), "scb_kmeans_step")
    return {"labels": labels, "dist": dist, "centroids": new_c, "wsum": wsum, "shift": shift}


# ------------------------------------------------------------------------------------------
def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, act: int = L.ACT_NONE,
           out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float16,
           alpha: float = 1.0) -> torch.Tensor:
    """out = act(alpha * a @ w.T + bias) + residual on tcgen05 tensor cores.

    a [M, K] or [B, M, K] fp16 (last dim contiguous), w [N, K] or [B, N, K] fp16.  A 2-D `w`
    with 3-D `a` is applied to the flattened rows.  fp32 accumulation in TMEM.
    """
    dev = _require_cuda(a, w, bias, residual, out)
    if a.dtype != torch.float16 or w.dtype != torch.float16:
        raise TypeError("linear: a and w must be float16")
    lead_shape = None
    if a.dim() == 3 and w.dim() == 2:
        lead_shape = a.shape[:2]
        a = a.reshape(-1, a.shape[-1])
        if residual is not None:
            residual = residual.reshape(-1, residual.shape[-1])
        if out is not None:
            out = out.view(-1, out.shape[-1])
    batched = a.dim() == 3
    if a.stride(-1) != 1:
        a = a.contiguous()
    if w.stride(-1) != 1:
        w = w.contiguous()
    if batched:
        bsz, m, k = a.shape
        n = w.shape[1]
        if w.shape[0] != bsz or w.shape[2] != k:
            raise ValueError(f"batched linear shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
        lda, ldw, sa, sw = a.stride(1), w.stride(1), a.stride(0), w.stride(0)
    else:
        bsz = 1
        m, k = a.shape
        n = w.shape[0]
        if w.shape[1] != k:
            raise ValueError(f"linear shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
        lda, ldw, sa, sw = a.stride(0), w.stride(0), 0, 0
    if lda % 8 or ldw % 8 or (batched and (sa % 8 or sw % 8)):
        raise ValueError("linear: row/batch strides of a and w must be multiples of 8 elements")
    n_out = n // 2 if act == L.ACT_SILU_PAIR else n
    oshape = (bsz, m, n_out) if batched else (m, n_out)
    if out is None:
        out = torch.empty(oshape, dtype=out_dtype, device=dev)
    elif tuple(out.shape) != oshape or out.stride(-1) != 1:
        raise ValueError(f"linear: out must have shape {oshape} with contiguous rows")
    if bias is not None:
        if bias.dtype != torch.float16 or bias.shape != (n,):
            raise ValueError("linear: bias must be float16 [N]")
        bias = bias.contiguous()
    ldr = sr = 0
    if residual is not None:
        if residual.dtype != torch.float16 or tuple(residual.shape) != oshape or residual.stride(-1) != 1:
            raise ValueError("linear: residual must be float16 with the output's shape")
        ldr = residual.stride(-2)
        sr = residual.stride(0) if batched else 0
    args = L.LinearArgs(
        a=_ptr(a), w=_ptr(w), bias=_ptr(bias), residual=_ptr(residual), out=_ptr(out),
        m=m, n=n, k=k, lda=lda, ldw=ldw, ldr=ldr, ldo=out.stride(-2),
        act=act, out_dtype=_dtype_code(out), batch=bsz,
        stride_a=sa, stride_w=sw, stride_r=sr, stride_o=out.stride(0) if batched else 0,
        alpha=alpha)
    with torch.cuda.device(dev):
        L.check(L.load().scb_linear(C.byref(args), _stream(dev)), "scb_linear")
    if lead_shape is not None:
        out = out.view(*lead_shape, n_out)
    return out


# ------------------------------------------------------------------------------------------
def _call(name: str, dev: torch.device, *args) -> None:
    with torch.cuda.device(dev):
        L.check(getattr(L.load(), name)(*args, _stream(dev)), name)


def _f16(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float16:
        raise TypeError(f"{what} must be float16")
    return t


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, causal: bool = False,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused attention.  q [B, Tq, Hq, D], k/v [B, Tk, Hkv, D] — any strided views whose last dim
    is contiguous (e.g. slices of a fused qkv buffer, or a [Hkv, T, D] cache permuted to this
    order).  Returns out [B, Tq, Hq, D] fp16 (contiguous unless `out` is given)."""
    dev = _require_cuda(q, k, v, out)
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _f16(t, n)
        if t.dim() != 4 or t.stride(3) != 1:
            raise ValueError(f"{n} must be [B, T, H, D] with a contiguous last dim")
    B, Tq, Hq, D = q.shape
    _, Tk, Hkv, _ = k.shape
    if out is None:
        out = torch.empty((B, Tq, Hq, D), dtype=torch.float16, device=dev)
    a = L.AttentionArgs(
        q=_ptr(q), k=_ptr(k), v=_ptr(v), out=_ptr(out), batch=B, n_q_heads=Hq, n_kv_heads=Hkv, head_dim=D,
        q_len=Tq, kv_len=Tk,
        q_stride_b=q.stride(0), q_stride_t=q.stride(1), q_stride_h=q.stride(2),
        k_stride_b=k.stride(0), k_stride_t=k.stride(1), k_stride_h=k.stride(2),
        v_stride_b=v.stride(0), v_stride_t=v.stride(1), v_stride_h=v.stride(2),
        o_stride_b=out.stride(0), o_stride_t=out.stride(1), o_stride_h=out.stride(2),
        scale=scale, causal=int(causal))
    _call("scb_attention", dev, C.byref(a))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _require_cuda(x, gamma, beta, out)
    x2 = _f16(x, "x").reshape(-1, x.shape[-1])
    if out is None:
        out = torch.empty_like(x2)
    o2 = out.view(-1, x.shape[-1])
    _call("scb_layernorm", dev, _ptr(x2), x2.stride(0), _ptr(gamma), _ptr(beta), eps, x2.shape[0],
          x2.shape[1], _ptr(o2), o2.stride(0))
    return out.view(x.shape)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _require_cuda(x, weight, out)
    x2 = _f16(x, "x").reshape(-1, x.shape[-1])
    if out is None:
        out = torch.empty_like(x2)
    o2 = out.view(-1, x.shape[-1])
    _call("scb_rmsnorm", dev, _ptr(x2), x2.stride(0), _ptr(weight), eps, x2.shape[0], x2.shape[1],
          _ptr(o2), o2.stride(0))
    return out.view(x.shape)


def im2col_patch(pixels: torch.Tensor, patch: int, ld_out: int) -> torch.Tensor:
    dev = _require_cuda(pixels)
    pixels = _f16(pixels, "pixels").contiguous()
    n, c, h, w = pixels.shape
    out = torch.empty((n * (h // patch) * (w // patch), ld_out), dtype=torch.float16, device=dev)
    _call("scb_im2col_patch", dev, _ptr(pixels), _ptr(out), n, c, h, w, patch, ld_out)
    return out


def vit_embed_layernorm(patch_embeds: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor,
                        gamma: torch.Tensor, beta: torch.Tensor, eps: float, n_img: int) -> torch.Tensor:
    dev = _require_cuda(patch_embeds, cls, pos, gamma, beta)
    tokens, dim = pos.shape
    out = torch.empty((n_img * tokens, dim), dtype=torch.float16, device=dev)
    _call("scb_vit_embed_layernorm", dev, _ptr(patch_embeds), patch_embeds.stride(0), _ptr(cls), _ptr(pos),
          _ptr(gamma), _ptr(beta), eps, n_img, tokens, dim, _ptr(out), out.stride(0))
    return out


def text_embed_layernorm(ids: torch.Tensor, position_ids: torch.Tensor, word: torch.Tensor, pos: torch.Tensor,
                         token_type_row: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                         eps: float) -> torch.Tensor:
    """BERT embeddings + LayerNorm for a packed token list: ids / position_ids int64 [n] -> [n, D] fp16."""
    dev = _require_cuda(ids, position_ids, word, pos, token_type_row, gamma, beta)
    if ids.dtype != torch.int64 or position_ids.dtype != torch.int64 or ids.shape != position_ids.shape:
        raise ValueError("ids and position_ids must be int64 tensors of the same shape")
    n, dim = ids.numel(), word.shape[1]
    out = torch.empty((n, dim), dtype=torch.float16, device=dev)
    _call("scb_text_embed_layernorm", dev, _ptr(ids), _ptr(position_ids), n, _ptr(word), _ptr(pos),
          _ptr(token_type_row), _ptr(gamma), _ptr(beta), eps, dim, _ptr(out), out.stride(0))
    return out


def mean_pool_normalize(hidden: torch.Tensor, offsets: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """Packed token rows [T, D] fp16 + int32 offsets [n + 1] -> [n, D] fp32 mean-pooled (and L2-normalised) rows
    (sentence-transformers Pooling(mean) + Normalize, the dialogue-memory embedding)."""
    dev = _require_cuda(hidden, offsets)
    if offsets.dtype != torch.int32 or offsets.dim() != 1 or offsets.numel() < 2:
        raise TypeError("offsets must be int32 [n + 1]")
    _f16(hidden, "hidden")
    n, dim = offsets.numel() - 1, hidden.shape[1]
    out = torch.empty((n, dim), dtype=torch.float32, device=dev)
    _call("scb_mean_pool_normalize", dev, _ptr(hidden), hidden.stride(0), _ptr(offsets), n, dim, _ptr(out), out.stride(0),
          int(normalize))
    return out


def avgpool2x2_tokens(x: torch.Tensor, n_img: int, tokens_per_img: int, tok_offset: int, grid: int) -> torch.Tensor:
    """x [n_img*tokens_per_img, D] -> [n_img*(grid/2)^2, D]."""
    dev = _require_cuda(x)
    dim = x.shape[-1]
    out = torch.empty((n_img * (grid // 2) ** 2, dim), dtype=torch.float16, device=dev)
    _call("scb_avgpool2x2_tokens", dev, _ptr(x), x.stride(0), tokens_per_img, tok_offset, grid, n_img, dim,
          _ptr(out), out.stride(0))
    return out


def silu_mul(gate_up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _require_cuda(gate_up, out)
    rows, two_i = gate_up.shape
    inter = two_i // 2
    if out is None:
        out = torch.empty((rows, inter), dtype=torch.float16, device=dev)
    _call("scb_silu_mul", dev, _ptr(gate_up), gate_up.stride(0), rows, inter, _ptr(out), out.stride(0))
    return out


def rope_append(qkv: torch.Tensor, n_q: int, n_kv: int, head_dim: int, inv_freq: Optional[torch.Tensor],
                theta: float, pos_base: int = 0, positions: Optional[torch.Tensor] = None,
                k_cache: Optional[torch.Tensor] = None, v_cache: Optional[torch.Tensor] = None) -> None:
    """In-place RoPE on the q,k slices of qkv [rows, (n_q+2*n_kv)*head_dim]; appends k,v to the
    caches [n_kv, max_pos, head_dim] when given."""
    dev = _require_cuda(qkv, inv_freq, positions, k_cache, v_cache)
    max_pos = k_cache.shape[1] if k_cache is not None else 0
    _call("scb_rope_append", dev, _ptr(qkv), qkv.stride(0), qkv.shape[0], n_q, n_kv, head_dim,
          _ptr(positions), pos_base, theta, _ptr(inv_freq), _ptr(k_cache), _ptr(v_cache), max_pos)


def gather_rows(table: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _require_cuda(table, ids, out)
    ids = ids.reshape(-1)
    if ids.dtype != torch.int64:
        raise TypeError("ids must be int64")
    if out is None:
        out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float16, device=dev)
    _call("scb_gather_rows", dev, _ptr(table), _ptr(ids), ids.numel(), table.shape[1], _ptr(out), out.stride(0))
    return out


def gather_rows_rmsnorm(table: torch.Tensor, ids: torch.Tensor, gamma: torch.Tensor, eps: float,
                        out: torch.Tensor, norm_out: torch.Tensor) -> None:
    dev = _require_cuda(table, ids, gamma, out, norm_out)
    ids = ids.reshape(-1)
    _call("scb_gather_rows_rmsnorm", dev, _ptr(table), _ptr(ids), ids.numel(), table.shape[1], _ptr(out),
          out.stride(0), _ptr(gamma), eps, _ptr(norm_out), norm_out.stride(0))


def splitk_tail_workspace(rows: int, n: int, device: torch.device) -> torch.Tensor:
    """Zeroed ticket / row-statistics buffer of the fused split-K linear (scb_linear_splitk leaves it zero, so one buffer
    serves every launch of a stream-ordered chain with at most `rows` rows and `n` columns)."""
    return torch.zeros(L.load().scb_linear_splitk_workspace_bytes(rows, n) // 4, dtype=torch.int32, device=device)


def linear_splitk_fits_one_launch(rows: int, n: int, n_splits: int, device: torch.device) -> bool:
    """One resident CTA per (n tile, K slice): the fused reduction needs every CTA of the launch on the GPU at once."""
    return -(-rows // 128) * -(-n // 128) * n_splits <= torch.cuda.get_device_properties(device).multi_processor_count


def linear_splitk(a: torch.Tensor, w: torch.Tensor, n_splits: int, partial: torch.Tensor, out: torch.Tensor,
                  bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                  post_norm: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None,
                  tail_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = a @ w.T (+bias) (+residual) for few rows and a long K: the K range is cut into `n_splits`
    slices that run as a batched tcgen05 GEMM (one CTA wave streams every weight byte once, fp32 partials
    in `partial` [n_splits, rows, N]).  With `tail_ws` (splitk_tail_workspace) and one CTA per tile the last CTA of each
    n tile sums the slices in order and applies the epilogue — and the RMSNorm of the finished rows — inside the same
    launch (scb_linear_splitk); otherwise a second launch does (scb_splitk_reduce)."""
    dev = _require_cuda(a, w, partial, out, bias, residual)
    rows, k = a.shape
    n = w.shape[0]
    ks = k // n_splits
    if ks * n_splits != k or ks % 8:
        raise ValueError(f"k={k} cannot be cut into {n_splits} slices of a multiple of 8")
    if tail_ws is not None and linear_splitk_fits_one_launch(rows, n, n_splits, dev):
        if partial.numel() < n_splits * rows * n or partial.dtype != torch.float32 or not partial.is_contiguous():
            raise ValueError("linear_splitk: partial must be a contiguous fp32 buffer of n_splits*rows*n elements")
        pg = po = None
        pld, pe = 0, 0.0
        if post_norm is not None:
            pg, po, pe = post_norm
            pld = po.stride(0)
        args = L.LinearSplitKArgs(a=_ptr(a), w=_ptr(w), bias=_ptr(bias), residual=_ptr(residual), partial=_ptr(partial),
                                  out=_ptr(out), post_norm_gamma=_ptr(pg), post_norm_out=_ptr(po), workspace=_ptr(tail_ws),
                                  workspace_bytes=tail_ws.numel() * tail_ws.element_size(), m=rows, n=n, k=k,
                                  lda=a.stride(0), ldw=w.stride(0), ldr=residual.stride(0) if residual is not None else 0,
                                  ldo=out.stride(0), post_norm_ld=pld, n_splits=n_splits, post_norm_eps=pe)
        with torch.cuda.device(dev):
            L.check(L.load().scb_linear_splitk(C.byref(args), _stream(dev)), "scb_linear_splitk")
        return out
    args = L.LinearArgs(a=_ptr(a), w=_ptr(w), bias=None, residual=None, out=_ptr(partial), m=rows, n=n, k=ks,
                        lda=a.stride(0), ldw=w.stride(0), ldr=0, ldo=partial.stride(1), act=L.ACT_NONE,
                        out_dtype=L.SCB_F32, batch=n_splits, stride_a=ks, stride_w=ks, stride_r=0,
                        stride_o=partial.stride(0), alpha=1.0)
    with torch.cuda.device(dev):
        L.check(L.load().scb_linear(C.byref(args), _stream(dev)), "scb_linear")
    pg = po = None
    pld, pe = 0, 0.0
    if post_norm is not None:
        pg, po, pe = post_norm
        pld = po.stride(0)
    _call("scb_splitk_reduce", dev, _ptr(partial), partial.stride(0), partial.stride(1), n_splits, rows, n,
          _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0),
          _ptr(pg), _ptr(po), pld, pe)
    return out


def splitk_reduce(partial: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor] = None,
                  residual: Optional[torch.Tensor] = None,
                  post_norm: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None) -> torch.Tensor:
    """out[r, :] = fp16(sum_s partial[s, r, :] (+ bias) (+ residual[r, :])) summed in slice order, optionally followed by
    post_norm_out[r, :] = RMSNorm(out[r, :]) * gamma (scb_splitk_reduce).  partial [S, rows, N] fp32 — K-slices of a
    split-K GEMM or the ranks' partial sums of a row-parallel layer."""
    dev = _require_cuda(partial, out, bias, residual)
    if partial.dtype != torch.float32 or partial.dim() != 3 or partial.stride(2) != 1:
        raise TypeError("splitk_reduce: partial must be fp32 [S, rows, N] with contiguous rows")
    S, rows, n = partial.shape
    pg = po = None
    pld, pe = 0, 0.0
    if post_norm is not None:
        pg, po, pe = post_norm
        pld = po.stride(0)
    _call("scb_splitk_reduce", dev, _ptr(partial), partial.stride(0), partial.stride(1), S, rows, n, _ptr(bias),
          _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), _ptr(pg), _ptr(po),
          pld, pe)
    return out


def argmax_rows(logits: torch.Tensor, out: Optional[torch.Tensor] = None,
                counter: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _require_cuda(logits, out, counter)
    rows, n = logits.shape
    if out is None:
        out = torch.empty((rows,), dtype=torch.int64, device=dev)
    _call("scb_argmax_rows", dev, _ptr(logits), _dtype_code(logits), logits.stride(0), rows, n, _ptr(out),
          _ptr(counter), 0 if counter is None else counter.numel())
    return out


def gemv(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, prologue: int = L.PRO_NONE,
         gamma: Optional[torch.Tensor] = None, eps: float = 0.0, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.float16, argmax_out: Optional[torch.Tensor] = None,
         counters: Optional[torch.Tensor] = None, epilogue: int = L.EPI_NONE,
         post_norm: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None,
         prefetch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode-step linear layer for 1..8 activation rows (see scb_gemv).  `argmax_out` (int64
    [rows]) fuses the greedy token choice over the stored output; `counters` (int32) are then
    incremented by one.  `post_norm=(gamma, out_tensor, eps)` makes the last block write
    RMSNorm(out)*gamma into out_tensor; `prefetch` is a tensor (the next kernel's weights) pulled into
    L2 by blocks that finish early."""
    dev = _require_cuda(a, w, bias, residual, gamma, out, argmax_out, counters)
    rows = a.shape[0]
    n, k = w.shape
    if out is None:
        out = torch.empty((rows, n // 2 if epilogue == L.EPI_SILU_PAIR else n), dtype=out_dtype, device=dev)
    args = L.GemvArgs(a=_ptr(a), w=_ptr(w), bias=_ptr(bias), residual=_ptr(residual), gamma=_ptr(gamma),
                      out=_ptr(out), n=n, k=k, lda=a.stride(0), ldw=w.stride(0),
                      ldr=residual.stride(0) if residual is not None else 0, ldo=out.stride(0),
                      rows=rows, prologue=prologue, out_dtype=_dtype_code(out), eps=eps, epilogue=epilogue)
    key = (dev.index, "gemv_ws")
    ws = _workspaces.get(key)
    if ws is None:      # zero-initialised once: tickets are self-resetting
        ws = torch.zeros(L.load().scb_gemv_workspace_bytes(16), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    args.workspace = _ptr(ws)
    args.workspace_bytes = ws.numel()
    if post_norm is not None:
        pg, po, pe = post_norm
        args.post_norm_gamma, args.post_norm_out, args.post_norm_ld, args.post_norm_eps = _ptr(pg), _ptr(po), po.stride(0), pe
    if prefetch is not None:
        args.prefetch = _ptr(prefetch)
        args.prefetch_bytes = min(prefetch.numel() * prefetch.element_size(), 48 << 20)
    if argmax_out is not None:
        args.argmax_out = _ptr(argmax_out)
        args.counters = _ptr(counters)
        args.n_counters = 0 if counters is None else counters.numel()
    _call("scb_gemv", dev, C.byref(args))
    return out


def decode_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, scale: float,
                     kv_len: int = 0, kv_len_dev: Optional[torch.Tensor] = None, n_splits: int = 32,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [Hq, D] against caches [Hkv, max_pos, D]; context length from `kv_len_dev` (int32 device
    scalar, graph-replayable) or `kv_len`."""
    dev = _require_cuda(q, k_cache, v_cache, kv_len_dev, out)
    hq, d = q.shape
    hkv, max_pos, _ = k_cache.shape
    if out is None:
        out = torch.empty((hq, d), dtype=torch.float16, device=dev)
    lib = L.load()
    ws = _workspace(dev, "decode_attn", lib.scb_decode_attention_workspace_bytes(hq, d, n_splits))
    args = L.DecodeAttentionArgs(q=_ptr(q), k_cache=_ptr(k_cache), v_cache=_ptr(v_cache), out=_ptr(out),
                                 kv_len_dev=_ptr(kv_len_dev), max_pos=max_pos, kv_len=kv_len,
                                 n_q_heads=hq, n_kv_heads=hkv, head_dim=d, n_splits=n_splits, scale=scale,
                                 workspace=_ptr(ws), workspace_bytes=ws.numel())
    _call("scb_decode_attention", dev, C.byref(args))
    return out


def decode_attention_fused(qkv: torch.Tensor, inv_freq: Optional[torch.Tensor], theta: float,
                           k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len_dev: torch.Tensor,
                           n_q: int, scale: float, workspace: torch.Tensor, n_splits: int = 37,
                           out: Optional[torch.Tensor] = None, prefetch: Optional[torch.Tensor] = None,
                           ctx_cap: int = 0) -> torch.Tensor:
    """RoPE + KV append + split attention + combine for ONE new token of each of S sequences (see
    scb_decode_attention_fused).  qkv [S, (n_q+2*n_kv)*hd]; caches [n_kv, max_pos, hd] (S == 1) or
    [S, n_kv, max_pos, hd]; kv_len_dev int32 [S].  `workspace` is a caller-held uint8 tensor whose
    first 256 bytes are zero (use `decode_attention_fused_workspace`)."""
    dev = _require_cuda(qkv, inv_freq, k_cache, v_cache, kv_len_dev, workspace, out)
    n_seqs = qkv.shape[0]
    if k_cache.dim() == 3:
        hkv, max_pos, d = k_cache.shape
        seq_stride = 0
    else:
        _, hkv, max_pos, d = k_cache.shape
        seq_stride = k_cache.stride(0)
    if out is None:
        out = torch.empty((n_seqs, n_q, d), dtype=torch.float16, device=dev)
    args = L.DecodeAttentionFusedArgs(
        qkv=_ptr(qkv), inv_freq=_ptr(inv_freq), k_cache=_ptr(k_cache), v_cache=_ptr(v_cache), out=_ptr(out),
        kv_len_dev=_ptr(kv_len_dev), max_pos=max_pos, n_q_heads=n_q, n_kv_heads=hkv, head_dim=d,
        n_splits=n_splits, scale=scale, theta=theta, n_seqs=n_seqs, ld_qkv=qkv.stride(0),
        cache_seq_stride=seq_stride, workspace=_ptr(workspace), workspace_bytes=workspace.numel())
    args.ctx_cap = ctx_cap
    if prefetch is not None:
        args.prefetch = _ptr(prefetch)
        args.prefetch_bytes = min(prefetch.numel() * prefetch.element_size(), 48 << 20)
    _call("scb_decode_attention_fused", dev, C.byref(args))
    return out


def decode_attention_fused_workspace(dev, n_q: int, n_kv: int, head_dim: int, n_splits: int,
                                     n_seqs: int = 1) -> torch.Tensor:
    n = L.load().scb_decode_attention_fused_workspace_bytes(n_seqs, n_q, n_kv, head_dim, n_splits)
    if n == 0:
        raise ValueError("invalid decode_attention_fused configuration")
    return torch.zeros(n, dtype=torch.uint8, device=dev)


# ------------------------------------------------------------------------------------------
class Qwen2Decoder:
    """Handle of the persistent decode-step kernel (scb_qwen2_decoder_*): one launch per cached decode
    step for 1..max_rows sequences.  Holds references to every tensor whose pointer the handle keeps."""

    def __init__(self, *, embed: torch.Tensor, final_norm: torch.Tensor, lm_head: torch.Tensor, layers: list,
                 k_cache: list, v_cache: list, n_q: int, n_kv: int, head_dim: int, inter: int, eps: float,
                 theta: float, inv_freq: Optional[torch.Tensor], max_ctx: int):
        dev = _require_cuda(embed, final_norm, lm_head)
        n = len(layers)
        self._keep = (embed, final_norm, lm_head, layers, k_cache, v_cache, inv_freq)

        def table(ts):
            arr = (C.c_void_p * n)()
            for i, t in enumerate(ts):
                if t is not None:
                    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float16):
                        raise L.ScbError("Qwen2Decoder: weights / caches must be contiguous fp16 CUDA tensors")
                    arr[i] = t.data_ptr()
            return arr

        self._tables = [table([ly.get(k) for ly in layers]) for k in ("qkv_w", "qkv_b", "o_w", "gu_w", "down_w", "ln1", "ln2")]
        self._tables += [table(k_cache), table(v_cache)]
        d = L.Qwen2DecoderDesc()
        d.n_layers, d.hidden, d.n_q_heads, d.n_kv_heads = n, embed.shape[1], n_q, n_kv
        d.head_dim, d.inter, d.vocab, d.max_seqs = head_dim, inter, lm_head.shape[0], k_cache[0].shape[0]
        d.max_ctx, d.rms_eps, d.rope_theta = max_ctx, eps, theta
        d.inv_freq = _ptr(inv_freq)
        d.embed, d.final_norm, d.lm_head = embed.data_ptr(), final_norm.data_ptr(), lm_head.data_ptr()
        (d.qkv_w, d.qkv_b, d.o_w, d.gu_w, d.down_w, d.ln1, d.ln2, d.k_cache, d.v_cache) = [
            C.cast(t, C.POINTER(C.c_void_p)) for t in self._tables]
        self.device = dev
        self.vocab = lm_head.shape[0]
        self.max_rows = int(L.load().scb_qwen2_decoder_max_rows())
        h = C.c_void_p()
        with torch.cuda.device(dev):
            L.check(L.load().scb_qwen2_decoder_create(C.byref(d), C.byref(h)), "scb_qwen2_decoder_create")
        self._h = h

    def step(self, rows: int, next_id: torch.Tensor, kv_len: torch.Tensor, logits: torch.Tensor,
             counters: Optional[torch.Tensor] = None, ctx_cap: int = 0, ctx_sum: int = 0) -> None:
        """One decode step for cache slots 0..rows-1: consumes next_id[:rows] (int64), writes the greedy
        next tokens back into it, fp16 logits into logits[:rows], and adds 1 to every int32 of `counters`."""
        _require_cuda(next_id, kv_len, logits, counters)
        if next_id.dtype != torch.int64 or kv_len.dtype != torch.int32 or logits.dtype != torch.float16:
            raise TypeError("Qwen2Decoder.step: next_id int64, kv_len int32, logits fp16")
        a = L.Qwen2DecodeStepArgs()
        a.decoder, a.rows = self._h, rows
        a.next_id, a.kv_len, a.logits, a.ld_logits = next_id.data_ptr(), kv_len.data_ptr(), logits.data_ptr(), logits.stride(0)
        a.counters, a.n_counters = _ptr(counters), (counters.numel() if counters is not None else 0)
        a.ctx_cap = ctx_cap
        a.ctx_sum = ctx_sum          # upper bound of sum(kv_len) over the rows (0: rows * ctx_cap)
        with torch.cuda.device(self.device):
            L.check(L.load().scb_qwen2_decode_step(C.byref(a), _stream(self.device)), "scb_qwen2_decode_step")

    def prefill(self, inputs_embeds: torch.Tensor, slot: int, pos0: int, logits_row: torch.Tensor) -> None:
        """scb_qwen2_prefill: the whole prefill of one sequence as one native call (see the header)."""
        _require_cuda(inputs_embeds, logits_row)
        if inputs_embeds.dtype != torch.float16 or not inputs_embeds.is_contiguous() or logits_row.dtype != torch.float16:
            raise TypeError("Qwen2Decoder.prefill: contiguous fp16 inputs_embeds, fp16 logits row")
        lib = L.load()
        T = inputs_embeds.shape[0]
        need = int(lib.scb_qwen2_prefill_workspace_bytes(self._h, T))
        ws = getattr(self, "_prefill_ws", None)
        if ws is None or ws.numel() < need:
            ws = torch.zeros(need, dtype=torch.uint8, device=self.device)     # zero: the lm_head GEMV's ticket
            self._prefill_ws = ws
        with torch.cuda.device(self.device):
            L.check(lib.scb_qwen2_prefill(self._h, inputs_embeds.data_ptr(), T, slot, pos0, logits_row.data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream(self.device)), "scb_qwen2_prefill")

    def prefill_batch(self, stacked: torch.Tensor, n_tokens: list, slots: list, pos0: list, logits: torch.Tensor) -> None:
        """scb_qwen2_prefill_batch: several prompts (rows stacked in `stacked` [sum(n_tokens), H]) in ONE pass; sequence q's
        last-position logits land in logits[q]."""
        _require_cuda(stacked, logits)
        if stacked.dtype != torch.float16 or not stacked.is_contiguous() or logits.dtype != torch.float16:
            raise TypeError("Qwen2Decoder.prefill_batch: contiguous fp16 rows, fp16 logits")
        lib = L.load()
        n = len(n_tokens)
        if stacked.shape[0] != sum(n_tokens) or len(slots) != n or len(pos0) != n or logits.shape[0] < n:
            raise ValueError("prefill_batch: inconsistent sequence lists")
        need = int(lib.scb_qwen2_prefill_workspace_bytes(self._h, int(stacked.shape[0])))
        ws = getattr(self, "_prefill_ws", None)
        if ws is None or ws.numel() < need:
            self._prefill_ws = None
            ws = torch.zeros(need, dtype=torch.uint8, device=self.device)     # zero: the lm_head GEMV's ticket
            self._prefill_ws = ws
        nt = (C.c_int64 * n)(*[int(t) for t in n_tokens])
        sl = (C.c_int32 * n)(*[int(t) for t in slots])
        p0 = (C.c_int64 * n)(*[int(t) for t in pos0])
        with torch.cuda.device(self.device):
            L.check(lib.scb_qwen2_prefill_batch(self._h, stacked.data_ptr(), n, nt, sl, p0, logits.data_ptr(), logits.stride(0),
                                                ws.data_ptr(), ws.numel(), _stream(self.device)), "scb_qwen2_prefill_batch")

    def close(self) -> None:
        if getattr(self, "_h", None):
            L.load().scb_qwen2_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VitEncoder:
    """Handle of the native `encode_images` entry point (scb_vit_encoder_* / scb_vit_encode): the whole CLIP tower ->
    projector -> 2x2 pool pass as one C call into a reused workspace.  Holds references to every weight tensor."""

    LAYER_KEYS = ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_g", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")

    def __init__(self, *, layers: list, patch_w: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, pre_g: torch.Tensor,
                 pre_b: torch.Tensor, proj1_w: torch.Tensor, proj1_b: torch.Tensor, proj2_w: torch.Tensor,
                 proj2_b: torch.Tensor, n_heads: int, image_size: int, patch: int, channels: int, ln_eps: float):
        dev = _require_cuda(patch_w, cls, pos, proj1_w, proj2_w)
        n = len(layers)
        self._keep = (layers, patch_w, cls, pos, pre_g, pre_b, proj1_w, proj1_b, proj2_w, proj2_b)

        def ok(t):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float16):
                raise L.ScbError("VitEncoder: weights must be contiguous fp16 CUDA tensors")
            return t.data_ptr()

        self._tables = []
        for key in self.LAYER_KEYS:
            arr = (C.c_void_p * max(n, 1))()
            for i, ly in enumerate(layers):
                arr[i] = ok(ly[key])
            self._tables.append(arr)
        d = L.VitEncoderDesc()
        d.n_layers, d.hidden, d.n_heads = n, patch_w.shape[0], n_heads
        d.inter = layers[0]["fc1_w"].shape[0] if n else 8
        d.image_size, d.patch, d.channels, d.kpad = image_size, patch, channels, patch_w.shape[1]
        d.out_dim, d.ln_eps = proj2_w.shape[0], ln_eps
        d.patch_w, d.cls, d.pos, d.pre_g, d.pre_b = ok(patch_w), ok(cls), ok(pos), ok(pre_g), ok(pre_b)
        d.proj1_w, d.proj1_b, d.proj2_w, d.proj2_b = ok(proj1_w), ok(proj1_b), ok(proj2_w), ok(proj2_b)
        for key, arr in zip(self.LAYER_KEYS, self._tables):
            setattr(d, key, C.cast(arr, C.POINTER(C.c_void_p)))
        self.device = dev
        self.tokens = (image_size // patch) ** 2 + 1
        self.pooled = (image_size // patch // 2) ** 2
        self.hidden, self.out_dim = patch_w.shape[0], proj2_w.shape[0]
        self._ws: Optional[torch.Tensor] = None
        h = C.c_void_p()
        L.check(L.load().scb_vit_encoder_create(C.byref(d), C.byref(h)), "scb_vit_encoder_create")
        self._h = h

    def encode(self, pixels: torch.Tensor, want_tower: bool = False):
        """pixels [n, C, S, S] fp16 contiguous -> features [n, pooled, out_dim] fp16 (and the tower's hidden state
        [n * tokens, hidden] when asked)."""
        _require_cuda(pixels)
        if pixels.dtype != torch.float16 or not pixels.is_contiguous() or pixels.dim() != 4:
            raise TypeError("VitEncoder.encode: contiguous fp16 [n, C, S, S] pixels")
        lib = L.load()
        n = pixels.shape[0]
        need = int(lib.scb_vit_encode_workspace_bytes(self._h, n))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None                           # release before growing
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty((n, self.pooled, self.out_dim), dtype=torch.float16, device=self.device)
        tower = torch.empty((n * self.tokens, self.hidden), dtype=torch.float16, device=self.device) if want_tower else None
        with torch.cuda.device(self.device):
            L.check(lib.scb_vit_encode(self._h, pixels.data_ptr(), n, out.data_ptr(), _ptr(tower), self._ws.data_ptr(),
                                       self._ws.numel(), _stream(self.device)), "scb_vit_encode")
        return (out, tower) if want_tower else out

    def close(self) -> None:
        if getattr(self, "_h", None):
            L.load().scb_vit_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _RawCudaBuffer:
    """`__cuda_array_interface__` view of library-allocated device memory (lets torch wrap it without a copy)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """Gather buffer of the frame-sharded stream, shared over NVLink peer memory (scb_peer_* / scb_allgather_embed).

    `buffer` is a torch view [world, *slice_shape] of library-allocated memory: rank r fills buffer[r]; `allgather()`
    pushes it into every peer's buffer and returns (asynchronously, on the current stream) with every slice valid.
    The 64-byte IPC handles travel over `torch.distributed` (control plane).  Construction is a collective: every
    rank of `group` must call it; it raises on every rank if ANY rank cannot map its peers (the caller then lets NCCL
    move the bytes)."""

    def __init__(self, shape, dtype: torch.dtype, rank: int, world: int, device, group=None, device_epoch: bool = False):
        """`device_epoch`: the call number of `allgather()` lives on the device, so the launch can be captured in a CUDA
        graph and replayed (the per-layer exchanges of tp_decode.TensorParallelDecoder)."""
        import torch.distributed as dist
        self.device_epoch = device_epoch
        if shape[0] != world:
            raise ValueError("PeerGather: shape[0] must be the world size")
        self.rank, self.world, self.device = rank, world, torch.device(device)
        lib = L.load()
        elem = torch.empty((), dtype=dtype).element_size()
        n = 1
        for d in shape[1:]:
            n *= int(d)
        self.slice_bytes = n * elem
        total = int(lib.scb_allgather_embed_buffer_bytes(self.slice_bytes, world))
        if total == 0:
            raise L.ScbError(f"PeerGather: slice of {self.slice_bytes} B x {world} ranks is not a valid configuration")
        self._ptr = C.c_void_p()
        self._mapped = {}
        handle = (C.c_ubyte * 64)()
        ok, err = 1, ""
        with torch.cuda.device(self.device):
            try:
                L.check(lib.scb_peer_alloc(total, C.byref(self._ptr), handle), "scb_peer_alloc")
            except L.ScbError as e:
                ok, err = 0, str(e)
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
            if ok and all(h is not None for h in handles):
                for r, h in enumerate(handles):
                    if r == rank:
                        continue
                    p = C.c_void_p()
                    rc = lib.scb_peer_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p))
                    if rc != 0:
                        ok, err = 0, lib.scb_last_error().decode("utf-8", "replace")
                        break
                    self._mapped[r] = p
            else:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.close()
                raise L.ScbError(f"PeerGather: peer memory unavailable on some rank ({err or 'a peer failed'})")
        self._total = total
        raw = torch.as_tensor(_RawCudaBuffer(self._ptr.value, self.slice_bytes * world, self), device=self.device)
        self.buffer = raw.view(dtype).view(*shape)
        self._table = (C.c_void_p * world)()
        for r in range(world):
            self._table[r] = self._ptr.value if r == rank else self._mapped[r].value
        self.epoch = 0

    def allgather(self) -> None:
        self.epoch += 1
        a = L.AllgatherEmbedArgs(peer_bufs=C.cast(self._table, C.POINTER(C.c_void_p)), rank=self.rank, world=self.world,
                                 slice_bytes=self.slice_bytes, epoch=0 if self.device_epoch else self.epoch)
        with torch.cuda.device(self.device):
            L.check(L.load().scb_allgather_embed(C.byref(a), _stream(self.device)), "scb_allgather_embed")

    def close(self) -> None:
        lib = L.load()
        with torch.cuda.device(self.device):
            for p in self._mapped.values():
                lib.scb_peer_close(p)
            self._mapped = {}
            if self._ptr:
                torch.cuda.synchronize(self.device)
                lib.scb_peer_free(self._ptr)
                self._ptr = C.c_void_p()
