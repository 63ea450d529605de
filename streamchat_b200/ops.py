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
    key = (dev.index, name)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
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
    oshape = (bsz, m, n) if batched else (m, n)
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
        out = out.view(*lead_shape, n)
    return out
