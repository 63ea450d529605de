"""Host-side mirrors of the reference's two k-means entry points, running on the sm_100a
`scb_kmeans_step` kernels (no CPU fallback).

  weighted_kmeans_feature  <- utiles.py:291-330   (selective frame stacking / tree merge, a-S3)
  KMeans, ClusterResult    <- torch_kmeans/clustering/kmeans.py:24-644, utils/utils.py:11-31 (a-S4)

Same names, argument order and return arity as the reference.  Deviations, all stated in
DESIGN.md: arithmetic is fp32 with fp16/fp32 I/O (the reference's fp16 distance sum overflows,
hazard H3), and every random draw can be injected (`init_indices`, `repair_indices`, `centers`)
so a CPU oracle can be driven with identical choices (hazard H1).
"""
from __future__ import annotations

import random
from typing import Any, NamedTuple, Optional, Sequence, Union
from warnings import warn

import torch
from torch import LongTensor, Tensor

from . import ops

__all__ = ["weighted_kmeans_feature", "KMeans", "ClusterResult", "LpDistance"]


def weighted_kmeans_feature(img_feature: Tensor, video_max_frames: int, weights: Optional[Tensor] = None,
                            *, init_indices: Optional[Sequence[int]] = None,
                            repair_indices: Optional[Sequence[int]] = None,
                            tol: float = 1e-4, max_iter: int = 10):
    """Cluster T frames (each flattened to P*D dims) into `video_max_frames` centroids.

    Returns `(img_feature, weights, [[[i]...]])` when T <= K (utiles.py:321-322) and
    `(reduced_feature [K,P,D], labels [T])` otherwise (utiles.py:330) — the reference's arity quirk
    is kept on purpose.
    """
    if weights is None:
        weights = torch.ones(img_feature.size(0), dtype=img_feature.dtype, device=img_feature.device)
    T, P, D = img_feature.shape
    K = video_max_frames
    if T <= K:
        return img_feature, weights, [[[i] for i in range(T)]]
    X = img_feature.reshape(T, -1)
    if X.dtype not in (torch.float16, torch.float32):
        X = X.float()
    if init_indices is None:
        indices = torch.randperm(T, device=X.device)[:K]                    # utiles.py:295
    else:
        indices = torch.as_tensor(list(init_indices), dtype=torch.long, device=X.device)
        if indices.numel() != K:
            raise ValueError(f"init_indices must hold {K} indices")
    repair = list(repair_indices) if repair_indices is not None else None
    centroids = X[indices].float().unsqueeze(0).contiguous()               # [1,K,F]
    w32 = weights.float()
    labels = None
    for _ in range(max_iter):
        r = ops.kmeans_step(X, centroids, w32, dist_eps=0.0, do_update=True)
        labels = r["labels"][0]
        new_c = r["centroids"]
        wsum = r["wsum"][0]
        shift = r["shift"][0]
        empty = (wsum <= 0).nonzero().flatten().tolist()                   # host sync, as the reference's
        if empty:                                                          # `if mask.sum() < num_clusters`
            for j in empty:                                                # utiles.py:312-313
                ridx = repair.pop(0) if repair else random.randint(0, T - 1)
                new_c[0, j] = X[ridx].float()
                d = ops.kmeans_step(X[ridx:ridx + 1], centroids[:, j:j + 1], None, do_update=False,
                                    want_dist=True)["dist"]
                shift[j] = d[0, 0]
        diff = float(shift.sum())                                          # utiles.py:314
        if diff < tol:
            break                                                          # keeps the pre-update centroids
        centroids = new_c
    reduced = centroids[0].to(img_feature.dtype).view(K, P, D)
    return reduced, labels.long()


# ------------------------------------------------------------------------------------------
class ClusterResult(NamedTuple):
    """torch_kmeans/utils/utils.py:11-31."""
    labels: LongTensor
    centers: Tensor
    inertia: Tensor
    x_org: Tensor
    x_norm: Tensor
    k: LongTensor
    soft_assignment: Optional[Tensor] = None


class LpDistance:
    """Marker for the only distance the sm_100a kernel implements (p=2, torch_kmeans/utils/distances.py:109)."""

    def __init__(self, p: int = 2, **kwargs):
        if p != 2:
            raise NotImplementedError("only the Euclidean (p=2) LpDistance runs on the sm_100a kernel")
        self.p = p


class KMeans(torch.nn.Module):
    """Batched Lloyd k-means with `num_init` restarts (torch_kmeans/clustering/kmeans.py:24-644).

    Supported: init_method 'rnd' (same torch.multinomial draw as the reference, kmeans.py:386-419)
    or explicit `centers=`; LpDistance p=2; normalize None/'mean'/'minmax'/'unit'; one common k per
    batch.  Not supported (raise): k-means++ init, per-instance k, other distances.
    """

    INIT_METHODS = ["rnd", "k-means++"]
    NORM_METHODS = ["mean", "minmax", "unit"]

    def __init__(self, init_method: str = "rnd", num_init: int = 8, max_iter: int = 100,
                 distance=LpDistance, p_norm: int = 2, tol: float = 1e-4,
                 normalize: Optional[Union[str, bool]] = None, n_clusters: Optional[int] = 8,
                 verbose: bool = True, seed: Optional[int] = 123, **kwargs):
        super().__init__()
        self.init_method = init_method.lower()
        self.num_init = num_init
        self.max_iter = max_iter
        self.p_norm = p_norm
        self.tol = tol
        self.normalize = normalize
        self.n_clusters = n_clusters
        self.verbose = verbose
        self.seed = seed
        self.distance = distance(**kwargs) if isinstance(distance, type) else distance
        self._result = None
        self.eps = None
        self._k_max = None
        # kmeans.py:108-134
        if self.init_method not in self.INIT_METHODS:
            raise ValueError(f"unknown <init_method>: {self.init_method}. Please choose one of {self.INIT_METHODS}")
        if self.num_init <= 0:
            raise ValueError(f"num_init should be > 0, but got {self.num_init}.")
        if self.max_iter <= 0:
            raise ValueError(f"max_iter should be > 0, but got {self.max_iter}.")
        if self.p_norm <= 0:
            raise ValueError(f"p_norm should be > 0, but got {self.p_norm}.")
        if self.tol is not None and (self.tol < 0 or self.tol > 1):
            raise ValueError(f"tol should be > 0 and < 1, but got {self.tol}.")
        if isinstance(self.normalize, bool):
            self.normalize = "mean" if self.normalize else None
        if self.normalize is not None and self.normalize not in self.NORM_METHODS:
            raise ValueError(f"unknown <normalize> method: {self.normalize}. Please choose one of {self.NORM_METHODS}")
        if self.n_clusters is not None and self.n_clusters < 2:
            raise ValueError(f"n_clusters should be > 1, but got {self.n_clusters}.")
        if not isinstance(self.distance, LpDistance) or self.p_norm != 2:
            raise NotImplementedError("only LpDistance with p_norm=2 is implemented on sm_100a")

    @property
    def is_fitted(self) -> bool:
        return self._result is not None

    @property
    def num_clusters(self) -> Union[int, Tensor, Any]:
        if not self.is_fitted:
            return None
        return self._result.k

    # kmeans.py:136-151
    def _check_x(self, x) -> Tensor:
        if not isinstance(x, Tensor):
            raise TypeError(f"x has to be a torch.Tensor but got {type(x)}.")
        shp = x.shape
        if len(shp) < 3:
            raise ValueError(f"input <x> should be at least of shape (BS, N, D) with batch size BS, "
                             f"number of points N and number of dimensions D but got {shp}.")
        elif len(shp) > 3:
            x = self._check_x(x.squeeze())
        self.eps = torch.finfo(x.dtype).eps
        return x

    # kmeans.py:153-185
    def _check_k(self, k, dims, device=torch.device("cpu")) -> LongTensor:
        bs, n, d = dims
        if not isinstance(k, Tensor):
            if k is None:
                if self.n_clusters is None:
                    raise ValueError("Did not provide number of clusters k on call and "
                                     "did not specify default 'n_clusters' at initialization.")
                k = self.n_clusters
            if isinstance(k, int):
                k = torch.empty(bs, dtype=torch.long).fill_(k)
            else:
                raise TypeError(f"k has to be int, torch.Tensor or None but got {type(k)}.")
        if len(k.shape) > 1:
            k = k.squeeze()
            assert len(k.shape) == 1
        if k.shape[0] == 1:
            k = k.repeat(bs)
        if (k >= n).any():
            raise ValueError(f"Specified 'k' must be smaller than number of samples n={n}, but got: {k}.")
        if (k <= 1).any():
            raise ValueError("Clustering for k=1 is ambiguous.")
        if int(k.min()) != int(k.max()):
            raise NotImplementedError("per-instance k is not implemented on sm_100a (one common k per batch)")
        self._k_max = int(k.max())
        return k.to(dtype=torch.long, device=device)

    # kmeans.py:187-241
    def _check_centers(self, centers, dims, dtype=torch.float32, device=torch.device("cpu")) -> Tensor:
        if not isinstance(centers, Tensor):
            raise TypeError(f"centers has to be a torch.Tensor but got {type(centers)}.")
        bs, n, d = dims
        if len(centers.shape) == 3:
            if centers.size(0) != bs or centers.size(1) != self._k_max or centers.size(2) != d:
                raise ValueError(f"centers needs to be of shape ({bs}, {self._k_max}, {d}),"
                                 f"but got {tuple(centers.shape)}.")
            if self.num_init > 1:
                warn(f"Specified num_init={self.num_init} > 1 but provided only 1 center configuration "
                     f"per instance. Using same center configuration for all {self.num_init} runs.")
                centers = centers[:, None, :, :].expand(centers.size(0), self.num_init,
                                                        centers.size(1), centers.size(2))
            else:
                centers = centers.unsqueeze(1)
        elif len(centers.shape) == 4:
            if (centers.size(0) != bs or centers.size(1) != self.num_init
                    or centers.size(2) != self._k_max or centers.size(3) != d):
                raise ValueError(f"centers needs to be of shape ({bs}, {self.num_init}, {self._k_max}, {d}),"
                                 f"but got {tuple(centers.shape)}.")
        else:
            raise ValueError(f"centers have unsupported shape of {tuple(centers.shape)} instead of "
                             f"({bs}, {self.num_init}, {self._k_max}, {d}).")
        return centers.contiguous().to(dtype=dtype, device=device)

    @staticmethod
    def _normalize(x: Tensor, normalize: str, eps: float = 1e-8):
        # kmeans.py:363-384 (input conditioning, element-wise; in place for mean/minmax as upstream)
        if normalize == "mean":
            x -= x.mean(dim=1)[:, None, :]
        elif normalize == "minmax":
            x -= x.min(-1, keepdims=True).values
            x /= x.max(-1, keepdims=True).values
        elif normalize == "unit":
            z_msk = x == 0
            x = x.clone()
            x[z_msk] = eps
            x = torch.diag_embed(1.0 / (torch.norm(x, p=2, dim=-1))) @ x
        else:
            raise ValueError(f"unknown normalization type {normalize}.")
        return x

    def _init_rnd(self, x: Tensor, k: LongTensor) -> Tensor:
        # kmeans.py:386-419 — identical draw (same generator seed, same multinomial call)
        bs, n, d = x.size()
        k_max = int(torch.max(k))
        gen = None
        if self.seed is not None:
            gen = torch.Generator(device=x.device)
            gen.manual_seed(self.seed)
        rnd_idx = torch.multinomial(
            torch.empty((bs * self.num_init, n), device=x.device, dtype=x.dtype).fill_(1 / n),
            num_samples=k_max, replacement=False, generator=gen)
        return x.gather(index=rnd_idx.view(bs, -1)[:, :, None].expand(bs, -1, d), dim=1).view(
            bs, self.num_init, k_max, d)

    def _center_init(self, x: Tensor, k: LongTensor, **kwargs) -> Tensor:
        if self.init_method == "rnd":
            return self._init_rnd(x, k)
        raise NotImplementedError("init_method 'k-means++' is not implemented on sm_100a; pass centers=")

    @torch.no_grad()
    def _cluster(self, x: Tensor, centers: Tensor, k: LongTensor, **kwargs):
        """kmeans.py:517-570 on the fused assign/update kernel; eps=1e-6 reproduces
        F.pairwise_distance (distances.py:134-141)."""
        bs, n, d = x.shape
        xs = [x[b].contiguous() for b in range(bs)]
        cur = [centers[b].float().contiguous() for b in range(bs)]       # [num_init, K, D] each
        for i in range(self.max_iter):
            shifts = []
            for b in range(bs):
                r = ops.kmeans_step(xs[b], cur[b], None, dist_eps=1e-6, do_update=True)
                cur[b] = r["centroids"]
                shifts.append(r["shift"])                                # [num_init, K]
            if self.tol is not None:
                shift = torch.stack(shifts).mean(dim=-1)                 # kmeans.py:601-607
                if bool((shift < self.tol).all()):
                    if self.verbose:
                        print(f"Full batch converged at iteration {i + 1}/{self.max_iter} "
                              f"with center shifts = {shift.view(-1, self.num_init).mean(-1)}.")
                    break
        labels, out_c, inertia = [], [], []
        for b in range(bs):
            r = ops.kmeans_step(xs[b], cur[b], None, dist_eps=1e-6, do_update=False)
            r0 = ops.kmeans_step(xs[b], cur[b], None, dist_eps=0.0, do_update=False, want_dist=True)
            inert = (r0["dist"] ** 2).sum(-1)                             # kmeans.py:611-632
            best = int(torch.argmin(inert))
            labels.append(r["labels"][best].long())
            out_c.append(cur[b][best].to(x.dtype))
            inertia.append(inert[best].to(x.dtype))
        return torch.stack(labels), torch.stack(out_c), torch.stack(inertia), None

    def forward(self, x: Tensor, k=None, centers: Optional[Tensor] = None, **kwargs) -> ClusterResult:
        x = self._check_x(x)
        x_ = x
        k = self._check_k(k, dims=x.shape, device=x.device)
        if self.normalize is not None:
            x = self._normalize(x, self.normalize, self.eps)
        if centers is None:
            centers = self._center_init(x, k, **kwargs)
        centers = self._check_centers(centers, dims=x.shape, dtype=x.dtype, device=x.device)
        labels, new_centers, inertia, soft_assign = self._cluster(x, centers, k, **kwargs)
        return ClusterResult(labels=labels, centers=new_centers, inertia=inertia, x_org=x_, x_norm=x,
                             k=k, soft_assignment=soft_assign)

    def fit(self, x: Tensor, k=None, centers: Optional[Tensor] = None, **kwargs):
        self._result = self(x, k=k, centers=centers, **kwargs)
        return self

    def predict(self, x: Tensor, **kwargs) -> LongTensor:
        assert self.is_fitted
        x = self._check_x(x)
        out = []
        for b in range(x.shape[0]):
            r = ops.kmeans_step(x[b].contiguous(), self._result.centers[b:b + 1].float().contiguous(),
                                None, dist_eps=1e-6, do_update=False)
            out.append(r["labels"][0].long())
        return torch.stack(out)

    def fit_predict(self, x: Tensor, k=None, centers: Optional[Tensor] = None, **kwargs) -> LongTensor:
        return self(x, k=k, centers=centers, **kwargs).labels

    def __repr__(self):
        return (f"{self.__class__.__name__}(init: '{self.init_method}', num_init: {self.num_init}, "
                f"max_iter: {self.max_iter}, distance: {self.distance}, tolerance: {self.tol}, "
                f"normalize: {self.normalize})")
