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

    def __init__(self, init_method: str = "rnd", num_init: int = 8, max_iter: int = 100,
                 distance=LpDistance, p_norm: int = 2, tol: float = 1e-4,
                 normalize: Optional[Union[str, bool]] = None, n_clusters: Optional[int] = 8,
                 verbose: bool = True, seed: Optional[int] = 123, **kwargs):
        super().__init__()
        self.init_method = init_method.lower()
        self.num_init, self.max_iter, self.p_norm, self.tol = num_init, max_iter, p_norm, tol
        self.normalize = ("mean" if normalize else None) if isinstance(normalize, bool) else normalize
        self.n_clusters, self.verbose, self.seed = n_clusters, verbose, seed
        self.distance = distance(**kwargs) if isinstance(distance, type) else distance
        self._result = None
        self.eps = None
        self._k_max = None
        # argument contract of the upstream constructor (kmeans.py:108-134), one check per argument
        bad = [("init_method", self.init_method not in ("rnd", "k-means++")),
               ("num_init", num_init <= 0), ("max_iter", max_iter <= 0), ("p_norm", p_norm <= 0),
               ("tol", tol is not None and not 0 <= tol <= 1),
               ("normalize", self.normalize not in (None, "mean", "minmax", "unit")),
               ("n_clusters", n_clusters is not None and n_clusters < 2)]
        for name, wrong in bad:
            if wrong:
                raise ValueError(f"KMeans: invalid value for {name}: {getattr(self, name)!r}")
        if not isinstance(self.distance, LpDistance) or self.p_norm != 2:
            raise NotImplementedError("only LpDistance with p_norm=2 is implemented on sm_100a")

    @property
    def is_fitted(self) -> bool:
        return self._result is not None

    @property
    def num_clusters(self) -> Union[int, Tensor, Any]:
        return self._result.k if self.is_fitted else None

    def _check_x(self, x) -> Tensor:
        """x must be a [BS, N, D] tensor (extra singleton dims are squeezed away, kmeans.py:136-151)."""
        if not isinstance(x, Tensor):
            raise TypeError(f"x must be a torch.Tensor, got {type(x)}")
        if x.dim() > 3:
            x = x.squeeze()
        if x.dim() != 3:
            raise ValueError(f"x must have shape (BS, N, D), got {tuple(x.shape)}")
        self.eps = torch.finfo(x.dtype).eps
        return x

    def _check_k(self, k, dims, device=torch.device("cpu")) -> LongTensor:
        """Number of clusters per instance as a [BS] long tensor (kmeans.py:153-185); the kernel takes one common k."""
        bs, n, _ = dims
        if k is None:
            k = self.n_clusters
            if k is None:
                raise ValueError("no k given and no default n_clusters set")
        if isinstance(k, int):
            k = torch.full((bs,), k, dtype=torch.long)
        elif not isinstance(k, Tensor):
            raise TypeError(f"k must be an int, a tensor or None, got {type(k)}")
        k = k.reshape(-1)
        if k.numel() == 1:
            k = k.repeat(bs)
        if bool((k >= n).any()) or bool((k <= 1).any()):
            raise ValueError(f"k must satisfy 1 < k < n={n}, got {k.tolist()}")
        if int(k.min()) != int(k.max()):
            raise NotImplementedError("per-instance k is not implemented on sm_100a (one common k per batch)")
        self._k_max = int(k.max())
        return k.to(dtype=torch.long, device=device)

    def _check_centers(self, centers, dims, dtype=torch.float32, device=torch.device("cpu")) -> Tensor:
        """Initial centers as [BS, num_init, K, D]; a [BS, K, D] tensor is shared by all restarts (kmeans.py:187-241)."""
        if not isinstance(centers, Tensor):
            raise TypeError(f"centers must be a torch.Tensor, got {type(centers)}")
        bs, _, d = dims
        if centers.dim() == 3:
            if self.num_init > 1:
                warn(f"one center configuration per instance given: all {self.num_init} restarts start from it")
            centers = centers.unsqueeze(1).expand(-1, self.num_init, -1, -1)
        want = (bs, self.num_init, self._k_max, d)
        if tuple(centers.shape) != want:
            raise ValueError(f"centers must have shape {want} (or without the restart dim), got {tuple(centers.shape)}")
        return centers.contiguous().to(dtype=dtype, device=device)

    @staticmethod
    def _normalize(x: Tensor, normalize: str, eps: float = 1e-8):
        """Input conditioning of kmeans.py:363-384 (torch element-wise ops on the caller's tensor, off the hot loop;
        'mean' / 'minmax' modify x in place as upstream does)."""
        if normalize == "mean":
            x -= x.mean(dim=1, keepdim=True)
            return x
        if normalize == "minmax":
            x -= x.min(-1, keepdim=True).values
            x /= x.max(-1, keepdim=True).values
            return x
        if normalize == "unit":
            x = torch.where(x == 0, torch.full_like(x, eps), x)
            return x / torch.norm(x, p=2, dim=-1, keepdim=True)
        raise ValueError(f"unknown normalization {normalize!r}")

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
