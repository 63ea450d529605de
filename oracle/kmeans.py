"""CPU oracle for the two k-means variants on the path (TEST INFRASTRUCTURE).

weighted_kmeans     restates weighted_kmeans_feature / weighted_kmeans_torch, utiles.py:291-330,
                    in fp32 (the reference runs it in fp16 where the 516 096-d squared-distance
                    sum overflows to inf — stated deviation H3) and with the two RNG draws
                    (torch.randperm :295, random.randint :313) injected (hazard H1).
torch_kmeans_forward restates torch_kmeans.KMeans.forward with given centers,
                    torch_kmeans/clustering/kmeans.py:517-632 + utils/utils.py:33-66 +
                    utils/distances.py:134-141 (F.pairwise_distance adds eps=1e-6 to the
                    difference before the norm).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


def weighted_kmeans(x: torch.Tensor, num_clusters: int, init_indices: Sequence[int],
                    weights: Optional[torch.Tensor] = None, tol: float = 1e-4, max_iter: int = 10,
                    repair_index: Optional[Callable[[], int]] = None):
    """x [T, F] -> (centroids [K, F] fp32, labels [T] int64, weights_sum [K], exit_iter)."""
    X = x.float()
    T = X.size(0)
    if weights is None:
        weights = torch.ones(T, dtype=torch.float32)
    weights = weights.float()
    centroids = X[torch.as_tensor(list(init_indices), dtype=torch.long)]          # :295-296
    labels = None
    weights_sum = None
    i = 0
    for i in range(max_iter):                                                     # :297
        dists = ((X.unsqueeze(1) - centroids.unsqueeze(0)) ** 2).sum(dim=2).sqrt()  # :299
        labels = torch.argmin(dists, dim=1)                                       # :302
        weighted_sum = torch.zeros_like(centroids)
        weights_sum = torch.zeros(num_clusters, dtype=torch.float32)
        for j in range(num_clusters):                                             # :305-308
            m = labels == j
            weighted_sum[j] = torch.sum(weights[m, None] * X[m], dim=0)
            weights_sum[j] = torch.sum(weights[m])
        mask = weights_sum > 0
        new_centroids = torch.zeros_like(weighted_sum)
        new_centroids[mask] = weighted_sum[mask] / weights_sum[mask, None]        # :311
        if mask.sum() < num_clusters:                                             # :312-313
            assert repair_index is not None, "empty cluster but no repair_index source injected"
            new_centroids[~mask] = torch.stack(
                [X[repair_index()] for _ in range(num_clusters - int(mask.sum()))])
        diff = torch.norm(centroids - new_centroids, dim=1).sum()                 # :314
        if diff < tol:                                                            # :315-316
            break
        centroids = new_centroids                                                 # :317
    return centroids, labels, weights_sum, i


def weighted_kmeans_feature(img_feature: torch.Tensor, video_max_frames: int,
                            init_indices: Sequence[int], weights=None, repair_index=None):
    """Mirror of utiles.py:319-330 including the 3-tuple / 2-tuple arity quirk."""
    T, P, D = img_feature.shape
    if weights is None:
        weights = torch.ones(T, dtype=img_feature.dtype)
    if T <= video_max_frames:
        return img_feature, weights, [[[i] for i in range(T)]]
    X = img_feature.reshape(T, -1)
    centroids, labels, _, _ = weighted_kmeans(X, video_max_frames, init_indices, weights,
                                              repair_index=repair_index)
    return centroids.view(video_max_frames, P, D), labels


def group_by_label_mean(x: torch.Tensor, labels: torch.Tensor, k: int) -> torch.Tensor:
    """utils/utils.py:33-66: one-hot membership, L1-normalised rows (eps 1e-12 -> an empty
    cluster yields an all-zero centre), matmul.  x [BS,N,D], labels [BS,M,N] -> [BS,M,K,D]."""
    bs, n, d = x.shape
    m = labels.shape[1]
    rng = torch.arange(k)
    M = (labels[:, :, :, None] == rng[None, None, None, :]).permute(0, 1, 3, 2).to(x.dtype)
    M = torch.nn.functional.normalize(M, p=1.0, dim=-1)
    return torch.matmul(M, x[:, None, :, :].expand(bs, m, n, d))


def _assign(x: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """kmeans.py:572-597 with LpDistance.pairwise_distance (distances.py:134-141)."""
    bs, n, d = x.shape
    _, num_init, k, _ = centers.shape
    xe = x[:, None, :, None, :].expand(bs, num_init, n, k, d).reshape(-1, d)
    ce = centers[:, :, None, :, :].expand(bs, num_init, n, k, d).reshape(-1, d)
    dist = torch.nn.functional.pairwise_distance(xe, ce, p=2).view(bs, num_init, n, k)
    return torch.argmin(dist, dim=-1)


def torch_kmeans_forward(x: torch.Tensor, centers: torch.Tensor, max_iter: int = 100,
                         tol: float = 1e-4):
    """x [BS,N,D] fp32, centers [BS,num_init,K,D].  Returns labels [BS,N], centers [BS,K,D],
    inertia [BS] of the best restart (kmeans.py:537-570)."""
    x = x.float()
    centers = centers.float().clone()
    bs, n, d = x.shape
    k = centers.shape[2]
    for _ in range(max_iter):
        old = centers.clone()
        c_assign = _assign(x, centers)
        centers = group_by_label_mean(x, c_assign, k)
        dd = torch.norm(centers - old, p=2, dim=-1)                               # :601-607
        dd[dd == float("inf")] = 0
        shift = torch.mean(dd, dim=-1)
        if (shift < tol).all():
            break
    c_assign = _assign(x, centers)
    assigned = centers.gather(index=c_assign[:, :, :, None].expand(-1, -1, -1, d), dim=2)
    dist = torch.norm(x[:, None, :, :].expand(bs, centers.shape[1], n, d) - assigned, p=2, dim=-1) ** 2
    dist[dist == float("inf")] = 0
    inertia = dist.sum(-1)                                                        # :611-632
    best = torch.argmin(inertia, dim=-1)
    b = torch.arange(bs)
    return c_assign[b, best], centers[b, best], inertia[b, best]
