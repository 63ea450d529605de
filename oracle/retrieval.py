"""CPU oracle for cosine retrieval (TEST INFRASTRUCTURE — see oracle/__init__.py).

cos_sim restates sentence_transformers.util.cos_sim (sentence-transformers==2.2.2,
requirements.txt:126; not vendored, not installed): both operands are made 2-D, L2-normalised
with torch.nn.functional.normalize(p=2, dim=1) (eps 1e-12) and multiplied.  Reference call
sites: utiles.py:17 (import), :732 and :764 (`cos_sim(text_embedding[0], query_embedding[0])`).
PARITY UNPINNED: the reference has no test or fixture for this function.
"""
from __future__ import annotations

import numpy as np
import torch


def cos_sim(a, b) -> torch.Tensor:
    a = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a)
    b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    a_norm = torch.nn.functional.normalize(a.float(), p=2, dim=1)
    b_norm = torch.nn.functional.normalize(b.float(), p=2, dim=1)
    return torch.mm(a_norm, b_norm.transpose(0, 1))


def cos_topk(index: torch.Tensor, queries: torch.Tensor, k: int):
    """Scores [Q,k] and row indices [Q,k]; ties resolved to the lowest row (stable sort)."""
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    sims = cos_sim(queries, index)  # [Q, N]
    n = sims.shape[1]
    order = torch.argsort(sims, dim=1, descending=True, stable=True)[:, :k]
    scores = torch.gather(sims, 1, order)
    if n < k:
        pad = k - n
        order = torch.cat([order, torch.full((sims.shape[0], pad), -1, dtype=order.dtype)], 1)
        scores = torch.cat([scores, torch.full((sims.shape[0], pad), float("-inf"))], 1)
    return scores, order


def running_argmax_top1(sims, floor: float = 0.0, default=None):
    """The reference's selection rule (utiles.py:719-742 / :751-772): best_sim starts at 0,
    strict '>' so the first maximum wins and non-positive similarities never win."""
    best, best_i = floor, default
    for i, s in enumerate(sims):
        if float(s) > best:
            best, best_i = float(s), i
    return best_i
