"""Multi-GPU plumbing of the hot path: frames shard by rank with no data-path
collective; the single exchange step is one all-gather of the per-rank memory embeddings before
retrieval.  The reference has no counterpart (it runs two model replicas serially, fact F9).
Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests) — no arithmetic here."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items for `rank` (keeps 40-frame chunks on one rank); the first
    n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pack_memory_embeddings(embeddings: Sequence[torch.Tensor], max_rows: int, dim: int,
                           device) -> torch.Tensor:
    """[max_rows, dim + 1] fp32: rows = node text embeddings (padded with zeros), last column = 1 for
    live rows.  Fixed shape so that one all_gather_into_tensor moves every rank's memory."""
    out = torch.zeros((max_rows, dim + 1), dtype=torch.float32, device=device)
    n = min(len(embeddings), max_rows)
    for i in range(n):
        out[i, :dim] = embeddings[i].to(device=device, dtype=torch.float32)
        out[i, dim] = 1.0
    return out


def gather_memory_embeddings(packed: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks' packed blocks concatenated in rank order: [world * max_rows, dim + 1]."""
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    if dist.get_backend(group) == "gloo":
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed, group=group)
        out.copy_(torch.cat(parts, 0))
    else:
        dist.all_gather_into_tensor(out, packed, group=group)
    return out


def live_rows(gathered: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(index [n_live, dim], owner rank [n_live], local row [n_live]) of the gathered block."""
    dim = gathered.shape[1] - 1
    rows = (gathered[:, dim] > 0).nonzero().flatten()
    world = dist.get_world_size() if dist.is_initialized() else 1
    per = gathered.shape[0] // world
    return gathered[rows, :dim].contiguous(), rows // per, rows % per


def exchange_memory_index(tree, embedding_model, embedding_tokenizer, device, max_rows: int = 64, group=None):
    """The ONE exchange step of the sharded stream (§8e): every rank embeds the root nodes of its own
    memory tree (pending captions / summaries are generated first, in lock-step batches — never one
    node at a time — and all texts go through one packed encoder forward), then a single all-gather
    gives every rank the global index.  Returns (index [n_live, D] fp32, owner rank, local row)."""
    from . import memory as M
    nodes = list(tree or [])[:max_rows]
    if nodes:
        embs = M.node_embeddings(nodes, embedding_model, embedding_tokenizer)
        dim = embs.shape[1]
    else:
        embs, dim = [], getattr(embedding_model, "dim", 1024)
    packed = pack_memory_embeddings(list(embs), max_rows, dim, device)
    return live_rows(gather_memory_embeddings(packed, group))
