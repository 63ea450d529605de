"""Multi-GPU plumbing of the hot path: ONE stream, frame-sharded over the ranks of a box.

Partition (north_star: "incoming frames shard naturally across the 8 GPUs ... a single all-gather of per-shard
memory embeddings before retrieval"): chunk c of the stream (`chunk_size` = 40 consecutive frames, the unit the
reference captions, utiles.py:539-559) belongs to rank c mod W.  A rank encodes only its chunks' frames and generates
only their captions (no collective on the data path: ViT encode, caption prefill and caption decode are independent
per chunk).  The ONE exchange step is an all-gather, in place into the buffer every consumer then reads, of what a
cross-rank memory needs: the chunks' pooled frame features [40*144, H] fp16 — the depth-0 node centroids the merges
(k-means) and the answer's visual context read, 1.03 GB per 1 000 frames — and the caption token ids.  After it every
rank holds the whole memory and replays the tree construction (forgetting-curve short memory, merges) redundantly,
which is cheaper than a second collective (two 413 MB k-means passes); retrieval runs on the gathered index and the
answer is generated on rank 0 while the ranks 1.. generate the merge summaries the answer does not depend on.

The reference has no counterpart (it runs two replicas serially).  torch.distributed (NCCL on GPUs;
gloo in the CPU tests) is the control plane; the gather itself is `allgather_embed`: the library's peer-memory kernel
(scb_allgather_embed over NVLink, ops.PeerGather) when the ranks can map each other's buffers, NCCL otherwise.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items for `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def owned_chunks(first: int, count: int, rank: int, world: int) -> List[int]:
    """Chunk ids in [first, first + count) that belong to `rank` (chunk c -> rank c mod world)."""
    return [c for c in range(first, first + count) if c % world == rank]


def allgather_inplace(buf: torch.Tensor, rank: int, group=None) -> None:
    """All-gather IN PLACE: `buf` is [world, ...]; every rank has filled buf[rank] and leaves with every slice filled.
    NCCL takes the rank's slice of the output as its input (no staging copy); gloo (CPU tests) goes through a list."""
    world = dist.get_world_size(group)
    if buf.shape[0] != world:
        raise ValueError(f"buffer has {buf.shape[0]} slices for a world of {world}")
    if world == 1:
        return
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(buf.view(world, -1), buf[rank].reshape(-1), group=group)
    else:                                   # gloo gathers host tensors only
        mine = buf[rank].detach().cpu().contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        for r, p in enumerate(parts):
            if r != rank:
                buf[r].copy_(p)


def gather_shape(world: int, total_frames: int, chunk: int, tokens_per_frame: int, hidden: int) -> Tuple[int, ...]:
    """[world, chunks per rank, chunk * tokens_per_frame, hidden]: global chunk c lives at [c % world, c // world]."""
    n_chunks = math.ceil(total_frames / chunk)
    return (world, math.ceil(n_chunks / world), chunk * tokens_per_frame, hidden)


def make_gather_buffer(model, chunk: int, total_frames: int, rank: int, world: int, group=None, peer: bool = True):
    """The gather buffer of a sharded stream, allocated ONCE per process: an `ops.PeerGather` (NVLink peer memory,
    moved by the library's own scb_allgather_embed kernel) when every rank can map its peers' buffers, else a plain
    tensor moved by NCCL.  A collective when world > 1 (every rank calls it)."""
    mc = model.config
    shape = gather_shape(world, total_frames, chunk, (mc.vision.grid // 2) ** 2, mc.hidden_size)
    if world > 1 and peer and model.device.type == "cuda":
        from . import _lib as L
        from . import ops
        try:
            return ops.PeerGather(shape, torch.float16, rank, world, model.device, group)
        except L.ScbError:
            pass                                         # no P2P / IPC between some pair: NCCL moves the bytes
    return torch.zeros(shape, dtype=torch.float16, device=model.device)


class ShardedStreamingSession:
    """One stream over `world` ranks.  Drive it like the single-GPU session, segment by segment:

        sh = ShardedStreamingSession(session, rank, world, total_frames)
        for each segment:  sh.ingest_segment(first_frame, n_frames, frames_of_my_chunks)
        sh.finalize()                       # my captions -> the ONE all-gather -> replay of the memory updates
        text, ids = sh.answer(question)     # rank 0 answers, ranks 1.. generate the deferred merge summaries

    Results (tree, texts, short memory, answer ids) are those of `StreamingSession` fed the whole stream on one GPU,
    up to greedy ties below fp16 resolution (a caption decodes as one of <= ceil(25/W) rows here, of 25 there).
    """

    def __init__(self, session, rank: int, world: int, total_frames: int, group=None, kmeans_init_indices=None,
                 gather_buffer=None):
        """`gather_buffer`: what `make_gather_buffer` returned (long-lived: allocate it once per process and reuse it
        for every stream); None allocates a plain tensor moved by NCCL."""
        from . import memory as M
        self.s, self.rank, self.world, self.group = session, rank, world, group
        self.kmeans_init_indices = kmeans_init_indices
        cfg = session.cfg
        self.chunk = cfg.chunk_size
        self.n_chunks = math.ceil(total_frames / self.chunk)
        self.per_rank = math.ceil(self.n_chunks / world)
        self.device = session.model.device
        mc = session.model.config
        self._P, self._H = (mc.vision.grid // 2) ** 2, mc.hidden_size
        self.segments: List[Tuple[int, int]] = []          # (first frame, n frames) in arrival order
        self.frames_in_chunk = [0] * self.n_chunks
        shape = gather_shape(world, total_frames, self.chunk, self._P, self._H)
        self._peer = None
        if gather_buffer is None:
            self._feats = torch.zeros(shape, dtype=torch.float16, device=self.device)
        elif isinstance(gather_buffer, torch.Tensor):
            self._feats = gather_buffer
        else:
            self._peer, self._feats = gather_buffer, gather_buffer.buffer
        if tuple(self._feats.shape) != shape:
            raise ValueError(f"gather buffer has shape {tuple(self._feats.shape)}, the stream needs {shape}")
        self._ids = torch.full((world, self.per_rank, M.CAPTION_TOKENS), -1, dtype=torch.int64, device=self.device)
        self.exchanged_bytes = 0

    # ---- [A] ---------------------------------------------------------------------------------
    def slot(self, chunk: int) -> Tuple[int, int]:
        return chunk % self.world, chunk // self.world

    @torch.no_grad()
    def ingest_segment(self, first_frame: int, n_frames: int, my_frames: Optional[torch.Tensor]) -> None:
        """A segment of the stream arrives: frames [first_frame, first_frame + n_frames).  `my_frames` holds, in
        stream order, ONLY the frames of the chunks this rank owns ([k, 3, S, S]; None / empty when it owns none).
        Segments must start on a chunk boundary (the reference's updates do: 200-frame segments, 40-frame chunks)."""
        if first_frame % self.chunk:
            raise ValueError("segments must start on a chunk boundary")
        self.segments.append((first_frame, n_frames))
        c0, nc = first_frame // self.chunk, math.ceil(n_frames / self.chunk)
        for c in range(c0, c0 + nc):
            self.frames_in_chunk[c] = min(self.chunk, first_frame + n_frames - c * self.chunk)
        mine = owned_chunks(c0, nc, self.rank, self.world)
        want = sum(self.frames_in_chunk[c] for c in mine)
        got = 0 if my_frames is None else int(my_frames.shape[0])
        if got != want:
            raise ValueError(f"rank {self.rank} owns {want} frames of this segment, got {got}")
        if not mine:
            return
        feats = self.s.model.encode_images(my_frames.half())                  # [want, P, H]
        self.s.frames_seen += want
        off = 0
        for c in mine:
            k = self.frames_in_chunk[c]
            r, j = self.slot(c)
            self._feats[r, j, : k * feats.shape[1]].copy_(feats[off:off + k].reshape(-1, feats.shape[2]))
            off += k

    def chunk_features(self, chunk: int) -> torch.Tensor:
        """[frames, P, H] view of a chunk's pooled features inside the gather buffer."""
        r, j = self.slot(chunk)
        k = self.frames_in_chunk[chunk]
        return self._feats[r, j, : k * self._P].view(k, self._P, self._H)

    # ---- the exchange ------------------------------------------------------------------------
    def finalize(self) -> None:
        """My captions -> the ONE all-gather -> replay of the memory updates."""
        self.caption_my_chunks()
        self.allgather_embed()
        self.replay_updates()

    @torch.no_grad()
    def caption_my_chunks(self) -> None:
        """Captions of the chunks this rank owns, decoded in lock-step (utiles.py:539-559; greedy: hazard H2)."""
        from . import memory as M
        s = self.s
        mine = [c for c in range(self.n_chunks) if c % self.world == self.rank and self.frames_in_chunk[c] > 0]
        if mine:
            reqs = [(s._caption_ids, [self.chunk_features(c).reshape(-1, self._H)], M.CAPTION_TOKENS) for c in mine]
            outs = s.summarizer.generate_requests(reqs, modalities=["video"])
            for c, ids in zip(mine, outs):
                r, j = self.slot(c)
                self._ids[r, j, : ids.shape[1]].copy_(ids[0])

    @torch.no_grad()
    def replay_updates(self) -> None:
        """The reference's per-segment memory updates (inference_streaming_longva_v2.py:267-378) on the gathered
        memory — every rank, redundantly: short-memory draw, depth-0 nodes with their captions, merges."""
        from . import memory as M
        s = self.s
        for first, n in self.segments:
            c0, nc = first // self.chunk, math.ceil(n / self.chunk)
            buffer_cache: List[torch.Tensor] = []
            nodes = []
            for c in range(c0, c0 + nc):
                f = self.chunk_features(c)
                buffer_cache.extend(torch.split(f, 1))
                r, j = self.slot(c)
                ids = self._ids[r, j]
                ids = ids[ids >= 0].unsqueeze(0)
                text = s.tokenizer.batch_decode(ids, skip_special_tokens=True)[0].strip()
                nodes.append(M.MultimodalTreeNode(f, text, depth=0))
            s.update_memory(buffer_cache, kmeans_init_indices=self.kmeans_init_indices, chunk_nodes=nodes)

    def allgather_embed(self) -> None:
        """Every rank's chunk features and caption ids into every rank's buffers (rank r's slice is [r])."""
        self.exchanged_bytes = (self._feats[0].numel() * 2 + self._ids[0].numel() * 8) * (self.world - 1)
        if self.world == 1:
            return
        if self._peer is not None:
            self._peer.allgather()                          # scb_allgather_embed: push over NVLink peer memory
        else:
            allgather_inplace(self._feats, self.rank, self.group)
        allgather_inplace(self._ids, self.rank, self.group)

    # ---- [C] ---------------------------------------------------------------------------------
    def _summaries_across_ranks(self, ready, ranks: Sequence[int], extra_rows: int = 0) -> torch.Tensor:
        """Generate the ready pending summaries `ready[i]` on rank `ranks[i % len(ranks)]` (each rank decodes its share
        in one lock-step batch), all-gather the ids and set the texts on EVERY rank.  Returns the gathered id block
        [world, slots, width] (slot rows beyond a rank's summaries — `extra_rows` of them — are the caller's)."""
        from . import memory as M
        s = self.s
        share = {r: [i for i in range(len(ready)) if ranks[i % len(ranks)] == r] for r in range(self.world)}
        n_slots = max(1, max(len(v) for v in share.values()) + extra_rows)
        out = torch.full((self.world, n_slots, max(M.SUMMARY_TOKENS, s.cfg.answer_tokens)), -1, dtype=torch.int64,
                         device=self.device)
        mine = share[self.rank]
        if mine:
            outs = s.summarizer.generate_requests([M.summary_request(ready[i]) for i in mine], modalities=["video"])
            for k, ids in enumerate(outs):
                out[self.rank, extra_rows + k, : ids.shape[1]].copy_(ids[0])
        return out, share

    def _set_summary_texts(self, ready, out: torch.Tensor, share, extra_rows: int = 0) -> None:
        for r, idxs in share.items():
            for k, i in enumerate(idxs):
                ids = out[r, extra_rows + k]
                ready[i].set_text(self.s.tokenizer.batch_decode(ids[ids >= 0].unsqueeze(0), skip_special_tokens=True)[0].strip())

    @torch.no_grad()
    def answer(self, question: str, history_prompt: Optional[str] = None):
        """Texts the search reads that are still pending (summaries below a deeper root) are generated level by level,
        spread over all ranks.  Then rank 0 generates the answer while the roots' merge summaries — never read by
        the search — are generated on ranks 1..; one small all-gather of ids leaves every rank with the complete
        tree and the answer ids."""
        from . import memory as M
        s = self.s
        if self.world == 1:
            return s.answer(question, history_prompt)
        tree = s.long_memory_tree
        if tree is not None:
            needed = {id(n) for n in M._all_nodes(M.texts_read_by_search(tree))}
            while True:
                ready = [n for n in M.ready_summaries(tree) if id(n) in needed]
                if not ready:
                    break
                out, share = self._summaries_across_ranks(ready, list(range(self.world)))
                allgather_inplace(out, self.rank, self.group)
                self._set_summary_texts(ready, out, share)
        input_ids, image_embeddings = s.build_question_inputs(question, history_prompt)
        ready = M.ready_summaries(tree) if tree is not None else []
        out, _ = self._answer_round(ready, input_ids, image_embeddings)
        a = out[0, 0]
        output_ids = a[a >= 0].unsqueeze(0)
        text = s.tokenizer.batch_decode(output_ids, skip_special_tokens=True)[0].strip()
        return text, output_ids

    def _answer_round(self, ready, input_ids, image_embeddings):
        """Row [0, 0] of the gathered block is the answer (rank 0); ranks 1.. fill their rows with root summaries."""
        s = self.s
        helpers = list(range(1, self.world))
        # the same slot layout on every rank: slot 0 of rank 0 is the answer, summaries start at slot 0 on ranks 1..
        out, share = self._summaries_across_ranks(ready, helpers)
        if self.rank == 0:
            ids = s.model.generate_with_image_embedding(
                input_ids, image_embeddings=[image_embeddings], modalities=["video"], do_sample=False, temperature=0,
                top_p=None, num_beams=1, max_new_tokens=s.cfg.answer_tokens, use_cache=False)
            out[0, 0, : ids.shape[1]].copy_(ids[0])
        allgather_inplace(out, self.rank, self.group)
        self._set_summary_texts(ready, out, share)
        return out, share
