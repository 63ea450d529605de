"""Asynchronous stages of the streaming loop with CUDA streams and events instead of the Python
threads the reference's earlier drivers used (previous_version/streaming_demo_llava_next_3.py:968-989 reader / updater /
QA threads over threading.Event; utiles.py:2129-2230) — the v2 driver runs the stages serially
(inference_streaming_longva_v2.py:862-911).

What can overlap on this path, and what is built:
  * host -> device staging of the NEXT segment's frames while the current one is encoded — `FramePrefetcher`
    (copy stream + two device buffers + ready / free events);
  * the answer on rank 0 while the other ranks generate the merge summaries it does not depend on —
    distributed.ShardedStreamingSession.answer; on one GPU the same summaries ride in the answer's lock-step batch
    (stream.StreamingSession.answer), which removes a whole 255-step decode loop;
  * NOT built, because it cannot pay here: co-scheduling the HBM-bound caption decode with the tensor-bound
    encode / prefill of later segments.  Captions are lazy — all 25 of a stream decode in ONE lock-step batch at the
    question (one weight pass per step for 25 rows: 0.71 s); decoding them segment by segment so that they could
    overlap with the next segment's encode would take 5 x 127 steps of a 5-row batch (2.4 s of HBM time) to hide
    0.5 s of encode — and both the persistent decode kernel and the tcgen05 GEMMs are one-CTA-per-SM residents, so
    sharing the GPU splits the SMs rather than filling idle ones (DESIGN.md §6b has the numbers).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import torch


class FramePrefetcher:
    """Iterate over the segments of a pinned host frame tensor as DEVICE tensors, copying segment i+1 on a side stream
    while the consumer works on segment i.

        for i, seg in FramePrefetcher(host_frames, slices, device):
            feats = session.ingest(seg)          # enqueued on the current stream

    Two device buffers alternate.  Per segment: the copy stream waits until the consumer's work on the buffer it is
    about to overwrite has finished (a `free` event recorded on the consumer's stream when the NEXT segment is
    requested), copies, and records `ready`; the consumer's stream waits on `ready` — no host synchronisation."""

    def __init__(self, host_frames: torch.Tensor, slices: Sequence[Tuple[int, int]], device):
        if not host_frames.is_pinned():
            raise ValueError("FramePrefetcher needs pinned host memory (page-locked frames copy asynchronously)")
        self.host, self.slices, self.device = host_frames, list(slices), torch.device(device)
        longest = max((b - a for a, b in self.slices), default=0)
        shape = (max(longest, 1),) + tuple(host_frames.shape[1:])
        self._bufs = [torch.empty(shape, dtype=host_frames.dtype, device=self.device) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self.bytes_copied = 0

    def __iter__(self) -> Iterator[Tuple[int, Optional[torch.Tensor]]]:
        dev = self.device
        ready: List[Optional[torch.cuda.Event]] = [None, None]
        free: List[Optional[torch.cuda.Event]] = [None, None]

        def start(i: int) -> None:
            a, b = self.slices[i]
            if b <= a:
                return
            j = i & 1
            with torch.cuda.stream(self._copy_stream):
                if free[j] is not None:
                    self._copy_stream.wait_event(free[j])
                self._bufs[j][: b - a].copy_(self.host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                ready[j] = ev
            self.bytes_copied += (b - a) * self.host[0].numel() * self.host.element_size()

        if self.slices:
            start(0)
        for i, (a, b) in enumerate(self.slices):
            if i + 1 < len(self.slices):
                start(i + 1)                      # overlaps with the consumer's work on segment i
            j = i & 1
            if b <= a:
                yield i, None
                continue
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready[j])
            yield i, self._bufs[j][: b - a]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))   # everything the consumer enqueued on this buffer so far
            free[j] = ev
