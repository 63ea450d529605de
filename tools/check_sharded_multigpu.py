#!/usr/bin/env python
"""Multi-GPU check of the frame-sharded stream (run under torchrun on >= 2 GPUs, e.g. `gpurun --gpus 2`):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded_multigpu.py

Every rank (own GPU, NCCL) runs ONE tiny stream sharded by chunks, once with the library's peer-memory gather
(scb_allgather_embed over NVLink) and once with NCCL's all-gather, and must end with the tree / texts / centroids /
short memory / answer ids of the single-GPU StreamingSession fed the whole stream (computed on every rank).  Also times
the two gathers on a LongVA-7B-sized buffer.  Prints one JSON line on rank 0; exit code != 0 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_FRAMES, SEG, CHUNK = 88, 24, 8


def describe(sess, ids):
    def node(n):
        return (n.depth, n.text, float(n.centroids.float().sum()), [node(c) for c in n.children])
    return ([node(n) for n in sess.long_memory_tree], [float(t.float().sum()) for t in sess.short_memory_buffer],
            ids.cpu().tolist())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    import streamchat_b200.memory as M
    from streamchat_b200 import ops
    from streamchat_b200.distributed import (ShardedStreamingSession, allgather_inplace, make_gather_buffer,
                                             owned_chunks)
    from streamchat_b200.longva.model.builder import load_pretrained_model
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession
    M.CAPTION_TOKENS, M.SUMMARY_TOKENS = 8, 8
    tokenizer, model, _, _ = load_pretrained_model("synthetic:tiny", None, "llava_qwen", device_map=str(device))
    cfg = StreamConfig(chunk_size=CHUNK, interval=3, answer_tokens=10)
    g = torch.Generator(device=device).manual_seed(0)
    frames = torch.randn(N_FRAMES, 3, 56, 56, device=device, dtype=torch.float16, generator=g)
    embedder = HashTextEmbedder(1024, device)

    def session():
        np.random.seed(7)
        return StreamingSession(model, tokenizer, embedder, None, None, cfg)

    sess = session()
    for s0 in range(0, N_FRAMES, SEG):
        sess.update_memory(sess.ingest(frames[s0:s0 + SEG]), kmeans_init_indices=[0, 1, 2, 3, 4])
    want = describe(sess, sess.answer("what did I see?")[1])
    result = {"world": world, "modes": {}}
    ok = True
    for mode in ("peer", "nccl"):
        buf = make_gather_buffer(model, CHUNK, N_FRAMES, rank, world, peer=(mode == "peer"))
        used = "nccl" if isinstance(buf, torch.Tensor) else "peer"
        sess = session()
        sh = ShardedStreamingSession(sess, rank, world, N_FRAMES, kmeans_init_indices=[0, 1, 2, 3, 4], gather_buffer=buf)
        for first in range(0, N_FRAMES, SEG):
            n = min(SEG, N_FRAMES - first)
            mine = owned_chunks(first // CHUNK, -(-n // CHUNK), rank, world)
            idx = [f for c in mine for f in range(c * CHUNK, min((c + 1) * CHUNK, first + n))]
            sh.ingest_segment(first, n, frames[idx] if idx else None)
        sh.finalize()
        got = describe(sess, sh.answer("what did I see?")[1])
        same = torch.tensor([int(got == want)], device=device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        result["modes"][mode] = {"gather_used": used, "equals_single_gpu": bool(same.item())}
        ok = ok and bool(same.item())
        if not isinstance(buf, torch.Tensor):
            buf.close()
    # ---- gather timing on the LongVA-7B stream's buffer: [world, ceil(25/world), 5760, 3584] fp16
    per = -(-25 // world)
    shape = (world, per, 5760, 3584)
    timing = {}
    for mode in ("peer", "nccl"):
        try:
            pg = ops.PeerGather(shape, torch.float16, rank, world, device) if mode == "peer" else None
        except Exception as e:
            timing[mode] = f"unavailable: {e}"
            continue
        buf = pg.buffer if pg else torch.zeros(shape, dtype=torch.float16, device=device)
        buf[rank].fill_(rank + 1)
        ms = []
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if pg:
                pg.allgather()
            else:
                allgather_inplace(buf, rank)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t))
        good = all(bool((buf[r] == r + 1).all()) for r in range(world))
        flag = torch.tensor([int(good)], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        recv = buf[0].numel() * 2 * (world - 1)
        timing[mode] = {"ms_best": min(ms[1:]), "ms_all": ms, "bytes_received_per_rank": recv,
                        "GBps_per_rank": recv / (min(ms[1:]) * 1e-3) / 1e9, "correct": bool(flag.item())}
        dist.barrier()
        if pg:
            pg.close()
    result["gather_timing"] = timing
    result["ok"] = ok
    if rank == 0:
        print(json.dumps(result))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
