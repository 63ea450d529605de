"""GPU: the frame-sharded stream (distributed.ShardedStreamingSession) with the real engines — two processes share
cuda:0 (gloo as the control plane, so one GPU is enough) and run ONE tiny stream sharded by chunks; every rank must end
with the tree, texts, centroids, short memory and answer ids of the single-GPU StreamingSession fed the whole stream.
(The NCCL / NVLink peer-memory gather needs two GPUs: tools/check_sharded_multigpu.py, run under `gpurun --gpus 2`.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_FRAMES, SEG = 88, 24          # segments of 3 chunks of 8 frames + a tail of 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup():
    import streamchat_b200.memory as M
    from streamchat_b200.longva.model.builder import load_pretrained_model
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession
    M.CAPTION_TOKENS, M.SUMMARY_TOKENS = 8, 8
    tokenizer, model, _, _ = load_pretrained_model("synthetic:tiny", None, "llava_qwen", device_map="cuda:0")
    cfg = StreamConfig(chunk_size=8, interval=3, answer_tokens=10)
    np.random.seed(7)
    sess = StreamingSession(model, tokenizer, HashTextEmbedder(1024, "cuda:0"), None, None, cfg)
    g = torch.Generator(device="cuda:0").manual_seed(0)
    frames = torch.randn(N_FRAMES, 3, 56, 56, device="cuda:0", dtype=torch.float16, generator=g)
    return sess, frames


def _describe(sess, ids):
    def node(n):
        return (n.depth, n.text, float(n.centroids.float().sum()), [node(c) for c in n.children])
    return ([node(n) for n in sess.long_memory_tree], [float(t.float().sum()) for t in sess.short_memory_buffer],
            ids.cpu().tolist())


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from streamchat_b200.distributed import ShardedStreamingSession, make_gather_buffer, owned_chunks
    sess, frames = _setup()
    buf = make_gather_buffer(sess.model, 8, N_FRAMES, rank, world, peer=False)
    sh = ShardedStreamingSession(sess, rank, world, N_FRAMES, kmeans_init_indices=[0, 1, 2, 3, 4], gather_buffer=buf)
    for first in range(0, N_FRAMES, SEG):
        n = min(SEG, N_FRAMES - first)
        mine = owned_chunks(first // 8, -(-n // 8), rank, world)
        idx = [f for c in mine for f in range(c * 8, min((c + 1) * 8, first + n))]
        sh.ingest_segment(first, n, frames[idx] if idx else None)
    sh.finalize()
    text, ids = sh.answer("what did I see?")
    ret[rank] = _describe(sess, ids)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stream_on_real_engines_equals_single_gpu(cuda):
    import streamchat_b200.memory as M
    from streamchat_b200.distributed import ShardedStreamingSession
    sess, frames = _setup()
    try:
        with pytest.raises(ValueError, match="chunk boundary"):          # updates must start on a chunk boundary
            ShardedStreamingSession(sess, 0, 1, N_FRAMES).ingest_segment(22, 22, frames[22:44])
        for s0 in range(0, N_FRAMES, SEG):
            sess.update_memory(sess.ingest(frames[s0:s0 + SEG]), kmeans_init_indices=[0, 1, 2, 3, 4])
        text, ids = sess.answer("what did I see?")
        want = _describe(sess, ids)
        assert max(n[0] for n in want[0]) >= 1 and ids.shape == (1, 10)   # merges happened
        # world 1 through the sharded session == the plain session
        sess1, _ = _setup()
        sh = ShardedStreamingSession(sess1, 0, 1, N_FRAMES, kmeans_init_indices=[0, 1, 2, 3, 4])
        for s0 in range(0, N_FRAMES, SEG):
            sh.ingest_segment(s0, min(SEG, N_FRAMES - s0), frames[s0:s0 + SEG])
        sh.finalize()
        assert _describe(sess1, sh.answer("what did I see?")[1]) == want
    finally:
        M.CAPTION_TOKENS, M.SUMMARY_TOKENS = 128, 256
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == want
