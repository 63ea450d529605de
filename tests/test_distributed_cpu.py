"""CPU, gloo, world_size 2 and 3: host logic of the frame-sharded stream — chunk ownership, the ONE
in-place all-gather of chunk features + caption ids, the redundant replay of the memory updates, answer on rank 0 with
the deferred merge summaries generated on the other ranks — against the single-process StreamingSession on the same
stream: identical tree shape, texts, centroids, short memory and answer ids.  Kernels are not involved: a fake model
stands in for the engines and the k-means of the merges is the CPU oracle (checker infrastructure, tests only)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

P, H, CHUNK = 4, 8, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeModel:
    """CPU stand-in with the engine's surface: encode_images, generate_requests, generate_with_image_embedding.
    Every output is a pure function of its inputs, so that ranks and the single-process run agree."""
    device = torch.device("cpu")

    class config:
        mm_use_im_start_end = False
        hidden_size = H

        class vision:
            grid = 4                                   # (4 // 2) ** 2 = P tokens per frame

    def __init__(self):
        self.log = []

    def encode_images(self, frames):
        base = frames.float().mean(dim=(1, 2, 3))
        return (base[:, None, None] + torch.arange(P * H).view(1, P, H) / 100.0).half()

    def _gen(self, ids, feats, n):
        s = int(ids.sum()) % 97 + (0 if feats is None else int(feats[0].float().sum().abs() * 10) % 89)
        return torch.tensor([[(s + 3 * i) % 200 + 3 for i in range(min(n, 6))]], dtype=torch.long)

    def generate_requests(self, reqs, modalities=None, eos_token_id="config"):
        self.log.append(["caption" if f is not None else "summary" for _, f, _ in reqs])
        return [self._gen(i, f, n) for i, f, n in reqs]

    def generate_with_image_embedding(self, inputs, image_embeddings=None, modalities=None, **kw):
        self.log.append(["answer"])
        return self._gen(inputs, image_embeddings, int(kw.get("max_new_tokens", 6)))


def _patch_kmeans():
    """merges call the CUDA k-means; on the CPU the oracle restatement (same init indices) stands in."""
    from oracle import kmeans as ok
    import streamchat_b200.memory as M

    def wkf(x, k, weights=None, *, init_indices=None, **kw):
        T, p_, d_ = x.shape
        c, labels = ok.weighted_kmeans(x.float().reshape(T, -1), k, init_indices=list(init_indices))[:2]
        return c.to(x.dtype).view(k, p_, d_), labels
    M.weighted_kmeans_feature = wkf
    # ... and the tree search calls the CUDA cosine top-k: the oracle's restatement stands in
    import types
    from oracle import retrieval as oret
    M.ops = types.SimpleNamespace(cosine_topk=lambda index, q, k=1: oret.cos_topk(index.float(), q.float(), k))


def _frames(n):
    g = torch.Generator().manual_seed(5)
    return torch.randn(n, 3, 8, 8, generator=g)


def _session(model):
    import fakes
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession
    cfg = StreamConfig(chunk_size=CHUNK, num_clusters=2, interval=3, short_window=6, remember_window=2, answer_tokens=6)
    return StreamingSession(model, fakes.FakeTokenizer(), HashTextEmbedder(16, "cpu"), None, None, cfg)


def _describe(sess, ids):
    def node(n):
        return (n.depth, n.text, round(float(n.centroids.float().sum()), 3), [node(c) for c in n.children])
    return ([node(n) for n in sess.long_memory_tree],
            [round(float(t.float().sum()), 3) for t in sess.short_memory_buffer], ids.tolist())


SEGMENTS = [(0, 12), (12, 12), (24, 12), (36, 6)]          # 42 frames: 11 chunks, the last one short (2 frames)


def _single(segments=None, total=42):
    segments = segments or SEGMENTS
    _patch_kmeans()
    np.random.seed(11)
    frames = _frames(total)
    sess = _session(_FakeModel())
    for first, n in segments:
        sess.update_memory(sess.ingest(frames[first:first + n]), kmeans_init_indices=[0, 1])
    text, ids = sess.answer("where are my keys")
    return _describe(sess, ids)


def _worker(rank, world, port, ret, segments=None, total=42):
    segments = segments or SEGMENTS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from streamchat_b200.distributed import ShardedStreamingSession, make_gather_buffer, owned_chunks
    _patch_kmeans()
    np.random.seed(11)
    frames = _frames(total)
    model = _FakeModel()
    sess = _session(model)
    buf = make_gather_buffer(model, CHUNK, total, rank, world)
    n_chunks = -(-total // CHUNK)
    assert isinstance(buf, torch.Tensor) and tuple(buf.shape) == (world, -(-n_chunks // world), CHUNK * P, H)
    sh = ShardedStreamingSession(sess, rank, world, total, kmeans_init_indices=[0, 1], gather_buffer=buf)
    n_mine = 0
    for first, n in segments:
        mine = owned_chunks(first // CHUNK, -(-n // CHUNK), rank, world)
        idx = [f for c in mine for f in range(c * CHUNK, min((c + 1) * CHUNK, first + n))]
        n_mine += len(idx)
        sh.ingest_segment(first, n, frames[idx] if idx else None)
    sh.finalize()
    text, ids = sh.answer("where are my keys")
    ret[rank] = (_describe(sess, ids), n_mine, model.log, sh.exchanged_bytes)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, segments=None, total=42):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret, segments, total), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_sharded_stream_equals_single_process_world2():
    want = _single()
    res = _run(2)
    assert res[0][0] == res[1][0] == want                       # every rank ends with the single-process memory + answer
    assert res[0][1] + res[1][1] == 42 and res[0][1] == 22      # chunks 0,2,..,10 (the short one) / 1,3,..,9
    # rank 0: its 6 captions in ONE batch ... the answer last; rank 1: 5 captions, never the answer.  Every merge summary
    # of the tree is generated exactly ONCE across the ranks (those the search reads are spread over all ranks, the
    # root's own — which no search reads — goes to rank 1 while rank 0 answers)
    assert res[0][2][0] == ["caption"] * 6 and res[0][2][-1] == ["answer"]
    assert res[1][2][0] == ["caption"] * 5 and ["answer"] not in res[1][2]

    def count(nodes):
        return sum((1 if n[0] > 0 else 0) + count(n[3]) for n in nodes)
    n_summaries = sum(b.count("summary") for r in res for b in r[2])
    assert n_summaries == count(want[0]) >= 4 and res[1][2][-1] == ["summary"]
    assert res[0][3] == res[1][3] > 0


def test_sharded_stream_equals_single_process_world3():
    want = _single()
    res = _run(3)
    assert all(r[0] == want for r in res)
    assert [r[1] for r in res] == [16, 14, 12]                  # chunks {0,3,6,9} {1,4,7,10 (2 frames)} {2,5,8}
    assert sum(r[1] for r in res) == 42


def test_more_ranks_than_chunks():
    """A short stream (2 chunks) on 3 ranks: the rank that owns nothing still takes part in the gather and the replay and
    ends with the same memory and answer (no merge happens: the tree is two captioned depth-0 nodes)."""
    segs = [(0, 8)]
    want = _single(segs, 8)
    res = _run(3, segs, 8)
    assert all(r[0] == want for r in res) and [r[1] for r in res] == [4, 4, 0]
    assert res[2][2] == [] or all("caption" not in b for b in res[2][2])      # rank 2 generated no caption


def test_chunk_ownership_and_shapes():
    from streamchat_b200.distributed import gather_shape, owned_chunks, shard_range
    assert owned_chunks(5, 5, 1, 8) == [9] and owned_chunks(0, 25, 0, 8) == [0, 8, 16, 24]
    assert sorted(c for r in range(8) for c in owned_chunks(0, 25, r, 8)) == list(range(25))
    assert gather_shape(8, 1000, 40, 144, 3584) == (8, 4, 5760, 3584)
    assert gather_shape(1, 1000, 40, 144, 3584) == (1, 25, 5760, 3584)
    seen = []
    for r in range(3):
        b, e = shard_range(10, r, 3)
        seen += list(range(b, e))
    assert seen == list(range(10))
