"""CPU, world_size 2, gloo: frame sharding and the single memory-embedding all-gather (§8e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from streamchat_b200.distributed import gather_memory_embeddings, live_rows, pack_memory_embeddings, shard_range
    b, e = shard_range(1001, rank, world)
    n_nodes = 3 + rank                                   # ranks hold different numbers of memory nodes
    embs = [torch.full((8,), float(100 * rank + i)) for i in range(n_nodes)]
    packed = pack_memory_embeddings(embs, max_rows=6, dim=8, device="cpu")
    gathered = gather_memory_embeddings(packed)
    index, owner, local = live_rows(gathered)
    ret[rank] = (b, e, gathered.shape, index[:, 0].tolist(), owner.tolist(), local.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (b0, e0, shp0, vals0, own0, loc0), (b1, e1, shp1, vals1, own1, loc1) = ret[0], ret[1]
    assert (b0, e0, b1, e1) == (0, 501, 501, 1001)       # contiguous, disjoint, complete
    assert tuple(shp0) == tuple(shp1) == (12, 9)
    assert vals0 == vals1 == [0.0, 1.0, 2.0, 100.0, 101.0, 102.0, 103.0]   # every rank sees the full index
    assert own0 == [0, 0, 0, 1, 1, 1, 1] and loc0 == [0, 1, 2, 0, 1, 2, 3]


def _exchange_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from streamchat_b200 import memory as M
    from streamchat_b200.distributed import exchange_memory_index
    from streamchat_b200.stream import HashTextEmbedder
    emb = HashTextEmbedder(32, "cpu")
    calls = []

    class Summ:                       # records how the pending texts were generated (one batch, not per node)
        device = torch.device("cpu")

        def generate_requests(self, reqs, modalities=None):
            calls.append(len(reqs))
            return [torch.tensor([[rank, i]]) for i in range(len(reqs))]

    class Tok:
        def batch_decode(self, ids, skip_special_tokens=True):
            return ["rank%d text%d" % tuple(ids[0].tolist())]

    nodes = []
    for i in range(2 + rank):
        n = M.MultimodalTreeNode(torch.zeros(1, 4, 8), None)
        n._pending = {"kind": "caption", "summarizer": Summ(), "tokenizer": Tok(), "input_ids": torch.zeros(1, 3),
                      "features": torch.zeros(4, 8), "max_new": 4}
        nodes.append(n)
    index, owner, local = exchange_memory_index(nodes, emb, None, "cpu", max_rows=4)
    expect = torch.stack([emb(None, f"rank{r} text{i}") for r in range(world) for i in range(2 + r)])
    ret[rank] = (calls, torch.equal(index, expect), owner.tolist(), local.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_memory_index_world2():
    """Pending texts are generated in ONE lock-step batch per rank before the all-gather, and every rank ends
    up with the same global index in rank order."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == [2] and ret[1][0] == [3]
    assert ret[0][1] and ret[1][1]
    assert ret[0][2] == ret[1][2] == [0, 0, 1, 1, 1]
    assert ret[0][3] == ret[1][3] == [0, 1, 0, 1, 2]


def test_shard_range_properties():
    from streamchat_b200.distributed import shard_range
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
