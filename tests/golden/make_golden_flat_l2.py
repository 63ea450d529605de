"""Fixture pinning the flat-L2 search semantics the dialogue retrieval relies on (faiss.IndexFlatL2 under langchain
0.0.142 `FAISS.similarity_search_with_score_by_vector`, reached at memory_bank/memory_retrieval/local_doc_qa.py:140).
faiss-cpu (requirements.txt:32, ==1.8.0.post1) is NOT installed here and not vendored in the reference, so the fixture
is produced by an independent float64 brute force of the PUBLISHED definition — exhaustive search, metric = squared
Euclidean distance, results ordered by increasing distance, label -1 / distance +inf(3.4e38 in faiss) when fewer than
k vectors exist — on inputs without distance ties:

    python tests/golden/make_golden_flat_l2.py        -> flat_l2_ref.npz

Both the oracle (oracle/memory_bank.index_flat_l2_search) and the GPU mirror (cosine_topk on unit rows, score
2 - 2 cos) are tested against it.  This pins the restatement against an independent computation; it cannot pin it
against the faiss binary itself (stated in DESIGN.md §2).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20240501)
    xb = rng.standard_normal((257, 384))
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)              # MiniLM embeddings are unit rows
    xq = rng.standard_normal((9, 384))
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    xq[0] = xb[100]                                               # an exact hit: distance 0
    xb32, xq32 = xb.astype(np.float32), xq.astype(np.float32)
    d = ((xq32.astype(np.float64)[:, None, :] - xb32.astype(np.float64)[None, :, :]) ** 2).sum(-1)
    k = 8
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    dist = np.take_along_axis(d, order, axis=1)
    gaps = np.diff(np.sort(d, axis=1)[:, : k + 1], axis=1)
    assert gaps.min() > 1e-6, "fixture must not contain distance ties"
    np.savez_compressed(os.path.join(HERE, "flat_l2_ref.npz"), xb=xb32, xq=xq32, k=np.array(k), labels=order.astype(np.int64),
                        distances=dist.astype(np.float32), short_labels=np.array([[1, 0, -1, -1]], dtype=np.int64))
    print("wrote flat_l2_ref.npz", order[0][:3], dist[0][:3])


if __name__ == "__main__":
    main()
