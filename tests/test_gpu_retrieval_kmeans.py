"""GPU parity: cosine_topk and k-means kernels (through the C-ABI) vs the CPU oracle and the
reference-generated golden vectors.  Index/label outputs must match bit-exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import kmeans as okm
from oracle import retrieval as oret

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,q,k,dtype", [
    (10000, 1024, 1, 1, torch.float32),      # BASELINE config #3
    (10000, 1024, 1, 8, torch.float32),
    (10000, 1024, 1, 8, torch.float16),
    (10000, 1024, 3, 4, torch.float32),
    (777, 384, 2, 5, torch.float32),         # MiniLM width (a-R2), generic vector path
    (333, 100, 1, 32, torch.float32),        # dim not a multiple of 128
    (50, 7, 2, 3, torch.float32),            # scalar path (rows not 16-byte aligned)
    (3, 64, 1, 8, torch.float32),            # fewer rows than k
    (1, 1024, 1, 1, torch.float16),
])
def test_cosine_topk_matches_oracle(cuda, n, d, q, k, dtype):
    from streamchat_b200 import ops
    torch.manual_seed(0)
    index = torch.randn(n, d).to(dtype)
    queries = torch.randn(q, d).to(dtype)
    es, ei = oret.cos_topk(index.float(), queries.float(), k)
    s, i = ops.cosine_topk(index.to(cuda), queries.to(cuda), k)
    torch.cuda.synchronize()
    assert torch.equal(i.cpu(), ei), (i.cpu(), ei)
    valid = ei >= 0
    np.testing.assert_allclose(s.cpu()[valid].numpy(), es[valid].numpy(), rtol=1e-5, atol=2e-6)
    assert torch.isinf(s.cpu()[~valid]).all()


def test_cosine_topk_edge_cases(cuda):
    from streamchat_b200 import ops
    torch.manual_seed(1)
    base = torch.randn(64, 256)
    # exact duplicates: lowest row wins (hazard H4); scaling a row does not change its cosine
    index = torch.cat([base, base[:4] * 1.0], 0).to(cuda)
    s, i = ops.cosine_topk(index, base[2].to(cuda), 2)
    assert i[0].tolist() == [2, 66]
    # empty index
    s, i = ops.cosine_topk(torch.empty(0, 256, device=cuda), base[0].to(cuda), 3)
    assert i[0].tolist() == [-1, -1, -1]
    # zero row: norm clamps to eps, similarity 0, never NaN
    z = torch.zeros(5, 256, device=cuda)
    s, i = ops.cosine_topk(z, base[0].to(cuda), 2)
    assert torch.equal(s.cpu(), torch.zeros(1, 2)) and i[0].tolist() == [0, 1]
    # strided index view (leading dimension > dim)
    wide = torch.randn(100, 512, device=cuda)
    s1, i1 = ops.cosine_topk(wide[:, :256], base[0].to(cuda), 4)
    s2, i2 = ops.cosine_topk(wide[:, :256].contiguous(), base[0].to(cuda), 4)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    # repeatability (workspace tickets are left clean)
    s3, i3 = ops.cosine_topk(wide[:, :256], base[0].to(cuda), 4)
    assert torch.equal(i1, i3) and torch.equal(s1, s3)


def test_cosine_topk_large_self_retrieval(cuda):
    """Full-size property check (1M x 1024 fp16, the stream-bound scale-up of config #3):
    a query copied from row r must retrieve r first with similarity 1."""
    from streamchat_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(3)
    index = torch.randn(1_000_000, 1024, device=cuda, dtype=torch.float16, generator=g)
    rows = [0, 123_457, 999_999]
    s, i = ops.cosine_topk(index, index[rows], 2)
    assert i[:, 0].tolist() == rows
    np.testing.assert_allclose(s[:, 0].cpu().numpy(), 1.0, atol=1e-3)
    assert (s[:, 1] < 0.5).all()


def test_weighted_kmeans_feature_matches_reference_golden(cuda, golden_dir):
    from streamchat_b200.kmeans import weighted_kmeans_feature
    g = np.load(os.path.join(golden_dir, "kmeans_ref.npz"))
    for name in "abcdef":
        x = torch.from_numpy(g[f"{name}_x"]).to(cuda)
        K = int(g[f"{name}_k"])
        res = weighted_kmeans_feature(x, K, init_indices=g[f"{name}_init"].tolist())
        assert len(res) == int(g[f"{name}_arity"]), name
        if len(res) == 2:
            np.testing.assert_array_equal(res[1].cpu().numpy(), g[f"{name}_labels"], err_msg=name)
            np.testing.assert_allclose(res[0].cpu().numpy(), g[f"{name}_centroids"], rtol=1e-5, atol=1e-5)
        else:
            assert res[0] is x


def test_weighted_kmeans_fp16_vs_oracle_and_empty_cluster(cuda):
    from streamchat_b200.kmeans import weighted_kmeans_feature
    torch.manual_seed(5)
    x = torch.randn(48, 6, 160).half()
    init = [3, 9, 27, 40, 11]
    ec, el = okm.weighted_kmeans_feature(x.float(), 5, init_indices=init)
    c, l = weighted_kmeans_feature(x.to(cuda), 5, init_indices=init)
    assert c.dtype == torch.float16
    np.testing.assert_array_equal(l.cpu().numpy(), el.numpy())
    np.testing.assert_allclose(c.float().cpu().numpy(), ec.numpy(), rtol=2e-3, atol=2e-3)
    # duplicate initial centroids -> an empty cluster -> repair path (utiles.py:312-313)
    x2 = torch.randn(20, 1, 32)
    x2[1] = x2[0]
    it = iter([7, 12, 3, 5, 9, 1, 2, 4, 6, 8])
    ec2, el2 = okm.weighted_kmeans_feature(x2, 3, init_indices=[0, 1, 5], repair_index=lambda: next(it))
    c2, l2 = weighted_kmeans_feature(x2.to(cuda), 3, init_indices=[0, 1, 5],
                                     repair_indices=[7, 12, 3, 5, 9, 1, 2, 4, 6, 8])
    np.testing.assert_array_equal(l2.cpu().numpy(), el2.numpy())
    np.testing.assert_allclose(c2.cpu().numpy(), ec2.numpy(), rtol=1e-5, atol=1e-5)


def test_torch_kmeans_dropin_matches_reference_golden(cuda, golden_dir):
    from streamchat_b200.kmeans import KMeans
    g = np.load(os.path.join(golden_dir, "torch_kmeans_ref.npz"))
    for name in ("cfg1", "small"):
        x = torch.from_numpy(g[f"{name}_x"]).to(cuda)
        model = KMeans(n_clusters=int(g[f"{name}_k"]), verbose=False)
        res = model(x, centers=torch.from_numpy(g[f"{name}_init"]).to(cuda))
        np.testing.assert_array_equal(res.labels.cpu().numpy(), g[f"{name}_labels"])
        np.testing.assert_allclose(res.centers.cpu().numpy(), g[f"{name}_centers"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(res.inertia.cpu().numpy(), g[f"{name}_inertia"], rtol=1e-4)
        assert res.labels.dtype == torch.int64 and res.x_org is x
        # fit / predict surface
        model.fit(x, centers=torch.from_numpy(g[f"{name}_init"]).to(cuda))
        np.testing.assert_array_equal(model.predict(x).cpu().numpy(), g[f"{name}_labels"])


def test_kmeans_merge_size_properties(cuda):
    """Full merge size of the memory tree (400 frames x 144 x 3584 fp16 = 413 MB, SURVEY §8a a-S3):
    size-independent properties — labels in range, every centroid is the mean of its members,
    well-separated planted clusters are recovered exactly."""
    from streamchat_b200 import ops
    T, F, K = 400, 144 * 3584, 5
    g = torch.Generator(device=cuda).manual_seed(11)
    planted = torch.arange(T, device=cuda) % K
    x = (torch.randn(T, F, device=cuda, generator=g) * 0.05)
    x += (planted[:, None].float() - 2.0)          # cluster j sits around the constant (j-2)
    x = x.half()
    cent = x[:K].float().unsqueeze(0).contiguous()  # rows 0..4 belong to clusters 0..4
    r = ops.kmeans_step(x, cent, None, do_update=True, want_dist=True)
    labels = r["labels"][0].long()
    assert torch.equal(labels, planted)
    assert torch.equal(r["wsum"][0].cpu(), torch.full((K,), T / K))
    for j in range(K):
        mean_j = x[planted == j].float().mean(0)
        assert torch.allclose(r["centroids"][0, j], mean_j, rtol=1e-4, atol=1e-4)
    # idempotence: a second step from the means moves nothing and keeps the labels
    r2 = ops.kmeans_step(x, r["centroids"], None, do_update=True)
    assert torch.equal(r2["labels"][0].long(), planted)
    assert float(r2["shift"].sum()) < 1e-2


def test_kmeans_dot_assign_matches_exact_assign(cuda):
    """Large fp16 problems that only need labels take the tensor-pipe dot-product assign
    (argmin_k |c_k|^2 - 2 x.c_k, centroids as fp16 hi+lo pairs); asking for distances forces the exact
    (x-c)^2 kernel.  Same labels, same updated centroids — on clustered AND on structureless points
    (ragged sizes: 70 points = 4.4 groups of 16, dim not a multiple of the 2048-column tile)."""
    from streamchat_b200 import ops
    torch.manual_seed(7)
    n, d, k = 70, 262144 + 1032, 5
    for clustered in (True, False):
        x = torch.randn(n, d, device=cuda)
        if clustered:
            centers = torch.randn(k, d, device=cuda) * 2
            x = x * 0.5 + centers[torch.randint(0, k, (n,), device=cuda)]
        x = x.half()
        cent = x[:k].float().unsqueeze(0).contiguous()
        for _ in range(3):
            fast = ops.kmeans_step(x, cent, None, do_update=True, want_dist=False)
            exact = ops.kmeans_step(x, cent, None, do_update=True, want_dist=True)
            assert torch.equal(fast["labels"], exact["labels"])
            assert torch.equal(fast["centroids"], exact["centroids"])
            cent = exact["centroids"]
            cent[exact["wsum"] == 0] = x[0].float()       # keep empty clusters alive (host repair rule, not under test)
