"""CPU: pins oracle/ against golden vectors produced by the reference itself
(tests/golden/make_golden.py ran /root/reference's code)."""
import os

import numpy as np
import torch

from oracle import kmeans as okm
from oracle import memory as omem
from oracle import retrieval as oret


def test_weighted_kmeans_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "kmeans_ref.npz"))
    for name in "abcdef":
        x = torch.from_numpy(g[f"{name}_x"])
        K = int(g[f"{name}_k"])
        res = okm.weighted_kmeans_feature(x, K, init_indices=g[f"{name}_init"].tolist())
        assert len(res) == int(g[f"{name}_arity"]), name
        if len(res) == 2:
            np.testing.assert_array_equal(res[1].numpy(), g[f"{name}_labels"], err_msg=name)
            np.testing.assert_allclose(res[0].numpy(), g[f"{name}_centroids"], rtol=1e-6, atol=1e-6)
        else:  # T <= K: input returned untouched with the 3-tuple arity (utiles.py:321-322)
            assert res[0] is x and res[2] == [[[i] for i in range(x.shape[0])]]


def test_torch_kmeans_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "torch_kmeans_ref.npz"))
    for name in ("cfg1", "small"):
        x = torch.from_numpy(g[f"{name}_x"])
        labels, centers, inertia = okm.torch_kmeans_forward(x, torch.from_numpy(g[f"{name}_init"]))
        np.testing.assert_array_equal(labels.numpy(), g[f"{name}_labels"])
        np.testing.assert_allclose(centers.numpy(), g[f"{name}_centers"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(inertia.numpy(), g[f"{name}_inertia"], rtol=1e-5)


def test_forgetting_curve_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "memory_ref.npz"))
    np.testing.assert_array_equal(omem.calculate_forgetting_probabilities(20, tau=5), g["p_20_5"])
    np.testing.assert_array_equal(omem.calculate_forgetting_probabilities(7), g["p_7_10"])
    np.testing.assert_array_equal(omem.calculate_forgetting_probabilities(1, tau=5), g["p_1_5"])
    p = omem.calculate_forgetting_probabilities(20, tau=5)
    assert p[0] == p.max() and abs(p.sum() - 1) < 1e-12          # oldest frame most likely (H9)
    np.random.seed(42)
    sel = omem.select_data_without_replacement(list(range(20)), p, 5)
    np.testing.assert_array_equal(np.array(sel), g["sel_20_5_seed42"])
    assert len(set(sel)) == 5


def test_cos_topk_oracle_properties():
    torch.manual_seed(0)
    idx = torch.randn(200, 64)
    q = torch.randn(3, 64)
    s, i = oret.cos_topk(idx, q, 4)
    ref = torch.nn.functional.cosine_similarity(q[:, None, :], idx[None, :, :], dim=-1)
    np.testing.assert_allclose(s.numpy(), ref.topk(4, dim=1).values.numpy(), rtol=1e-5, atol=1e-6)
    # scale invariance and tie -> lowest row
    idx2 = torch.cat([idx[:1] * 3.0, idx], 0)
    s2, i2 = oret.cos_topk(idx2, idx[0], 2)
    assert i2[0].tolist() == [0, 1]
    # fewer rows than k
    s3, i3 = oret.cos_topk(idx[:2], q[0], 4)
    assert i3[0, 2:].tolist() == [-1, -1] and torch.isinf(s3[0, 2:]).all()
    # reference selection rule: floor 0, strict >
    assert oret.running_argmax_top1([-0.5, -0.1]) is None
    assert oret.running_argmax_top1([0.2, 0.7, 0.7]) == 1
    assert oret.running_argmax_top1([-1.0], default=0) == 0


def test_dot_assign_formulation_matches_exact_assign_cpu():
    """The large-merge assign kernel ranks clusters by |c|^2 - 2 x.c with every fp32 centroid entering as two
    fp16 terms (hi = fp16(c), lo = fp16(c - hi)).  On CPU, in float64: (a) the hi+lo pair reproduces the
    centroid to ~2^-22 relative (absolute floor 2^-25: fp16 subnormals), (b) the ranking equals the arg-min of sum (x-c)^2 on clustered and on
    structureless fp16 points, including means of several points (not fp16-representable)."""
    import numpy as np
    rng = np.random.default_rng(3)
    n, d, k = 96, 4096, 5
    for clustered in (True, False):
        x = rng.standard_normal((n, d))
        if clustered:
            x = 0.5 * x + 2.0 * rng.standard_normal((k, d))[rng.integers(0, k, n)]
        x = x.astype(np.float16)
        lab0 = rng.integers(0, k, n)
        c = np.stack([x[lab0 == j].astype(np.float32).mean(0) for j in range(k)]).astype(np.float32)
        hi = c.astype(np.float16)
        lo = (c - hi.astype(np.float32)).astype(np.float16)
        rec = hi.astype(np.float64) + lo.astype(np.float64)
        assert np.all(np.abs(rec - c) <= np.maximum(2.0 ** -21 * np.abs(c), 2.0 ** -24))   # 22 bits, fp16-subnormal floor
        xd = x.astype(np.float64)
        exact = ((xd[:, None, :] - c.astype(np.float64)[None]) ** 2).sum(-1)
        score = (rec ** 2).sum(-1)[None, :] - 2.0 * xd @ rec.T
        assert np.array_equal(exact.argmin(1), score.argmin(1))
