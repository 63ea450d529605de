"""CPU: host-side argument contract of the drop-in `torch_kmeans.KMeans` (no kernel calls) — the checks config #1's
surface relies on (torch_kmeans/clustering/kmeans.py:108-241), in this repo's own wording."""
import warnings

import pytest
import torch

from streamchat_b200.kmeans import KMeans, LpDistance


def test_constructor_contract():
    for bad in (dict(num_init=0), dict(max_iter=0), dict(tol=2), dict(normalize="x"), dict(n_clusters=1), dict(init_method="foo")):
        with pytest.raises(ValueError):
            KMeans(**bad)
    with pytest.raises(NotImplementedError):
        KMeans(p_norm=1)
    km = KMeans(normalize=True, verbose=False)
    assert km.normalize == "mean" and isinstance(km.distance, LpDistance) and not km.is_fitted and km.num_clusters is None


def test_shape_checks_and_center_layout():
    km = KMeans(n_clusters=5, verbose=False)
    x = torch.randn(2, 64, 16)
    assert km._check_x(x[None]).shape == (2, 64, 16)
    with pytest.raises(ValueError):
        km._check_x(torch.randn(64, 16))
    with pytest.raises(TypeError):
        km._check_x([1, 2])
    assert km._check_k(None, x.shape).tolist() == [5, 5] and km._check_k(torch.tensor([7]), x.shape).tolist() == [7, 7]
    for bad in (64, 1):
        with pytest.raises(ValueError):
            km._check_k(bad, x.shape)
    with pytest.raises(NotImplementedError):
        km._check_k(torch.tensor([3, 4]), x.shape)
    km._check_k(5, x.shape)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        c = km._check_centers(torch.randn(2, 5, 16), x.shape)
    assert c.shape == (2, 8, 5, 16) and len(w) == 1 and torch.equal(c[:, 0], c[:, 7])
    assert km._check_centers(torch.randn(2, 8, 5, 16), x.shape).shape == (2, 8, 5, 16)
    with pytest.raises(ValueError):
        km._check_centers(torch.randn(2, 8, 4, 16), x.shape)


def test_normalize_modes_match_their_definitions():
    x = torch.randn(2, 6, 4)
    x[0, 0, 0] = 0
    z = x.clone()
    z[z == 0] = 1e-8
    want = torch.diag_embed(1.0 / torch.norm(z, p=2, dim=-1)) @ z          # upstream formulation (kmeans.py:375-381)
    assert torch.allclose(KMeans._normalize(x.clone(), "unit"), want, atol=1e-6)
    m = KMeans._normalize(x.clone(), "mean")
    assert torch.allclose(m.mean(1), torch.zeros(2, 4), atol=1e-6)
    mm = KMeans._normalize(x.clone(), "minmax")
    assert float(mm.min()) >= 0 and torch.allclose(mm.max(-1).values, torch.ones(2, 6))
