"""Generates the golden fixtures in this directory by running the REFERENCE's own code
(/root/reference, read-only) on CPU in the build container.  The GPU box has no /root/reference,
so the outputs are committed; this script is the provenance record.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes: kmeans_ref.npz (utiles.weighted_kmeans_feature), torch_kmeans_ref.npz (vendored
torch_kmeans.KMeans), memory_ref.npz (forgetting-curve helpers).  Model fixtures (ViT / Qwen2)
are produced by make_golden_models.py.
"""
import ast
import math
import os
import random
import sys
import time
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def extract(names):
    """AST-extract top-level defs from utiles.py (importing it needs absent packages)."""
    src = open(os.path.join(REF, "utiles.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "np": np, "math": math, "random": random, "time": time,
          "defaultdict": defaultdict}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), "utiles.py", "exec"), ns)
    return ns


def kmeans_cases():
    ns = extract({"weighted_kmeans_feature"})
    ref = ns["weighted_kmeans_feature"]
    out = {}
    cases = [("a", 64, 1, 1024, 5, 0), ("b", 50, 9, 64, 5, 1), ("c", 40, 4, 100, 3, 2),
             ("d", 17, 3, 7, 4, 3), ("e", 4, 2, 8, 5, 4)]
    for name, T, P, D, K, seed in cases:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(T, P, D, generator=g)
        torch.manual_seed(100 + seed)
        init = torch.randperm(T)[:K].clone()
        torch.manual_seed(100 + seed)       # the reference draws the same permutation (utiles.py:295)
        random.seed(seed)
        res = ref(x, K)
        out[f"{name}_x"] = x.numpy()
        out[f"{name}_k"] = np.int64(K)
        out[f"{name}_init"] = init.numpy()
        out[f"{name}_arity"] = np.int64(len(res))
        if len(res) == 2:
            out[f"{name}_centroids"] = res[0].numpy()
            out[f"{name}_labels"] = res[1].numpy()
    # clustered data -> converges before max_iter and exercises the early break
    g = torch.Generator().manual_seed(7)
    centers = torch.randn(4, 32, generator=g) * 10
    x = (centers[torch.arange(60) % 4] + 0.1 * torch.randn(60, 32, generator=g)).view(60, 2, 16)
    torch.manual_seed(107)
    init = torch.randperm(60)[:4].clone()
    torch.manual_seed(107)
    random.seed(7)
    res = ref(x, 4)
    out["f_x"], out["f_k"], out["f_init"] = x.numpy(), np.int64(4), init.numpy()
    out["f_arity"] = np.int64(2)
    out["f_centroids"], out["f_labels"] = res[0].numpy(), res[1].numpy()
    np.savez_compressed(os.path.join(HERE, "kmeans_ref.npz"), **out)
    print("kmeans_ref.npz", {k: v.shape for k, v in out.items() if k.endswith("labels")})


def torch_kmeans_cases():
    sys.path.insert(0, REF)
    from torch_kmeans import KMeans
    out = {}
    for name, bs, n, d, k, seed in [("cfg1", 1, 64, 1024, 5, 0), ("small", 2, 40, 16, 3, 1)]:
        torch.manual_seed(seed)
        x = torch.randn(bs, n, d)
        model = KMeans(n_clusters=k, verbose=False)
        kk = model._check_k(None, dims=x.shape)
        model._check_x(x)
        init = model._init_rnd(x, kk).clone()
        res = model(x)
        out[f"{name}_x"] = x.numpy()
        out[f"{name}_k"] = np.int64(k)
        out[f"{name}_init"] = init.numpy()
        out[f"{name}_labels"] = res.labels.numpy()
        out[f"{name}_centers"] = res.centers.numpy()
        out[f"{name}_inertia"] = res.inertia.numpy()
    np.savez_compressed(os.path.join(HERE, "torch_kmeans_ref.npz"), **out)
    print("torch_kmeans_ref.npz", out["cfg1_labels"].shape, out["cfg1_inertia"])


def memory_cases():
    ns = extract({"calculate_forgetting_probabilities", "select_data_without_replacement",
                  "count_nodes_by_depth", "MultimodalTreeNode"})
    out = {"p_20_5": ns["calculate_forgetting_probabilities"](20, tau=5),
           "p_7_10": ns["calculate_forgetting_probabilities"](7),
           "p_1_5": ns["calculate_forgetting_probabilities"](1, tau=5)}
    np.random.seed(42)
    out["sel_20_5_seed42"] = np.array(
        ns["select_data_without_replacement"](list(range(20)), out["p_20_5"], 5))
    np.savez_compressed(os.path.join(HERE, "memory_ref.npz"), **out)
    print("memory_ref.npz", out["sel_20_5_seed42"])


if __name__ == "__main__":
    kmeans_cases()
    torch_kmeans_cases()
    memory_cases()
