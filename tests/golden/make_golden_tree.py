"""Golden fixture for the memory-tree host logic: runs the REFERENCE's
fast_building_memory_tree_summarize_token (utiles.py:489-620) and
fast_search_tree_multi_modal_with_embedding (utiles.py:685-788), AST-extracted from
/root/reference/utiles.py, on CPU with a deterministic fake summariser / tokenizer / embedder
(tests/fakes.py — the same fakes drive the product in tests/test_gpu_stream.py).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tree.py     -> tree_ref.npz
"""
import ast
import math
import os
import random
import sys
import time
import types
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fakes  # noqa: E402
from oracle.retrieval import cos_sim  # noqa: E402  (restated third-party function, see oracle/retrieval.py)

sys.path.insert(0, REF)
qf = types.ModuleType("longva.model.multimodal_resampler.qformer")
qf.Qformer = object
sys.modules["longva.model.multimodal_resampler.qformer"] = qf
from longva.conversation import conv_templates  # noqa: E402


def extract(names):
    src = open(os.path.join(REF, "utiles.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "np": np, "math": math, "random": random, "time": time,
          "defaultdict": defaultdict, "conv_templates": conv_templates, "cos_sim": cos_sim,
          "RED": "", "BLUE": "", "RESET": ""}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), "utiles.py", "exec"), ns)
    return ns


def main():
    ns = extract({"MultimodalTreeNode", "weighted_kmeans_feature", "fast_building_memory_tree_summarize_token",
                  "fast_search_tree_multi_modal_with_embedding", "count_nodes_by_depth"})
    build = ns["fast_building_memory_tree_summarize_token"]
    search = ns["fast_search_tree_multi_modal_with_embedding"]
    torch.Tensor.cuda = lambda self, *a, **k: self        # the reference calls .cuda() on token ids (utiles.py:706)
    import io, contextlib
    P, D, chunk, K, interval = 4, 32, 6, 3, 4
    g = torch.Generator().manual_seed(5)
    summarizer, tok = fakes.FakeSummarizer("cpu"), fakes.FakeTokenizer()
    input_ids = torch.tensor([[1, 2, -200, 3]])
    tree = None
    out = {}
    inits = []
    all_feats = []
    seg_sizes = [13, 12, 11, 12, 9]     # -> 3+2+2+2+2 chunks; merges when 4 depth-0 nodes are present
    with contextlib.redirect_stdout(io.StringIO()):
        for si, n in enumerate(seg_sizes):
            feats = [torch.randn(1, P, D, generator=g) for _ in range(n)]
            all_feats.append(torch.cat(feats))
            chunks = [feats[i:i + chunk] for i in range(0, n, chunk)]
            kchunks = [torch.cat(c) for c in chunks]
            torch.manual_seed(50 + si)
            # record the permutation the reference will draw IF it merges (utiles.py:295): the merge
            # concatenates `interval` nodes; its row count is only known inside, so draw lazily below
            state = torch.get_rng_state()
            tree = build(kchunks, K, interval, summarizer, input_ids, tok, chunks, tree)
            torch.set_rng_state(state)
            merged = [nd for nd in tree if nd.depth > 0]
            T = sum(c.centroids.shape[0] for c in merged[-1].children) if merged else 0
            inits.append(torch.randperm(T)[:K].numpy() if T else np.zeros(K, dtype=np.int64))
        def dump(nodes, prefix):
            for i, nd in enumerate(nodes):
                key = f"{prefix}{i}"
                out[key + "_depth"] = np.int64(nd.depth)
                out[key + "_text"] = np.array(nd.text)
                out[key + "_centroids"] = nd.centroids.numpy()
                out[key + "_nchildren"] = np.int64(len(nd.children))
                dump(nd.children, key + "c")
        dump(tree, "n")
        out["n_roots"] = np.int64(len(tree))
        emb_model, emb_tok = fakes.FakeBert(64), fakes.FakeTokenizer()
        for qi, q in enumerate(["where are my keys", "what colour was the car", "t12 t7 t7"]):
            feats, texts = search(tree, q, None, emb_model, emb_tok)
            out[f"q{qi}_texts"] = np.array(texts)
            out[f"q{qi}_shapes"] = np.array([list(f.shape) for f in feats])
            out[f"q{qi}_sums"] = np.array([float(f.double().sum()) for f in feats])
    out["seg_sizes"] = np.array(seg_sizes)
    out["features"] = torch.cat(all_feats).numpy()
    out["kmeans_inits"] = np.stack(inits)
    out["params"] = np.array([P, D, chunk, K, interval])
    np.savez_compressed(os.path.join(HERE, "tree_ref.npz"), **out)
    print("roots:", [(int(out[f'n{i}_depth']), int(out[f'n{i}_nchildren'])) for i in range(len(tree))])
    print("q0:", out["q0_texts"], out["q0_shapes"].tolist())


if __name__ == "__main__":
    main()
