"""GPU: memory-tree build / search mirrors vs golden produced by the reference's own functions
(tests/golden/make_golden_tree.py), and the end-to-end StreamingSession on the tiny model."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fakes  # noqa: E402

pytestmark = pytest.mark.gpu


def _collect(nodes, prefix, out):
    for i, nd in enumerate(nodes):
        key = f"{prefix}{i}"
        out[key] = nd
        _collect(nd.children, key + "c", out)


def test_tree_build_and_search_match_reference(cuda, golden_dir):
    from streamchat_b200 import memory as M
    from streamchat_b200.longva.conversation import conv_templates
    g = np.load(os.path.join(golden_dir, "tree_ref.npz"))
    P, D, chunk, K, interval = [int(v) for v in g["params"]]
    feats_all = torch.from_numpy(g["features"]).to(cuda)
    summarizer, tok = fakes.FakeSummarizer(cuda), fakes.FakeTokenizer()
    input_ids = torch.tensor([[1, 2, -200, 3]])
    tree, off = None, 0
    for si, n in enumerate(g["seg_sizes"].tolist()):
        feats = [feats_all[off + i:off + i + 1] for i in range(n)]
        off += n
        chunks = [feats[i:i + chunk] for i in range(0, n, chunk)]
        kchunks = [torch.cat(c) for c in chunks]
        tree = M.fast_building_memory_tree_summarize_token(
            kchunks, K, interval, summarizer, input_ids, tok, chunks, tree,
            caption_do_sample=True, kmeans_init_indices=g["kmeans_inits"][si].tolist(),
            conv_templates=conv_templates)
    assert len(tree) == int(g["n_roots"])
    nodes = {}
    _collect(tree, "n", nodes)
    keys = sorted(k[:-6] for k in g.files if k.endswith("_depth"))
    assert sorted(nodes) == keys
    for k in keys:
        nd = nodes[k]
        assert nd.depth == int(g[k + "_depth"]) and len(nd.children) == int(g[k + "_nchildren"]), k
        assert nd.text == str(g[k + "_text"]), k
        np.testing.assert_allclose(nd.centroids.float().cpu().numpy(), g[k + "_centroids"], rtol=1e-5, atol=1e-5,
                                   err_msg=k)
    assert dict(M.count_nodes_by_depth(tree)) == {1: 2, 0: 11}
    emb_model, emb_tok = fakes.FakeBert(64, cuda), fakes.FakeTokenizer()
    for qi, q in enumerate(["where are my keys", "what colour was the car", "t12 t7 t7"]):
        feats, texts = M.fast_search_tree_multi_modal_with_embedding(tree, q, None, emb_model, emb_tok)
        assert texts == g[f"q{qi}_texts"].tolist()
        assert [list(f.shape) for f in feats] == g[f"q{qi}_shapes"].tolist()
        np.testing.assert_allclose([float(f.double().sum()) for f in feats], g[f"q{qi}_sums"], rtol=1e-4, atol=1e-3)


def test_search_no_positive_similarity_raises_like_reference(cuda):
    """utiles.py:746: children[None] -> TypeError when no child has a positive similarity."""
    from streamchat_b200 import memory as M
    root = M.MultimodalTreeNode(torch.zeros(1, 1, 8), "root", depth=1)
    child = M.MultimodalTreeNode(torch.zeros(1, 1, 8), "child", depth=0)
    root.children.append(child)
    q = torch.ones(8, device=cuda)
    child.text_embedding = -q

    def emb(tokenizer, text):
        return q
    with pytest.raises(TypeError):
        M.fast_search_tree_multi_modal_with_embedding([root], "q", None, emb, None)
    # depth-0 roots fall back to index 0 (utiles.py:751)
    feats, texts = M.fast_search_tree_multi_modal_with_embedding([child], "q", None, emb, None)
    assert texts == ["child"]


def test_streaming_session_tiny_end_to_end(cuda):
    from streamchat_b200.longva.model.builder import load_pretrained_model
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession
    import streamchat_b200.memory as M
    tokenizer, model, image_processor, ctx = load_pretrained_model("synthetic:tiny", None, "llava_qwen",
                                                                   device_map="cuda:0")
    M.CAPTION_TOKENS, M.SUMMARY_TOKENS = 8, 8
    try:
        cfg = StreamConfig(chunk_size=8, interval=3, answer_tokens=10)
        g = torch.Generator(device=cuda).manual_seed(0)
        frames = torch.randn(60, 3, 56, 56, device=cuda, dtype=torch.float16, generator=g)
        outs = []
        texts = []
        for lazy in (True, False, True):
            cfg.lazy_text = lazy
            np.random.seed(7)
            sess = StreamingSession(model, tokenizer, HashTextEmbedder(1024, cuda), None, None, cfg)
            for s0 in range(0, 60, 20):
                feats = sess.ingest(frames[s0:s0 + 20])
                assert len(feats) == 20 and feats[0].shape == (1, 4, 256)
                tree, short = sess.update_memory(feats, kmeans_init_indices=[0, 1, 2, 3, 4])
                assert len(short) == 5
            depths = dict(M.count_nodes_by_depth(sess.long_memory_tree))
            assert sum(depths.values()) >= 9 and max(depths) >= 1      # at least one merge happened
            text, ids = sess.answer("what did I see?")
            assert ids.shape == (1, 10)
            outs.append(ids.cpu())
            flat = {}
            _collect(sess.long_memory_tree, "n", flat)
            texts.append({k: v.text for k, v in flat.items()})
        # deferred + batched memory texts == one generation per chunk, and runs are deterministic
        assert texts[0] == texts[1] == texts[2]
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    finally:
        M.CAPTION_TOKENS, M.SUMMARY_TOKENS = 128, 256
