"""Dialogue-history retrieval (a-R2): CPU checks of the oracle restatement, GPU parity of the
LocalMemoryRetrieval mirror against it on the same injected embeddings."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fakes  # noqa: E402
from oracle import memory_bank as omb


def _memory_json(tmp_path):
    mem = {"User": {"name": "User", "history": {
        "2024-05-01": [{"query": "where did I leave my keys", "response": "on the kitchen table"},
                       {"query": "what colour is the car", "response": "red"},
                       {"query": "who was at the door", "response": "the postman"}],
        "2024-05-02": [{"query": "what did I cook", "response": "pasta with tomato sauce"},
                       {"query": "is the window open", "response": "yes, the left one"}]},
        "summary": {"2024-05-01": "keys, car and a visitor"}},
        "Other": {"name": "Other", "history": {"2024-05-01": [{"query": "x", "response": "y"}]}}}
    p = tmp_path / "memory.json"
    p.write_text(json.dumps(mem, ensure_ascii=False), encoding="utf-8")
    return str(p), mem


def _embed_np(texts, dim=384):
    v = np.stack([fakes.text_vector(t, dim) for t in texts])
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def test_oracle_flat_l2_and_expansion(tmp_path):
    from streamchat_b200.memory_bank import load_memory_file
    path, _ = _memory_json(tmp_path)
    docs = load_memory_file(path, "User", "en")
    assert len(docs) == 6                                    # 3 turns + summary, 2 turns; other users skipped
    assert docs[0].page_content.startswith("Conversation content on 2024-05-01:[|User|]: where did I leave my keys; [|AI|]: on")
    assert docs[3].page_content.startswith("The summary of the conversation on 2024-05-01 is:")
    xb = _embed_np([d.page_content for d in docs])
    s, i = omb.index_flat_l2_search(xb, xb[2:3], 3)
    assert i[0, 0] == 2 and abs(s[0, 0]) < 1e-6 and (np.diff(s[0]) >= 0).all()
    s, i = omb.index_flat_l2_search(xb[:2], xb[2:3], 3)
    assert i[0, 2] == -1
    out, dates = omb.search_memory(docs, xb, xb[4], top_k=1, chunk_size=200)
    assert dates == "2024-05-02" and len(out) == 1 and "pasta" in out[0]


def test_oracle_flat_l2_matches_fixture(golden_dir):
    """oracle.index_flat_l2_search against the committed brute-force fixture of the published IndexFlatL2 definition."""
    g = np.load(os.path.join(golden_dir, "flat_l2_ref.npz"))
    s, i = omb.index_flat_l2_search(g["xb"], g["xq"], int(g["k"]))
    np.testing.assert_array_equal(i, g["labels"])
    np.testing.assert_allclose(s, g["distances"], rtol=1e-4, atol=2e-6)
    s, i = omb.index_flat_l2_search(g["xb"][:2], g["xq"][1:2], 4)
    assert i[0, 2:].tolist() == [-1, -1] and sorted(i[0, :2].tolist()) == [0, 1]


@pytest.mark.gpu
def test_gpu_flat_l2_matches_fixture_and_index_is_append_only(cuda, golden_dir):
    """The device index against the same fixture (labels exact, squared-L2 scores 2 - 2 cos within 2e-6), then the
    append-only behaviour (§8f-4): rows appended in place give the index a fresh build gives, the buffer is not
    reallocated until it is full, and only the new rows are written."""
    from streamchat_b200 import ops
    from streamchat_b200.memory_bank import Document, MemoryIndex
    g = np.load(os.path.join(golden_dir, "flat_l2_ref.npz"))
    xb, xq, k = torch.from_numpy(g["xb"]).to(cuda), torch.from_numpy(g["xq"]).to(cuda), int(g["k"])
    docs = [Document(f"turn {i}", {"source": "2024-05-01"}) for i in range(xb.shape[0])]
    index = MemoryIndex(docs[:200], xb[:200], None)
    assert index.capacity == 256 and index.n == 200
    ptr = index._buf.data_ptr()
    index.append(docs[200:250], xb[200:250])                       # fits: written behind the live rows, same buffer
    assert index._buf.data_ptr() == ptr and index.reallocations == 0 and index.n == 250
    index.append(docs[250:], xb[250:])                             # 257 > 256: one doubling
    assert index.capacity == 512 and index.reallocations == 1 and len(index.docs) == 257
    assert torch.equal(index.embeddings, xb)
    scores, idx = ops.cosine_topk(index.embeddings, xq, k=k)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["labels"])
    np.testing.assert_allclose((2.0 - 2.0 * scores).cpu().numpy(), g["distances"], atol=2e-6)
    with pytest.raises(ValueError):
        index.append(docs[:2], xb[:3])


@pytest.mark.gpu
def test_local_memory_retrieval_matches_oracle(cuda, tmp_path):
    from streamchat_b200.memory_bank import (LocalMemoryRetrieval, build_prompt_with_search_memory_only_related,
                                             enter_name, load_memory_file, only_related_prompt_dict_ego,
                                             save_local_memory)
    path, mem = _memory_json(tmp_path)

    def embed_fn(texts):
        return torch.from_numpy(_embed_np(list(texts))).to(cuda)
    qa = LocalMemoryRetrieval()
    qa.init_cfg("minilm-l6", "cuda:0", 1, "en", embed_fn=embed_fn)
    args = types.SimpleNamespace(memory_basic_dir=str(tmp_path), memory_file="memory.json", language="en")
    msg, user_memory, memory, name, index = enter_name("User", mem, qa, args)
    assert index is not None and msg.startswith("Wellcome Back")
    docs = load_memory_file(path, "User", "en")
    xb = _embed_np([d.page_content for d in docs])
    for top_k in (1, 3):
        qa.top_k = top_k
        for chunk_size in (200, 120, 1000):
            index.chunk_size = chunk_size
            for q in ["what did I cook yesterday", "where are the keys", docs[1].page_content]:
                got = qa.search_memory(q, index)
                want = omb.search_memory(docs, xb, _embed_np([q])[0], top_k, chunk_size)
                assert got == want, (top_k, chunk_size, q, got, want)
    # prompt assembly + a new turn + index rebuild (the per-question cycle of the driver, :918-920)
    qa.top_k, index.chunk_size = 1, 200
    prompt = build_prompt_with_search_memory_only_related("where are the keys", "User", index, qa,
                                                          only_related_prompt_dict_ego()["en"], "[|User|]", "[|AI|]", "AI")
    assert "most relevant historical contextual conversation records" in prompt
    memory = save_local_memory(memory, [["new question", "new answer"]], "User", args)
    _, _, _, _, index2 = enter_name("User", memory, qa, args)
    assert len(index2.docs) == 7                               # 6 documents + the new turn
    assert index2 is index and index2.reallocations == 0        # the resident index was appended to, not rebuilt
    docs2 = load_memory_file(path, "User", "en")
    xb2 = _embed_np([d.page_content for d in docs2])
    assert len(docs2) == 7 and torch.allclose(index2.embeddings.cpu(), torch.from_numpy(xb2), atol=1e-6)
    assert qa.search_memory("new question", index2) == omb.search_memory(docs2, xb2, _embed_np(["new question"])[0], 1, 200)
    # unknown user -> fresh entry, no index
    _, um, _, _, idx = enter_name("Nobody", memory, qa, args)
    assert idx is None and um == {"name": "Nobody"}
    assert build_prompt_with_search_memory_only_related("q", "Nobody", None, qa, "x", "a", "b", "c") is None


def test_dropin_import_paths():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "dropin"))
    try:
        for m in [k for k in sys.modules if k.split(".")[0] in ("longva", "utiles", "torch_kmeans", "memory_bank")]:
            del sys.modules[m]
        from longva.model.builder import load_pretrained_model  # noqa: F401
        from longva.conversation import conv_templates, SeparatorStyle  # noqa: F401
        from longva.mm_utils import tokenizer_image_token  # noqa: F401
        from longva.constants import IMAGE_TOKEN_INDEX
        from torch_kmeans import KMeans  # noqa: F401
        from utiles import (weighted_kmeans_feature, fast_building_memory_tree_summarize_token,  # noqa: F401
                            fast_search_tree_multi_modal_with_embedding, calculate_forgetting_probabilities,
                            select_data_without_replacement, count_nodes_by_depth, MultimodalTreeNode,
                            build_prompt_with_search_memory_only_related, process_images)
        from memory_bank.memory_retrieval.local_doc_qa import LocalMemoryRetrieval  # noqa: F401
        from memory_bank.memory_utils import enter_name, save_local_memory  # noqa: F401
        from memory_bank.prompt_utils import only_related_prompt_dict_ego  # noqa: F401
        assert IMAGE_TOKEN_INDEX == -200 and "qwen_1_5_ego" in conv_templates
    finally:
        sys.path.remove(os.path.join(root, "dropin"))
        for m in [k for k in sys.modules if k.split(".")[0] in ("longva", "utiles", "torch_kmeans", "memory_bank")]:
            del sys.modules[m]


def test_embedding_cache_and_snapshot_cpu(tmp_path):
    """§8f-4: a second build of the index embeds only the turns it has not seen, and the index can be
    snapshotted and reloaded in a fresh retrieval object (no search here: that needs the GPU)."""
    import json

    from streamchat_b200.memory_bank import LocalMemoryRetrieval
    seen = []

    def embed_fn(texts):
        seen.append(list(texts))
        return torch.from_numpy(_embed_np(list(texts)))

    def write(n_turns):
        mem = {"alice": {"name": "alice", "history": {"2024-05-01": [{"query": f"q{i}", "response": f"a{i}"} for i in range(n_turns)]}}}
        p = tmp_path / "mem.json"
        p.write_text(json.dumps(mem))
        return str(p)

    qa = LocalMemoryRetrieval()
    qa._stores, qa._embed_cache = {}, {}
    qa.init_cfg("minilm-l6", "cpu", 1, "en", embed_fn=embed_fn)
    qa.device = torch.device("cpu")
    vs = str(tmp_path / "vs")
    qa.init_memory_vector_store(write(3), vs, "alice")
    n_first = sum(len(t) for t in seen)
    import shutil
    shutil.rmtree(vs)                                   # the driver deletes the directory before every rebuild
    qa.init_memory_vector_store(write(5), vs, "alice")
    n_second = sum(len(t) for t in seen) - n_first
    store = qa.load_memory_index(vs)
    assert n_second == len(store.docs) - n_first > 0    # only the new turns went through the encoder
    assert store.n == 5 and store.capacity == 256 and store.reallocations == 0    # appended in place, not restacked
    path = qa.save_snapshot(vs)
    qb = LocalMemoryRetrieval()
    qb._stores = {}
    qb.init_cfg("minilm-l6", "cpu", 1, "en", embed_fn=embed_fn)
    qb.device = torch.device("cpu")
    again = qb.load_memory_index(vs)
    assert [d.page_content for d in again.docs] == [d.page_content for d in store.docs]
    assert torch.equal(again.embeddings, store.embeddings) and os.path.exists(path)
