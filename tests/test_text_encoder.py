"""Retrieval text encoder: oracle/text_encoder.py pinned against HF BertModel outputs
(tests/golden/tiny_bert_ref.npz), and the sm_100a encoder against the same golden through the reference's
call surface (`model(**ids).last_hidden_state[:, 0]`, utiles.py:704-708)."""
import os

import numpy as np
import pytest
import torch

from oracle import text_encoder as ot


def _load(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_bert_ref.npz"))
    W = {k[2:]: g[k].astype(np.float32) for k in g.files if k.startswith("w:")}
    return g, W


def test_oracle_matches_hf_bert(golden_dir):
    g, W = _load(golden_dir)
    heads = int(g["cfg"][2])
    for i in range(3):
        out = ot.bert_last_hidden_state(W, g[f"ids{i}"], heads)
        np.testing.assert_allclose(out, g[f"hidden{i}"], rtol=1e-4, atol=2e-5)
    # a right-padded batch with its mask gives the same CLS rows as the unpadded texts
    for i in range(3):
        np.testing.assert_allclose(ot.cls_embedding(W, g[f"ids{i}"], heads), g["batch_cls"][i], rtol=1e-4, atol=3e-5)


def _encoder(golden_dir, cuda):
    from streamchat_b200.text_encoder import BertConfig, BertTextEncoder
    g, W = _load(golden_dir)
    c = g["cfg"]
    cfg = BertConfig(hidden_size=int(c[0]), num_hidden_layers=int(c[1]), num_attention_heads=int(c[2]),
                     intermediate_size=int(c[3]), vocab_size=int(c[4]), max_position_embeddings=int(c[5]))
    return g, BertTextEncoder({k: torch.from_numpy(v) for k, v in W.items()}, cfg, cuda)


@pytest.mark.gpu
def test_gpu_encoder_matches_hf_bert(cuda, golden_dir):
    """fp16 storage / fp32 accumulation against the fp32 reference: 2e-3 relative L2, 5e-3 of the feature scale at
    the worst element (post-LN activations are O(1)); the packed forward equals the one-text-at-a-time forward bit for bit."""
    g, enc = _encoder(golden_dir, cuda)
    ids = [torch.from_numpy(g[f"ids{i}"]) for i in range(3)]
    packed = enc.hidden_states(ids)
    for i, h in enumerate(packed):
        ref = torch.from_numpy(g[f"hidden{i}"])
        assert h.shape == ref.shape
        err = (h.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
        from parity import record, rel_l2
        record(f"tiny_bert_hidden{i}", rel_l2=rel_l2(h.float().cpu(), ref), rel_max=err)
        assert rel_l2(h.float().cpu(), ref) < 2e-3 and err < 5e-3, (rel_l2(h.float().cpu(), ref), err)
        alone = enc.hidden_states([ids[i]])[0]
        assert torch.equal(alone, h)
    # reference call surface: right-padded batch + mask -> last_hidden_state[:, 0]
    out = enc(input_ids=torch.from_numpy(g["batch_ids"]).to(cuda), attention_mask=torch.from_numpy(g["batch_mask"]).to(cuda))
    cls = out.last_hidden_state[:, 0].cpu()
    ref = torch.from_numpy(g["batch_cls"])
    assert (cls - ref).abs().max().item() / ref.abs().max().item() < 5e-3
    assert torch.equal(cls.to(cuda), enc.encode_ids(ids))


@pytest.mark.gpu
def test_gpu_retrieval_argmax_matches_oracle_embeddings(cuda, golden_dir):
    """The quantity retrieval consumes: arg-max of cos(query, node) over node texts (utiles.py:732-736).
    GPU embeddings + cosine_topk pick the node the fp32 oracle picks."""
    from streamchat_b200 import ops
    g, enc = _encoder(golden_dir, cuda)
    _, W = _load(golden_dir)
    heads = int(g["cfg"][2])
    rng = np.random.default_rng(3)
    texts = [rng.integers(1, int(g["cfg"][4]), size=int(n)) for n in rng.integers(4, 40, size=24)]
    query = rng.integers(1, int(g["cfg"][4]), size=11)
    ref = np.stack([ot.cls_embedding(W, t, heads) for t in texts])
    rq = ot.cls_embedding(W, query, heads)
    sims = ref @ rq / (np.linalg.norm(ref, axis=1) * np.linalg.norm(rq))
    order = np.argsort(-sims)
    emb = enc.encode_ids([torch.from_numpy(t) for t in texts])
    q = enc.encode_ids([torch.from_numpy(query)])[0]
    scores, idx = ops.cosine_topk(emb, q, k=1)
    if sims[order[0]] - sims[order[1]] > 5e-3:         # ranking decided above fp16 resolution
        assert int(idx[0, 0]) == int(order[0])
    assert abs(float(scores[0, 0]) - float(sims[int(idx[0, 0])])) < 5e-3


# ---- dialogue-memory sentence encoder (MiniLM shape: 32-wide heads, mean pool + L2 normalise) -----------------------
def _load_minilm(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_minilm_ref.npz"))
    W = {k[2:]: g[k].astype(np.float32) for k in g.files if k.startswith("w:")}
    return g, W


def test_oracle_sentence_embedding_matches_hf_bert_mean_pool(golden_dir):
    g, W = _load_minilm(golden_dir)
    heads = int(g["cfg"][2])
    assert int(g["cfg"][0]) // heads == 32                          # the MiniLM head width
    for i in range(4):
        np.testing.assert_allclose(ot.bert_last_hidden_state(W, g[f"ids{i}"], heads), g[f"hidden{i}"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ot.sentence_embedding(W, g[f"ids{i}"], heads), g["sentence_embedding"][i], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ot.sentence_embedding(W, g[f"ids{i}"], heads, normalize=False), g["sentence_mean"][i],
                                   rtol=1e-4, atol=2e-5)


@pytest.mark.gpu
def test_gpu_sentence_encoder_matches_hf_golden(cuda, golden_dir):
    """32-wide heads run as zero-padded 64-wide heads on the tcgen05 attention kernel; mean pool + L2 normalise in
    scb_mean_pool_normalize.  Against HF BertModel (fp32): hidden states and sentence embeddings within the fp16-execution
    tolerance of tests/parity.py; one packed forward == one text at a time."""
    from parity import record, rel_l2, rel_max
    from streamchat_b200.text_encoder import BertConfig, SentenceEncoder
    g, W = _load_minilm(golden_dir)
    c = g["cfg"]
    cfg = BertConfig(hidden_size=int(c[0]), num_hidden_layers=int(c[1]), num_attention_heads=int(c[2]),
                     intermediate_size=int(c[3]), vocab_size=int(c[4]), max_position_embeddings=int(c[5]))
    enc = SentenceEncoder({k: torch.from_numpy(v) for k, v in W.items()}, cfg, cuda)
    assert enc.encoder.hd_pad == 64 and enc.encoder.head_dim == 32
    ids = [torch.from_numpy(g[f"ids{i}"]) for i in range(4)]
    hs = enc.encoder.hidden_states(ids)
    for i, h in enumerate(hs):
        ref = torch.from_numpy(g[f"hidden{i}"])
        assert rel_l2(h.float().cpu(), ref) < 2e-3, (i, rel_l2(h.float().cpu(), ref))
    emb = enc.encode_ids(ids)
    ref = torch.from_numpy(g["sentence_embedding"])
    record("minilm_sentence_embedding", rel_l2=rel_l2(emb.cpu(), ref), rel_max=rel_max(emb.cpu(), ref))
    assert emb.dtype == torch.float32 and rel_l2(emb.cpu(), ref) < 2e-3
    np.testing.assert_allclose(emb.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    for i in range(4):
        assert torch.equal(enc.encode_ids([ids[i]])[0], emb[i])
    mean = SentenceEncoder({k: torch.from_numpy(v) for k, v in W.items()}, cfg, cuda, normalize=False).encode_ids(ids)
    assert rel_l2(mean.cpu(), torch.from_numpy(g["sentence_mean"])) < 2e-3
