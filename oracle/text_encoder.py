"""oracle/text_encoder.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU fp32 restatement of the retrieval text encoder: the reference calls a third-party model,
`AutoModel.from_pretrained(<mxbai-colbert-large-v1>)` (inference_streaming_longva_v2.py:703-705) =
HF `BertModel` (module `transformers`, pinned ==4.37.2 in requirements.txt:149, not vendored under
/root/reference; 5.5.0 in this image), and reads `model(**ids).last_hidden_state[:, 0]`
(utiles.py:704-708, 723-726, 757-760).  Restated from transformers/models/bert/modeling_bert.py
(BertEmbeddings.forward :86-112, BertSelfAttention, BertSelfOutput, BertIntermediate, BertOutput).
Pinned by tests/golden/tiny_bert_ref.npz = outputs of HF BertModel itself run here
(tests/golden/make_golden_bert.py).
"""
import math

import numpy as np


def _ln(x, g, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def _gelu_erf(x):
    erf = np.vectorize(math.erf, otypes=[np.float64])
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def bert_last_hidden_state(W, ids, n_heads, eps=1e-12):
    """W: dict name -> float array (HF BertModel state-dict names).  ids: int sequence [L] (one unpadded
    text, so no attention mask).  Returns last_hidden_state [L, D] float64."""
    f = lambda k: np.asarray(W[k], dtype=np.float64)  # noqa: E731
    ids = np.asarray(ids, dtype=np.int64)
    L = ids.shape[0]
    # BertEmbeddings: (word + token_type[0]) + position, LayerNorm (dropout is identity in eval)
    x = f("embeddings.word_embeddings.weight")[ids] + f("embeddings.token_type_embeddings.weight")[0]
    x = x + f("embeddings.position_embeddings.weight")[np.arange(L)]
    x = _ln(x, f("embeddings.LayerNorm.weight"), f("embeddings.LayerNorm.bias"), eps)
    D = x.shape[1]
    hd = D // n_heads
    i = 0
    while f"encoder.layer.{i}.attention.self.query.weight" in W:
        p = f"encoder.layer.{i}."
        q = x @ f(p + "attention.self.query.weight").T + f(p + "attention.self.query.bias")
        k = x @ f(p + "attention.self.key.weight").T + f(p + "attention.self.key.bias")
        v = x @ f(p + "attention.self.value.weight").T + f(p + "attention.self.value.bias")
        ctx = np.empty_like(q)
        for h in range(n_heads):
            sl = slice(h * hd, (h + 1) * hd)
            s = q[:, sl] @ k[:, sl].T / math.sqrt(hd)
            s = s - s.max(-1, keepdims=True)
            pr = np.exp(s)
            pr /= pr.sum(-1, keepdims=True)
            ctx[:, sl] = pr @ v[:, sl]
        # BertSelfOutput: LayerNorm(dense(ctx) + x)
        y = ctx @ f(p + "attention.output.dense.weight").T + f(p + "attention.output.dense.bias")
        x = _ln(y + x, f(p + "attention.output.LayerNorm.weight"), f(p + "attention.output.LayerNorm.bias"), eps)
        # BertIntermediate (gelu = erf form) + BertOutput: LayerNorm(dense(h) + x)
        h1 = _gelu_erf(x @ f(p + "intermediate.dense.weight").T + f(p + "intermediate.dense.bias"))
        y = h1 @ f(p + "output.dense.weight").T + f(p + "output.dense.bias")
        x = _ln(y + x, f(p + "output.LayerNorm.weight"), f(p + "output.LayerNorm.bias"), eps)
        i += 1
    return x


def cls_embedding(W, ids, n_heads, eps=1e-12):
    """utiles.py:688-690 pooling(strategy='cls'): outputs[:, 0]."""
    return bert_last_hidden_state(W, ids, n_heads, eps)[0]


def sentence_embedding(W, ids, n_heads, eps=1e-12, normalize=True):
    """langchain HuggingFaceEmbeddings / sentence-transformers all-MiniLM-L6-v2 (local_doc_qa.py:185-194): BertModel ->
    Pooling(mean over the attention mask: one unpadded text, so the plain mean) -> Normalize(p=2, eps 1e-12).
    sentence-transformers==2.2.2 is pinned in requirements.txt:126 and absent here: restated from its Pooling /
    Normalize modules; the encoder under it is pinned by tiny_minilm_ref.npz (HF BertModel, 32-wide heads)."""
    h = bert_last_hidden_state(W, ids, n_heads, eps)
    v = h.sum(0) / max(float(h.shape[0]), 1e-9)
    return v / max(float(np.linalg.norm(v)), 1e-12) if normalize else v
