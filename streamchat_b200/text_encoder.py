"""BERT text encoder for memory retrieval on the sm_100a kernels.

The reference embeds the question and every visited node text with a BERT-large model
(`AutoModel.from_pretrained(mxbai-colbert-large-v1)`, inference_streaming_longva_v2.py:703-705) and
takes `model(**ids).last_hidden_state[:, 0]` (utiles.py:704-708, 723-726, 757-760) — one batch-1
forward plus a device->host copy per text, repeated at every question.  Here the encoder is a
callable with the same surface (`model(input_ids=..., attention_mask=...) -> .last_hidden_state`) and a
packed entry point: `encode_texts` runs the linear layers once over the concatenated tokens of all
texts, attention per text, and returns the CLS rows without leaving the device; memory nodes cache
their embedding (`memory._node_embedding`), so a question costs one forward of the query text plus one
`cosine_topk` per tree level.

Operator order = HF `BertModel` (modeling_bert.py): embeddings (word + token_type[0] + position -> LN,
eps 1e-12) then per layer  x = LN(x + O(Attn(QKV(x)))) ; x = LN(x + W2 gelu_erf(W1 x)).  fp16 storage,
fp32 accumulation and statistics (the reference runs the encoder in fp32: the parity bar on the
embedding is 1e-2 of its scale, and exact agreement of the arg-max the retrieval takes).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import ops


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12


class BertTextEncoder:
    """`weights`: HF `BertModel` state dict (names `embeddings.*`, `encoder.layer.N.*`; a leading `bert.` is
    accepted).  The pooler is not used by the reference (`last_hidden_state`), so it is not loaded."""

    def __init__(self, weights: Dict[str, torch.Tensor], cfg: BertConfig, device):
        self.cfg = cfg
        self.device = torch.device(device)
        D = cfg.hidden_size
        H = cfg.num_attention_heads
        hd = D // H
        # The tcgen05 attention kernel takes head dims 64 / 128.  Narrower heads (MiniLM-L6: 384 / 12 = 32) run as
        # 64-wide heads whose extra dims are ZERO: q/k/v rows of each head are padded with zero rows (and zero biases),
        # the output projection with zero columns — scores, probabilities and outputs are exactly those of the narrow
        # head (a zero contributes exactly 0 to every fp32 accumulation); only the softmax scale keeps the true width.
        if hd in (64, 128):
            self.hd_pad = hd
        elif hd < 64 and hd % 8 == 0:
            self.hd_pad = 64
        else:
            raise ValueError(f"BERT head_dim {hd} unsupported by the attention kernel (<= 64 or 128)")
        self.head_dim = hd
        weights = {(k[5:] if k.startswith("bert.") else k): v for k, v in weights.items()}

        def w(name):
            return weights[name].to(device=self.device, dtype=torch.float16).contiguous()

        def pad_rows(t):        # [H*hd, ...] -> [H*hd_pad, ...], every head's block followed by zero rows
            if self.hd_pad == hd:
                return t
            t = t.view(H, hd, *t.shape[1:])
            z = torch.zeros((H, self.hd_pad - hd, *t.shape[2:]), dtype=t.dtype, device=t.device)
            return torch.cat([t, z], 1).reshape(H * self.hd_pad, *t.shape[2:]).contiguous()

        def pad_cols(t):        # [D, H*hd] -> [D, H*hd_pad]
            if self.hd_pad == hd:
                return t
            return pad_rows(t.t().contiguous()).t().contiguous()

        self.word = w("embeddings.word_embeddings.weight")
        self.pos = w("embeddings.position_embeddings.weight")
        self.type0 = w("embeddings.token_type_embeddings.weight")[0].contiguous()
        self.emb_g, self.emb_b = w("embeddings.LayerNorm.weight"), w("embeddings.LayerNorm.bias")
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"encoder.layer.{i}."
            self.layers.append(dict(
                qkv_w=torch.cat([pad_rows(w(p + "attention.self.query.weight")), pad_rows(w(p + "attention.self.key.weight")),
                                 pad_rows(w(p + "attention.self.value.weight"))], 0).contiguous(),
                qkv_b=torch.cat([pad_rows(w(p + "attention.self.query.bias")), pad_rows(w(p + "attention.self.key.bias")),
                                 pad_rows(w(p + "attention.self.value.bias"))], 0).contiguous(),
                o_w=pad_cols(w(p + "attention.output.dense.weight")), o_b=w(p + "attention.output.dense.bias"),
                ln1_g=w(p + "attention.output.LayerNorm.weight"), ln1_b=w(p + "attention.output.LayerNorm.bias"),
                fc1_w=w(p + "intermediate.dense.weight"), fc1_b=w(p + "intermediate.dense.bias"),
                fc2_w=w(p + "output.dense.weight"), fc2_b=w(p + "output.dense.bias"),
                ln2_g=w(p + "output.LayerNorm.weight"), ln2_b=w(p + "output.LayerNorm.bias")))

    def parameters(self):
        yield self.word

    def flops(self, lengths: Sequence[int]) -> float:
        c = self.cfg
        D, I = c.hidden_size, c.intermediate_size
        T = sum(lengths)
        return c.num_hidden_layers * (2 * T * D * (4 * D + 2 * I) + sum(4 * n * n * D for n in lengths))

    @torch.no_grad()
    def hidden_states(self, id_lists: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Packed forward: `id_lists[i]` int64 [L_i] (special tokens included) -> last hidden state
        [L_i, D] fp16 per text.  No padding exists on this path, hence no attention mask."""
        c = self.cfg
        D, H = c.hidden_size, c.num_attention_heads
        hd = self.hd_pad                                  # width of a head in the qkv / context buffers
        lens = [int(t.numel()) for t in id_lists]
        if not lens or min(lens) < 1:
            raise ValueError("every text needs at least one token")
        if max(lens) > c.max_position_embeddings:
            raise ValueError(f"text of {max(lens)} tokens exceeds max_position_embeddings={c.max_position_embeddings}")
        ids = torch.cat([t.reshape(-1).to(torch.int64) for t in id_lists]).to(self.device)
        pos = torch.cat([torch.arange(n, dtype=torch.int64) for n in lens]).to(self.device)
        x = ops.text_embed_layernorm(ids, pos, self.word, self.pos, self.type0, self.emb_g, self.emb_b,
                                     c.layer_norm_eps)
        T = x.shape[0]
        offs = [0]
        for n in lens:
            offs.append(offs[-1] + n)
        scale = self.head_dim ** -0.5                     # the TRUE head width (padded dims are zero)
        ctx = torch.empty((T, H, hd), dtype=torch.float16, device=self.device)
        for ly in self.layers:
            qkv = ops.linear(x, ly["qkv_w"], bias=ly["qkv_b"]).view(T, 3, H, hd)
            for i, n in enumerate(lens):
                s = qkv[offs[i]:offs[i + 1]]
                ops.attention(s[None, :, 0], s[None, :, 1], s[None, :, 2], scale, causal=False,
                              out=ctx[None, offs[i]:offs[i + 1]])
            y = ops.linear(ctx.view(T, H * hd), ly["o_w"], bias=ly["o_b"], residual=x)
            x = ops.layernorm(y, ly["ln1_g"], ly["ln1_b"], c.layer_norm_eps)
            y = ops.linear(x, ly["fc1_w"], bias=ly["fc1_b"], act=L.ACT_GELU_ERF)
            y = ops.linear(y, ly["fc2_w"], bias=ly["fc2_b"], residual=x)
            x = ops.layernorm(y, ly["ln2_g"], ly["ln2_b"], c.layer_norm_eps)
        self._last_packed = (x, offs)
        return [x[offs[i]:offs[i + 1]] for i in range(len(lens))]

    @torch.no_grad()
    def encode_ids(self, id_lists: Sequence[torch.Tensor]) -> torch.Tensor:
        """CLS embeddings [n, D] fp32 (device) of already tokenised texts."""
        hs = self.hidden_states(id_lists)
        return torch.stack([h[0] for h in hs]).float()

    def encode_texts(self, tokenizer, texts: Sequence[str]) -> torch.Tensor:
        """CLS embeddings [n, D] fp32 of `texts`, one packed forward (tokenizer = the HF tokenizer the
        reference passes beside the model, utiles.py:704)."""
        id_lists = []
        for t in texts:
            enc = tokenizer(t, padding=True, return_tensors="pt")
            id_lists.append(enc["input_ids"][0][: self.cfg.max_position_embeddings])
        return self.encode_ids(id_lists)

    # ---- the reference's call surface: model(**tokenizer(text, padding=True, return_tensors='pt'))
    def __call__(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 token_type_ids: Optional[torch.Tensor] = None, **_):
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, L]")
        if token_type_ids is not None and bool((token_type_ids != 0).any()):
            raise L.ScbError("token_type_ids other than 0 are not on the retrieval path")
        rows = []
        for b in range(input_ids.shape[0]):
            n = int(attention_mask[b].sum()) if attention_mask is not None else input_ids.shape[1]
            if attention_mask is not None and not bool(attention_mask[b, :n].all()):
                raise L.ScbError("only right-padded batches are supported")
            rows.append(input_ids[b, :n])
        hs = self.hidden_states(rows)
        Lmax = input_ids.shape[1]
        out = torch.zeros((len(rows), Lmax, self.cfg.hidden_size), dtype=torch.float32, device=self.device)
        for b, h in enumerate(hs):
            out[b, : h.shape[0]] = h.float()    # padded positions stay 0 (the reference never reads them: CLS pooling)
        return SimpleNamespace(last_hidden_state=out)

    forward = __call__


MINILM_L6 = dict(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536,
                 max_position_embeddings=512, layer_norm_eps=1e-12)


class SentenceEncoder:
    """Dialogue-memory sentence encoder on the sm_100a kernels: what langchain's `HuggingFaceEmbeddings(model_name=
    "sentence-transformers/all-MiniLM-L6-v2")` computes for the reference (memory_bank/memory_retrieval/
    local_doc_qa.py:185-194, configs/model_config.py:5-18) — sentence-transformers 2.2.2 = HF BertModel -> Pooling(mean
    over the attention mask) -> Normalize(p=2); texts truncated to `max_seq_length` (256 for all-MiniLM-L6-v2).
    MiniLM-L6 is a 6-layer BERT with 384 / 12 = 32-wide heads: the encoder runs them as zero-padded 64-wide heads
    (BertTextEncoder).  Callable as `embed_fn(list[str]) -> [n, D] fp32 unit rows` (memory_bank.LocalMemoryRetrieval)."""

    def __init__(self, weights: Dict[str, torch.Tensor], cfg: BertConfig, device, tokenizer=None, max_seq_length: int = 256,
                 normalize: bool = True):
        self.encoder = BertTextEncoder(weights, cfg, device)
        self.device = self.encoder.device
        self.tokenizer = tokenizer
        self.max_seq_length = min(max_seq_length, cfg.max_position_embeddings)
        self.normalize = normalize
        self.dim = cfg.hidden_size

    @classmethod
    def from_pretrained(cls, path: str, device, **kw) -> "SentenceEncoder":
        """A local HF / sentence-transformers directory: config.json + model.safetensors (or pytorch_model.bin) +
        tokenizer files (+ sentence_bert_config.json for max_seq_length)."""
        import json
        import os

        from transformers import AutoTokenizer
        c = json.load(open(os.path.join(path, "config.json")))
        cfg = BertConfig(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                         num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                         max_position_embeddings=c["max_position_embeddings"], layer_norm_eps=c.get("layer_norm_eps", 1e-12))
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            weights = load_file(st)
        else:
            weights = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        sb = os.path.join(path, "sentence_bert_config.json")
        if os.path.exists(sb) and "max_seq_length" not in kw:
            kw["max_seq_length"] = json.load(open(sb)).get("max_seq_length", 256)
        return cls(weights, cfg, device, tokenizer=AutoTokenizer.from_pretrained(path), **kw)

    @torch.no_grad()
    def encode_ids(self, id_lists: Sequence[torch.Tensor]) -> torch.Tensor:
        """[n, D] fp32 sentence embeddings of tokenised texts (special tokens included), one packed forward."""
        self.encoder.hidden_states([t[: self.max_seq_length] for t in id_lists])
        x, offs = self.encoder._last_packed
        return ops.mean_pool_normalize(x, torch.tensor(offs, dtype=torch.int32, device=self.device), self.normalize)

    def encode_texts(self, tokenizer, texts: Sequence[str]) -> torch.Tensor:
        ids = []
        for t in texts:
            enc = tokenizer(t, truncation=True, max_length=self.max_seq_length, return_tensors="pt")
            ids.append(enc["input_ids"][0])
        return self.encode_ids(ids)

    def __call__(self, texts: Sequence[str]) -> torch.Tensor:
        if self.tokenizer is None:
            raise L.ScbError("SentenceEncoder: no tokenizer (construct with tokenizer= or from_pretrained)")
        return self.encode_texts(self.tokenizer, list(texts))


def random_bert_weights(cfg: BertConfig, device, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random-init state dict of the BertModel shape (synthetic benchmarks: there is no checkpoint offline):
    normal(0, std) matrices and embeddings, zero biases, LayerNorm weight 1 / bias 0, fp16."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, I = cfg.hidden_size, cfg.intermediate_size

    def rnd(*shape):
        return (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).half()

    def zeros(n):
        return torch.zeros((n,), dtype=torch.float16, device=device)

    def ones(n):
        return torch.ones((n,), dtype=torch.float16, device=device)

    W = {"embeddings.word_embeddings.weight": rnd(cfg.vocab_size, D),
         "embeddings.position_embeddings.weight": rnd(cfg.max_position_embeddings, D),
         "embeddings.token_type_embeddings.weight": rnd(2, D),
         "embeddings.LayerNorm.weight": ones(D), "embeddings.LayerNorm.bias": zeros(D)}
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for name in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            W[p + name + ".weight"], W[p + name + ".bias"] = rnd(D, D), zeros(D)
        W[p + "intermediate.dense.weight"], W[p + "intermediate.dense.bias"] = rnd(I, D), zeros(I)
        W[p + "output.dense.weight"], W[p + "output.dense.bias"] = rnd(D, I), zeros(D)
        for name in ("attention.output.LayerNorm", "output.LayerNorm"):
            W[p + name + ".weight"], W[p + name + ".bias"] = ones(D), zeros(D)
    return W
