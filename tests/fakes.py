"""Deterministic stand-ins for the captioner, tokenizer and text embedder, shared by the golden
generator (driving the reference's code) and the tests (driving the product)."""
import hashlib

import numpy as np
import torch


class _Enc(dict):
    __getattr__ = dict.get

    def items(self):
        return dict.items(self)


class FakeTokenizer:
    bos_token_id = None

    def __call__(self, text, padding=False, return_tensors=None, **kw):
        ids = [3 + int(hashlib.md5(w.encode()).hexdigest()[:6], 16) % 900 for w in text.split()] or [3]
        if return_tensors == "pt":
            return _Enc(input_ids=torch.tensor([ids], dtype=torch.long),
                        attention_mask=torch.ones((1, len(ids)), dtype=torch.long))
        return _Enc(input_ids=ids)

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(f"t{int(t)}" for t in row) for row in ids.tolist()]


class FakeSummarizer:
    """generate_with_image_embedding -> ids that depend only on the call count and input sizes."""

    class _Cfg:
        mm_use_im_start_end = False

    def __init__(self, device):
        self.device = torch.device(device)
        self.config = self._Cfg()
        self.calls = 0

    def generate_with_image_embedding(self, inputs, image_embeddings=None, modalities=None, **kw):
        self.calls += 1
        rows = 0 if image_embeddings is None else int(image_embeddings[0].shape[0])
        n = int(kw.get("max_new_tokens", 8))
        base = (self.calls * 37 + rows * 11 + int(inputs.shape[-1])) % 800
        return torch.tensor([[(base + 7 * i) % 900 for i in range(min(n, 6))]], dtype=torch.long)


def text_vector(text: str, dim: int) -> np.ndarray:
    seed = int(hashlib.sha256(text.encode()).hexdigest()[:8], 16)
    return np.random.RandomState(seed).randn(dim).astype(np.float32)


class FakeBert:
    """`model(**ids).last_hidden_state` contract of the BERT embedder (utiles.py:707): the CLS
    row is a deterministic function of the token ids."""

    def __init__(self, dim: int, device="cpu"):
        self.dim = dim
        self.device = torch.device(device)

    def parameters(self):
        yield torch.zeros(1, device=self.device)

    def forward(self, **kw):
        return self(**kw)

    def __call__(self, input_ids=None, attention_mask=None, **kw):
        key = " ".join(str(int(t)) for t in input_ids[0].tolist())
        v = torch.from_numpy(text_vector(key, self.dim)).to(input_ids.device)

        class _Out:
            pass
        o = _Out()
        o.last_hidden_state = v.view(1, 1, -1).repeat(1, input_ids.shape[1], 1)
        return o
