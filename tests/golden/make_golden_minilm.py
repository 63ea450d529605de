"""Golden vectors for the dialogue-memory sentence encoder (row a-R2 / §8f-1), produced by HF `BertModel` itself — the
class sentence-transformers' all-MiniLM-L6-v2 wraps (memory_bank/memory_retrieval/local_doc_qa.py:185-194 via langchain
HuggingFaceEmbeddings) — on a tiny MiniLM-shaped config (32-wide heads), CPU fp32, fp16-representable random weights:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_minilm.py

The pooling tail is sentence-transformers 2.2.2 (pinned in requirements.txt:126; NOT installed here, restated from its
published modules): Pooling(mean) = sum(h * mask) / clamp(sum(mask), 1e-9); Normalize = F.normalize(p=2, dim=1).
Writes tiny_minilm_ref.npz: weights (fp16), id lists, last_hidden_state per text, sentence embeddings of the
right-padded batch.
"""
import os

import numpy as np
import torch
import transformers
from transformers import BertConfig, BertModel

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(vocab_size=150, hidden_size=96, num_hidden_layers=2, num_attention_heads=3, intermediate_size=192,
           max_position_embeddings=64, layer_norm_eps=1e-12, hidden_act="gelu")

torch.manual_seed(0)
model = BertModel(BertConfig(**CFG), add_pooling_layer=False).eval()
with torch.no_grad():
    for name, p in model.named_parameters():
        if "LayerNorm.weight" in name:
            p.copy_(1.0 + 0.1 * torch.randn_like(p))
        elif "LayerNorm.bias" in name or name.endswith(".bias"):
            p.copy_(0.05 * torch.randn_like(p))
        else:
            p.copy_(0.08 * torch.randn_like(p))
        p.copy_(p.half().float())

g = torch.Generator().manual_seed(2)
lens = [4, 17, 9, 30]
ids = [torch.randint(1, CFG["vocab_size"], (n,), generator=g) for n in lens]
out = {"cfg": np.array([CFG["hidden_size"], CFG["num_hidden_layers"], CFG["num_attention_heads"],
                        CFG["intermediate_size"], CFG["vocab_size"], CFG["max_position_embeddings"]]),
       "transformers_version": np.array(transformers.__version__)}
with torch.no_grad():
    for i, t in enumerate(ids):
        out[f"ids{i}"] = t.numpy()
        out[f"hidden{i}"] = model(input_ids=t[None]).last_hidden_state[0].numpy()
    L = max(lens)
    batch = torch.zeros((len(ids), L), dtype=torch.int64)
    mask = torch.zeros((len(ids), L), dtype=torch.int64)
    for i, t in enumerate(ids):
        batch[i, : t.numel()] = t
        mask[i, : t.numel()] = 1
    h = model(input_ids=batch, attention_mask=mask).last_hidden_state
    m = mask.unsqueeze(-1).float()
    pooled = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
    out["sentence_mean"] = pooled.numpy()
    out["sentence_embedding"] = torch.nn.functional.normalize(pooled, p=2, dim=1).numpy()
for k, v in model.state_dict().items():
    out["w:" + k] = v.half().numpy()
np.savez_compressed(os.path.join(HERE, "tiny_minilm_ref.npz"), **out)
print("wrote tiny_minilm_ref.npz", {k: v.shape for k, v in out.items() if not k.startswith("w:")})
