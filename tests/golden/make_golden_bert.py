"""Golden vectors for the retrieval text encoder, produced by HF `BertModel` itself (the class the
reference's `AutoModel.from_pretrained(<BERT-large checkpoint>)` resolves to,
inference_streaming_longva_v2.py:703-705) on CPU in fp32 with fp16-representable random weights:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_bert.py

Writes tiny_bert_ref.npz: weights (fp16), three unpadded id lists and their last_hidden_state, the
same texts as one right-padded batch (what `tokenizer([..], padding=True)` produces) with its output.
"""
import os

import numpy as np
import torch
import transformers
from transformers import BertConfig, BertModel

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
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
        p.copy_(p.half().float())                     # fp16-representable: both sides start from equal weights

g = torch.Generator().manual_seed(1)
lens = [5, 12, 9]
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
    out["batch_ids"], out["batch_mask"] = batch.numpy(), mask.numpy()
    out["batch_cls"] = model(input_ids=batch, attention_mask=mask).last_hidden_state[:, 0].numpy()
for k, v in model.state_dict().items():
    out["w:" + k] = v.half().numpy()
np.savez_compressed(os.path.join(HERE, "tiny_bert_ref.npz"), **out)
print("wrote tiny_bert_ref.npz", {k: v.shape for k, v in out.items() if not k.startswith("w:")})
