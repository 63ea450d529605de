"""Phase breakdown of the persistent decode-step kernel at LongVA-7B shapes (clock64 stamps per CTA).
   python tools/mega_probe.py [ctx] [rows]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200 import _lib as L  # noqa: E402
from streamchat_b200.longva.model.builder import SYNTHETIC, random_weights  # noqa: E402
from streamchat_b200.qwen2 import Qwen2Engine  # noqa: E402

dev = torch.device("cuda:0")
cfg = SYNTHETIC["longva7b"]
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 5800
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
W = random_weights(cfg, dev, seed=0)
W = {k: v for k, v in W.items() if not k.startswith("model.vision_tower") and "mm_projector" not in k}
eng = Qwen2Engine(W, cfg.llm, dev, max_ctx=8192, max_batch=R)
del W
x = (torch.randn(ctx, cfg.llm.hidden_size, device=dev) * 0.02).half()
lib = L.load()
assert lib.scb_qwen2_decoder_profile(eng._mega._h, None, 0) == 0      # arm
eng.generate_batch([x] * R, 6, use_graph=False)
torch.cuda.synchronize()
nl = cfg.llm.num_hidden_layers
n = 148 * nl * 16
buf = (C.c_longlong * n)()
got = lib.scb_qwen2_decoder_profile(eng._mega._h, buf, n)
assert got == n, got
t = np.frombuffer(buf, dtype=np.int64).reshape(148, nl, 16).astype(np.float64)
names = ["P_qkv", "qkv", "sync", "attn1", "sync", "attn2", "sync", "P_o", "o", "sync", "P_gu", "gu", "sync",
         "P_down", "down", "sync(next layer start)"]
d = np.diff(t, axis=2)                                   # [cta, layer, 15]
nxt = t[:, 1:, 0] - t[:, :-1, 15]                        # last sync of layer l
mhz = 1.9e3
print(f"PF={os.environ.get('SCB_MEGA_PF')} BAR={os.environ.get('SCB_MEGA_BAR')} ctx={ctx} rows={R}; cycles -> us at {mhz / 1e3:.1f} GHz; mean over CTAs and layers 1..{nl - 2} (min / max over CTAs)")
tot = 0.0
for i in range(15):
    v = d[:, 1:-1, i].mean(axis=1)
    print(f"  {names[i]:8s} {v.mean() / mhz:7.2f} us   ({v.min() / mhz:6.2f} / {v.max() / mhz:6.2f})")
    tot += v.mean() / mhz
v = nxt[:, 1:].mean(axis=1)
print(f"  {'sync':8s} {v.mean() / mhz:7.2f} us   ({v.min() / mhz:6.2f} / {v.max() / mhz:6.2f})")
tot += v.mean() / mhz
print(f"  layer total {tot:.2f} us  -> x{nl} = {tot * nl / 1e3:.3f} ms")
