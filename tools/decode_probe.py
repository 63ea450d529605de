"""Short LongVA-7B-shaped run for ncu: one small prefill + a few eager decode steps.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/decode_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200.longva.model.builder import SYNTHETIC, random_weights  # noqa: E402
from streamchat_b200.qwen2 import Qwen2Engine  # noqa: E402

dev = torch.device("cuda:0")
cfg = SYNTHETIC["longva7b"]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
W = random_weights(cfg, dev, seed=0)
W = {k: v for k, v in W.items() if not k.startswith("model.vision_tower") and "mm_projector" not in k}
eng = Qwen2Engine(W, cfg.llm, dev, max_ctx=8192, max_batch=R)
del W
x = (torch.randn(ctx, cfg.llm.hidden_size, device=dev) * 0.02).half()
out = eng.generate_batch([x] * R, steps + 1, use_graph=False)[0]
torch.cuda.synchronize()
print("ids", out.tolist())
