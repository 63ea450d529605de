"""Decode-step timing on the LongVA-7B-shaped engine: CUDA-graph replays for R sequences.
   python tools/decode_bench.py [ctx] [R ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200.longva.model.builder import SYNTHETIC, random_weights  # noqa: E402
from streamchat_b200.qwen2 import Qwen2Engine  # noqa: E402

dev = torch.device("cuda:0")
cfg = SYNTHETIC["longva7b"]
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 5800
Rs = [int(a) for a in sys.argv[2:]] or [1, 5]
W = random_weights(cfg, dev, seed=0)
W = {k: v for k, v in W.items() if not k.startswith("model.vision_tower") and "mm_projector" not in k}
eng = Qwen2Engine(W, cfg.llm, dev, max_ctx=8192, max_batch=max(Rs))
del W
x = (torch.randn(ctx, cfg.llm.hidden_size, device=dev) * 0.02).half()
for R in Rs:
    for use_graph in ((True,) if os.environ.get("DECODE_BENCH_GRAPH_ONLY") else (True, False)):
        steps = 64
        eng.generate_batch([x] * R, 4, use_graph=use_graph)     # warm-up / capture
        torch.cuda.synchronize()
        eng.profile = []
        t0 = time.perf_counter()
        eng.generate_batch([x] * R, steps + 1, use_graph=use_graph)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        p = [q for q in eng.profile if q.get("kind", "decode") == "decode"][0]
        n = p["steps"] if use_graph else steps
        ms = p["events"][0].elapsed_time(p["events"][1])
        by = eng.weight_bytes() + R * eng.kv_bytes_per_token(ctx + steps // 2)
        print(f"R={R} graph={use_graph} ctx={ctx}: {ms / steps:.3f} ms/step  {by / (ms / steps * 1e-3) / 1e9:.0f} GB/s "
              f"(weights {eng.weight_bytes() / 1e9:.2f} GB)  wall incl. prefill {wall:.2f}s PDL={os.environ.get('SCB_PDL', '1')} mega={eng._use_mega(R)}")
        eng.profile = None
