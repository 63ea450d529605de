"""Times the two tensor-core stages at the bench shapes: CLIP encode (frames) and Qwen2 prefill (tokens).
   python tools/model_bench.py [frames] [tokens] [--once]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200.longva.model.builder import SYNTHETIC, random_weights  # noqa: E402
from streamchat_b200.qwen2 import Qwen2Engine  # noqa: E402
from streamchat_b200.vision import VisionEncoder  # noqa: E402

dev = torch.device("cuda:0")
cfg = SYNTHETIC["longva7b"]
ncu_mode = len(sys.argv) > 1 and sys.argv[1] == "ncu"          # one encode pass + one prefill, no warm-up (for ncu launch lists)
argv = [a for a in sys.argv[1:] if a != "ncu"]
frames = int(argv[0]) if len(argv) > 0 and argv[0].isdigit() else 100
tokens = int(argv[1]) if len(argv) > 1 and argv[1].isdigit() else 5824
once = "--once" in sys.argv or ncu_mode
W = random_weights(cfg, dev, seed=0)
vis = VisionEncoder(W, cfg.vision, dev)
eng = Qwen2Engine({k: v for k, v in W.items() if not k.startswith("model.vision_tower")}, cfg.llm, dev, max_ctx=32768)
del W


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if frames:
    px = torch.randn(frames, 3, 336, 336, device=dev).half()
    if not once:
        vis.encode_images(px)
    ms = timed(lambda: vis.encode_images(px), 1 if once else 5)
    fl = vis.flops_per_frame() * frames
    print(f"encode_images {frames} frames: {ms:.2f} ms  {frames / ms * 1e3:.0f} frames/s  {fl / ms / 1e9:.0f} TFLOP/s "
          f"({vis.flops_per_frame() / 1e9:.2f} GF/frame)")
if tokens:
    x = (torch.randn(tokens, cfg.llm.hidden_size, device=dev) * 0.02).half()

    def pf():
        eng.reset()
        eng.prefill(x)
    if not once:
        pf()
    ms = timed(pf, 1 if once else 3)
    fl = eng.prefill_flops(tokens)
    print(f"prefill {tokens} tokens: {ms:.2f} ms  {fl / ms / 1e9:.0f} TFLOP/s")

if tokens and "--graph" in sys.argv:
    # same prefill captured in a CUDA graph: what do the host-side launch gaps cost?
    eng.reset()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.reset(); eng.prefill(x)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    eng.reset()
    with torch.cuda.graph(g):
        eng.prefill(x)
    eng.reset()
    ms = timed(lambda: g.replay(), 3)
    print(f"prefill {tokens} tokens, CUDA graph replay: {ms:.2f} ms  {eng.prefill_flops(tokens) / ms / 1e9:.0f} TFLOP/s")
