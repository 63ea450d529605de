"""Sweep of the decode GEMV launch shape (env SCB_GEMV_W = warps per block, SCB_GEMV_C = K chunks)
over the four Qwen2-7B layer shapes, 28 distinct weight matrices per shape (no L2 reuse), timed as a
CUDA graph of 28 launches.   python tools/gemv_sweep.py R"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = {"qkv": (4608, 3584), "o": (3584, 3584), "gate_up": (37888, 3584), "down": (3584, 18944)}
for name, (n, k) in shapes.items():
    ws = [(torch.randn(n, k, device=dev) * 0.02).half() for _ in range(28)]
    a = torch.randn(R, k, device=dev).half()
    out = torch.empty(R, n, device=dev, dtype=torch.float16)
    for w in ws[:2]:
        ops.gemv(a, w, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            ops.gemv(a, w, out=out)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 280
    print(f"W={os.environ.get('SCB_GEMV_W', '4')} C={os.environ.get('SCB_GEMV_C', 'auto')} PDL={os.environ.get('SCB_PDL', '1')} "
          f"R={R} {name:8s} {us:7.2f} us  {n * k * 2 / us / 1e6:6.0f} GB/s")
    del ws
