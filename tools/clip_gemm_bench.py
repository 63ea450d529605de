"""The four GEMMs of a CLIP ViT-L/14-336 layer at the rows of one encode pass (frames x 577 tokens) with their real
epilogues, against cuBLAS without any epilogue.   python tools/clip_gemm_bench.py [frames]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200 import _lib as L, ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
m = frames * 577
for name, n, k, act, res in [("qkv", 3072, 1024, L.ACT_NONE, False), ("out", 1024, 1024, L.ACT_NONE, True),
                             ("fc1", 4096, 1024, L.ACT_QUICK_GELU, False), ("fc2", 1024, 4096, L.ACT_NONE, True)]:
    a = (torch.randn(m, k, device=dev) * 0.5).half()
    w = (torch.randn(n, k, device=dev) * 0.05).half()
    b = torch.randn(n, device=dev).half()
    r = torch.randn(m, n, device=dev).half() if res else None
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    med, best = timeit(lambda: ops.linear(a, w, bias=b, residual=r, act=act, out=out), flush=False)
    tm, tb = timeit(lambda: torch.matmul(a, w.t(), out=out), flush=False)
    fl = 2.0 * m * n * k
    print(json.dumps({"gemm": name, "m": m, "n": n, "k": k, "us": round(med * 1e6, 1), "tflops": round(fl / med / 1e12, 1),
                      "cublas_no_epilogue_tflops": round(fl / tm / 1e12, 1), "tma_store": os.environ.get("SCB_LINEAR_TMA_STORE", "1")}))
