"""Per-kernel micro-benchmarks (CUDA events on the launching stream, L2 flushed between timed
launches).  Usage: python tools/kbench.py [linear] [topk] [kmeans] ..."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamchat_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _flush.zero_()


def timeit(fn, iters=10, warmup=3, flush=True):
    """Device time of one call: the call is captured in a CUDA graph (no host gaps inside the timed
    region), L2 flushed by a 256 MB memset before every timed replay."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_linear():
    shapes = [(18464, 3072, 1024), (18464, 1024, 1024), (18464, 4096, 1024), (18464, 1024, 4096),
              (18464, 3584, 1024), (4672, 4608, 3584), (4672, 3584, 3584), (4672, 37888, 3584),
              (4672, 3584, 18944), (8192, 8192, 8192)]
    for m, n, k in shapes:
        a = (torch.randn(m, k, device=dev) * 0.5).half()
        w = (torch.randn(n, k, device=dev) * 0.05).half()
        out = torch.empty(m, n, device=dev, dtype=torch.float16)
        med, best = timeit(lambda: ops.linear(a, w, out=out), flush=False)
        fl = 2.0 * m * n * k
        tm, tb = timeit(lambda: torch.matmul(a, w.t(), out=out), flush=False)
        print(json.dumps({"kernel": "linear", "m": m, "n": n, "k": k, "ms": med * 1e3,
                          "tflops": fl / med / 1e12, "best_tflops": fl / best / 1e12,
                          "cublas_tflops": fl / tm / 1e12}))


def bench_topk():
    for n, d, dt, k in [(10000, 1024, torch.float32, 8), (10000, 1024, torch.float16, 8),
                        (1_000_000, 1024, torch.float32, 8), (1_000_000, 1024, torch.float16, 8)]:
        index = torch.randn(n, d, device=dev, dtype=dt)
        q = torch.randn(d, device=dev, dtype=dt)
        med, best = timeit(lambda: ops.cosine_topk(index, q, k), iters=20)
        by = n * d * index.element_size() + d * index.element_size() + k * 12
        print(json.dumps({"kernel": "cosine_topk", "n": n, "d": d, "dtype": str(dt), "k": k,
                          "us": med * 1e6, "gbs": by / med / 1e9, "best_gbs": by / best / 1e9}))
    # config #3 as a batch of 64 questions in ONE launch (the single-query case is shorter than a
    # launch): the 41 MB index crosses HBM once and serves all queries from L2
    index = torch.randn(10000, 1024, device=dev)
    qs = torch.randn(64, 1024, device=dev)
    med, best = timeit(lambda: ops.cosine_topk(index, qs, 8), iters=20)
    print(json.dumps({"kernel": "cosine_topk", "n": 10000, "d": 1024, "dtype": "torch.float32", "k": 8, "queries": 64,
                      "us": med * 1e6, "us_per_query": med * 1e6 / 64,
                      "index_gbs_hbm": 10000 * 1024 * 4 / med / 1e9, "l2_gbs": 64 * 10000 * 1024 * 4 / med / 1e9}))


def bench_topk_small():
    """Config #3 at its own size (10 000 x 1024) as a retrieval loop would see it: 48 back-to-back searches in ONE graph,
    cycling over 6 index copies (246 MB fp32 > L2, so every search streams its index from HBM) — per-call latency at
    sustained clocks, without the idle-GPU clock drop and event overhead of timing a single 10 us launch."""
    for dt in (torch.float32, torch.float16):
        idxs = [torch.randn(10000, 1024, device=dev, dtype=dt) for _ in range(6 if dt == torch.float32 else 12)]
        q = torch.randn(1024, device=dev, dtype=dt)
        n_calls = 48

        def loop():
            for i in range(n_calls):
                ops.cosine_topk(idxs[i % len(idxs)], q, 8)
        med, best = timeit(loop, iters=10, flush=False)
        by = 10000 * 1024 * idxs[0].element_size()
        print(json.dumps({"kernel": "cosine_topk", "n": 10000, "d": 1024, "dtype": str(dt), "k": 8, "calls_per_graph": n_calls,
                          "us_per_call": med * 1e6 / n_calls, "best_us_per_call": best * 1e6 / n_calls,
                          "gbs": by / (med / n_calls) / 1e9}))


def bench_kmeans():
    T, F, K = 400, 144 * 3584, 5
    x = torch.randn(T, F, device=dev).half()
    cent = x[:K].float().unsqueeze(0).contiguous()
    med, best = timeit(lambda: ops.kmeans_step(x, cent, None, do_update=True), iters=10)
    by = 2 * T * F * 2 + 3 * K * F * 4
    print(json.dumps({"kernel": "kmeans_step", "T": T, "F": F, "K": K, "ms": med * 1e3,
                      "gbs": by / med / 1e9, "best_gbs": by / best / 1e9}))
    med, best = timeit(lambda: ops.kmeans_step(x, cent, None, do_update=False), iters=10)
    by = T * F * 2 + K * F * 4
    print(json.dumps({"kernel": "kmeans_assign_only", "ms": med * 1e3, "gbs": by / med / 1e9}))


def bench_attn():
    import math
    for name, B, T, Hq, Hkv, D, causal in [("qwen2_prefill", 1, 5824, 28, 4, 128, True), ("clip_vit", 64, 577, 16, 16, 64, False)]:
        q = torch.randn(B, T, Hq, D, device=dev).half()
        k = torch.randn(B, T, Hkv, D, device=dev).half()
        v = torch.randn(B, T, Hkv, D, device=dev).half()
        med, best = timeit(lambda: ops.attention(q, k, v, 1.0 / math.sqrt(D), causal), iters=10, flush=False)
        fl = 4.0 * B * Hq * T * T * D * (0.5 if causal else 1.0)
        print(json.dumps({"kernel": "attention", "shape": name, "us": med * 1e6, "tflops": fl / med / 1e12,
                          "variant": os.environ.get("SCB_ATTN128" if D == 128 else "SCB_ATTN64", "default")}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["linear", "topk", "kmeans"]
    for w in which:
        globals()["bench_" + w]()
