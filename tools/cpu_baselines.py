#!/usr/bin/env python
"""CPU baselines of BASELINE.json's configs, the REFERENCE's own code on the host cores:

    python tools/cpu_baselines.py            -> baseline/results.json

  #1 torch_kmeans.KMeans(n_clusters=5) on [1, 64, 1024] fp32 and utiles.weighted_kmeans_feature on [64, 1, 1024] — full size
  #2 LlavaQwenForCausalLM.encode_images, 32 synthetic 336x336 frames, CLIP ViT-L/14-336 with all 24 layers + projector
     + pool — full size
  #3 cos_sim top-k over a 10 000 x 1024 index (the reference's cos_sim, utiles.py:17 / sentence-transformers util,
     restated: normalise both, mm) — full size
  #4 generate_with_image_embedding: 4 608 visual + 64 text tokens prefill, 8 greedy tokens, Qwen2-7B widths with 2 of
     28 layers, extrapolated to full depth with per-layer timings (oracle/ref_runner.py)
Needs oracle/_ref (python -m oracle.stage_ref).  Every entry: {config, cpu_cores, cpu_seconds, extrapolated, formula}."""
import contextlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref_runner as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cpu_cores": cores, "torch": torch.__version__, "results": []}

    def add(config, seconds, extrapolated=False, formula="measured at full size", **extra):
        out["results"].append(dict(config=config, cpu_cores=cores, cpu_seconds=seconds, extrapolated=extrapolated,
                                   formula=formula, **extra))

    with contextlib.redirect_stdout(sys.stderr):
        # ---- #1
        ts = sorted(R.time_torch_kmeans()["seconds"] for _ in range(5))
        add("#1 torch_kmeans.KMeans(n_clusters=5) on [1,64,1024] fp32 (8 restarts, <=100 iterations)", ts[2], runs=ts)
        fn = R.reference_function("weighted_kmeans_feature")
        torch.manual_seed(0)
        x = torch.randn(64, 1, 1024)
        ts = []
        for _ in range(5):
            torch.manual_seed(1)
            t0 = time.perf_counter()
            fn(x, 5)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        add("#1 utiles.weighted_kmeans_feature on [64,1,1024] fp32, K=5", ts[2], runs=ts)
        # ---- #3
        torch.manual_seed(0)
        index, q = torch.randn(10000, 1024), torch.randn(1, 1024)

        def cos_topk(k):
            a = torch.nn.functional.normalize(index, p=2, dim=1)
            b = torch.nn.functional.normalize(q, p=2, dim=1)
            return torch.topk(torch.mm(b, a.t()), k)
        for k in (1, 8):
            cos_topk(k)
            ts = []
            for _ in range(7):
                t0 = time.perf_counter()
                cos_topk(k)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            add(f"#3 cos_sim (normalise both, mm) + top-{k} over a 10000x1024 fp32 index", ts[3], runs=ts)
        # ---- #2 full size
        model = R.build_reference_model(R.VIT_LAYERS_FULL, 1, llm=dict(R.QWEN2_7B, vocab_size=1024))
        enc = R.time_encode_images(model, 32)
        add("#2 reference encode_images: 32 frames 336x336, CLIP ViT-L/14-336 (24 layers) + mlp2x_gelu + 2x2 pool, fp32",
            enc["seconds"], frames_per_second=32 / enc["seconds"], raw=enc)
        del model
        # ---- #4 reduced depth
        Ll = 2
        model = R.build_reference_model(1, Ll)
        gen = R.time_generate(model, 4608, 64, 9)
        pre = (gen["prefill_seconds"] - gen["prefill_layer_seconds"]) + R.LLM_LAYERS_FULL * gen["prefill_layer_seconds"] / Ll
        tok = (gen["decode_seconds"] - gen["decode_layer_seconds"]) / gen["decode_steps"] + \
            R.LLM_LAYERS_FULL * gen["decode_layer_seconds"] / (Ll * gen["decode_steps"])
        add("#4 reference generate_with_image_embedding: 4608 visual + 64 text tokens prefill, greedy decode, Qwen2-7B widths",
            pre + 127 * tok, extrapolated=True,
            formula=f"measured with {Ll} of 28 layers; full depth = time outside layers + 28 x per-layer time; 128 new tokens = "
                    f"prefill + 127 x per-token time", prefill_seconds=pre, seconds_per_token=tok, raw=gen)
    os.makedirs(os.path.join(ROOT, "baseline"), exist_ok=True)
    path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "baseline", "results.json")
    json.dump(out, open(path, "w"), indent=1)
    for r in out["results"]:
        print(f"{r['cpu_seconds']:10.4f} s  {'(extrapolated) ' if r['extrapolated'] else ''}{r['config']}")
    print("wrote", path)


if __name__ == "__main__":
    main()
