#!/usr/bin/env python
"""Tensor-parallel answer decode (streamchat_b200/tp_decode.py) on T = WORLD_SIZE GPUs (2 or 4), under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tools/check_tp_decode.py [--full]

1. small model (3 layers, hidden 512): greedy ids of the TP decode == ids of the single-GPU engine, on every rank;
   the first decode step's logits against the persistent kernel's (relative L2);
2. --full: LongVA-7B shapes (28 layers): ms per decode step, TP vs the persistent one-GPU kernel, ctx 5 800 / 18 000.
Prints one JSON line on rank 0; exit code != 0 on a mismatch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    from streamchat_b200.longva.model.builder import SYNTHETIC, random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig
    from streamchat_b200.qwen2 import Qwen2Config, Qwen2Engine
    from streamchat_b200.tp_decode import TensorParallelDecoder
    from streamchat_b200.vision import ClipVisionConfig
    out = {"world": world}
    ok = True

    def llm_weights(cfg, seed, std):
        full = LlavaQwenConfig(llm=cfg, vision=ClipVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1,
                                                                num_attention_heads=2, image_size=56, patch_size=14))
        W = random_weights(full, device, seed=seed, std=std)
        return {k: v for k, v in W.items() if not k.startswith("model.vision_tower") and "mm_projector" not in k}

    # ---- 1. correctness on a small model
    cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4 * max(1, world // 2),
                      num_key_value_heads=world, vocab_size=1024 * world // 2 if world > 2 else 1008)
    eng = Qwen2Engine(llm_weights(cfg, 5, 0.05), cfg, device, max_ctx=1024, max_batch=2)
    tp = TensorParallelDecoder(eng, rank, world)
    g = torch.Generator(device=device).manual_seed(3)
    res = []
    for n_tok, n_new in ((37, 12), (150, 20)):
        x = (torch.randn(n_tok, cfg.hidden_size, device=device, generator=g) * 0.5).half()
        want = eng.generate(x, n_new)
        got = tp.generate(x, n_new)
        same = torch.equal(want, got)
        flag = torch.tensor([int(same)], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gathered = [torch.zeros_like(got) for _ in range(world)]
        dist.all_gather(gathered, got.contiguous())
        identical_across_ranks = all(torch.equal(gathered[0], t) for t in gathered)
        res.append({"prompt": n_tok, "new": n_new, "equals_single_gpu_on_all_ranks": bool(flag.item()),
                    "identical_across_ranks": identical_across_ranks,
                    "first_mismatch": next((i for i, (a, b) in enumerate(zip(want[0].tolist(), got[0].tolist())) if a != b), None)})
        ok = ok and bool(flag.item()) and identical_across_ranks
    out["small_model"] = res
    tp.close()
    del eng, tp
    torch.cuda.empty_cache()

    # ---- 2. LongVA-7B shapes: step time
    if "--full" in sys.argv:
        cfg = SYNTHETIC["longva7b"].llm
        eng = Qwen2Engine(llm_weights(cfg, 0, 0.02), cfg, device, max_ctx=32768, max_batch=1)
        tp = TensorParallelDecoder(eng, rank, world)
        # numeric check at full shapes: ONE decode step from the same prefilled state, TP logits vs the persistent kernel's
        x = (torch.randn(700, cfg.hidden_size, device=device) * 0.02).half()
        eng.generate(x, 2)                                   # persistent kernel: prefill + one step
        ref = eng._logits[0].float().clone()
        tp.generate(x, 2)                                    # same prefill, one TP step
        got = tp.px_l.buffer.view(-1).float()
        rel = float((got - ref).norm() / ref.norm())
        top2 = torch.topk(ref, 2).values
        out["one_step_logits"] = {"rel_l2_vs_persistent_kernel": rel, "argmax_equal": bool(int(got.argmax()) == int(ref.argmax())),
                                  "ref_top2_margin_rel": float((top2[0] - top2[1]) / ref.abs().max())}
        ok = ok and rel < 5e-3
        timing = {}
        for ctx in (5800, 18000):
            x = (torch.randn(ctx, cfg.hidden_size, device=device) * 0.02).half()
            steps = 64
            row = {}
            for name, fn in (("tp", lambda: tp.generate(x, steps + 1)), ("mega_one_gpu", lambda: eng.generate(x, steps + 1))):
                fn()
                eng.profile = []
                dist.barrier()
                ids = fn()
                torch.cuda.synchronize()
                p = [q for q in eng.profile if q.get("kind", "decode") == "decode"][0]
                ms = p["events"][0].elapsed_time(p["events"][1]) / p["steps"]
                t = torch.tensor([ms], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row[name + "_ms_per_step"] = float(t)
                row[name + "_ids_head"] = ids[0, :6].tolist()
                eng.profile = None
            row["speedup"] = row["mega_one_gpu_ms_per_step"] / row["tp_ms_per_step"]
            timing[str(ctx)] = row
        out["longva7b"] = timing
        tp.close()
    out["ok"] = ok
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
