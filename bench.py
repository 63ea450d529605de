#!/usr/bin/env python
"""bench.py — streamed frames/sec of the StreamChat hot path (ViT -> k-means -> retrieve -> LongVA
decode) on N B200s, plus the reference's CPU path timed beside it.

    python bench.py --gpus 1 --steps 3 --warmup 3                 # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 # the reference's own code on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over ONE synthetic stream (BASELINE.json configs[4]): `--frames` (1000)
336x336 frames arriving in segments of `--segment` (200); per segment [A] CLIP-ViT encode, [B] memory update
(forgetting-curve short memory, 40-frame chunks each captioned with a 128-token greedy LongVA generation, tree merge
= k-means over 400 frames + 256-token text summary); then [C] cosine top-k tree search and a 256-token answer.
N > 1 shards THAT stream (strong scaling, north_star): chunk c -> rank c mod N encodes and captions it, ONE
all-gather of the chunk features (node centroids, 1.03 GB) + caption ids over NVLink, memory updates replayed on the
gathered memory, answer on rank 0 while the other ranks generate the merge summaries; `--scaling weak` runs the
round-1 formulation (one independent 1000-frame stream per rank) and is also reported as the secondary key `weak`.
Random-init LongVA-7B-shaped weights, synthetic frames (no datasets/checkpoints exist here).  Prints ONE JSON line
(rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "streamed frames/sec (ViT->kmeans->retrieve->LongVA decode)"
QUESTION = "What objects did I see on the table before I left the room, and where did I put the keys?"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="synthetic:longva7b")
    ap.add_argument("--frames", type=int, default=None, help="frames per rank per step (default 1000; tiny: 160)")
    ap.add_argument("--segment", type=int, default=200, help="frames per ingest/update segment")
    ap.add_argument("--caption-tokens", type=int, default=128)
    ap.add_argument("--answer-tokens", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: shard ONE stream over the ranks (strong, default) or one stream per rank (weak)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1, strong: skip the secondary weak-scaling measurement")
    ap.add_argument("--no-peer", action="store_true", help="N > 1: move the gather with NCCL instead of scb_allgather_embed")
    ap.add_argument("--text-encoder", default="bert", choices=["bert", "hash"],
                    help="retrieval text embedder: BERT-large-shaped encoder on the GPU, or synthetic hash vectors")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense 16-bit tensor throughput a kernel timed INSIDE a long step is held against: the sustained figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if d.get("bf16_tflops_sustained"):
            return float(d["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
    return 1380.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
STAGE_EVENTS = []   # (stage name, start event, end event) of the last timed region


def _stage(name):
    class _S:
        def __enter__(self_):
            self_.e0 = torch.cuda.Event(enable_timing=True)
            self_.e1 = torch.cuda.Event(enable_timing=True)
            self_.e0.record()

        def __exit__(self_, *a):
            self_.e1.record()
            STAGE_EVENTS.append((name, self_.e0, self_.e1))
    return _S()


def workload_config(args, n_frames: int, tiny: bool, world: int) -> dict:
    """The `config` object of the JSON line — shared by both arms so the driver compares like with like."""
    strong = world == 1 or args.scaling == "strong"
    if world == 1:
        sharding = "single rank"
    elif strong:
        sharding = (f"ONE stream over {world} ranks: chunk c (40 frames) -> rank c mod {world} (encode + caption), one in-place "
                    "all-gather of chunk features + caption ids over NVLink, memory updates replayed on every rank, answer on "
                    "rank 0 while ranks 1.. generate the merge summaries")
    else:
        sharding = f"{world} independent streams, one per rank (weak scaling)"
    total = n_frames if strong else n_frames * world
    return {"workload": (f"stream: {total} frames 336x336 in segments of {args.segment}, chunk 40, "
                         f"{args.caption_tokens}-token greedy caption per chunk, tree merge every 10 chunks "
                         f"(k-means 400->5 + 256-token summary), 1 question ({args.answer_tokens} tokens)")
                        if not tiny else f"tiny smoke stream: {total} frames",
            "model": args.model, "weights": "random-init normal(0,0.02) fp16", "frames_total": total,
            "frames_per_rank": n_frames if not strong else None,
            "text_encoder": ("BERT-large shape on the GPU (random init), node embeddings cached"
                             if (args.text_encoder == "bert" and not tiny) else "synthetic hash vectors"),
            "sharding": sharding,
            "l2": "inputs larger than L2 (677 MB frames + 15 GB weights per step); no explicit flush"}


def run_stream(session_factory, frames_dev, host_frames, segment, device, e2e: bool):
    """One step on one rank: the whole stream through a StreamingSession.  Returns (answer ids, h2d bytes, d2h bytes)."""
    from streamchat_b200.pipeline import FramePrefetcher
    sess = session_factory()
    n = frames_dev.shape[0] if frames_dev is not None else host_frames.shape[0]
    h2d = d2h = 0
    slices = [(s0, min(s0 + segment, n)) for s0 in range(0, n, segment)]
    # e2e: pinned host frames -> HBM inside the timed region, segment i+1 staged on a copy stream while i is encoded
    pre = FramePrefetcher(host_frames, slices, device) if e2e else None
    for i, seg in (pre if e2e else ((i, frames_dev[a:b]) for i, (a, b) in enumerate(slices))):
        with _stage("encode"):
            feats = sess.ingest(seg)
        with _stage("memory_update"):
            sess.update_memory(feats, kmeans_init_indices=list(range(sess.cfg.num_clusters)))
    if pre is not None:
        h2d = pre.bytes_copied
    with _stage("answer"):
        text, ids = sess.answer(QUESTION)
    if e2e:
        ids_host = ids.cpu()                                                  # device -> host read of the result
        d2h += ids_host.numel() * ids_host.element_size()
    return ids, h2d, d2h


def run_stream_sharded(session_factory, gather_buffer, my_frames_dev, my_frames_host, seg_slices, n_frames, segment,
                       world, rank, device, e2e: bool):
    """One step of the frame-sharded stream (distributed.ShardedStreamingSession): this rank ingests only the frames of
    the chunks it owns (`seg_slices[i]` = its rows for segment i)."""
    from streamchat_b200.distributed import ShardedStreamingSession
    sess = session_factory()
    sh = ShardedStreamingSession(sess, rank, world, n_frames, kmeans_init_indices=list(range(sess.cfg.num_clusters)),
                                 gather_buffer=gather_buffer)
    from streamchat_b200.pipeline import FramePrefetcher
    h2d = d2h = 0
    pre = FramePrefetcher(my_frames_host, seg_slices, device) if e2e else None
    it = pre if e2e else ((i, (my_frames_dev[a:b] if b > a else None)) for i, (a, b) in enumerate(seg_slices))
    for i, seg in it:
        first = i * segment
        with _stage("encode"):
            sh.ingest_segment(first, min(segment, n_frames - first), seg)
    if pre is not None:
        h2d = pre.bytes_copied
    with _stage("captions"):
        sh.caption_my_chunks()
    with _stage("exchange"):
        sh.allgather_embed()
    with _stage("memory_update"):
        sh.replay_updates()
    with _stage("answer"):
        text, ids = sh.answer(QUESTION)
    if e2e and rank == 0:
        ids_host = ids.cpu()
        d2h += ids_host.numel() * ids_host.element_size()
    return ids, h2d, d2h, sh.exchanged_bytes


def ours(args):
    import torch.distributed as dist
    from streamchat_b200 import _lib
    from streamchat_b200.distributed import make_gather_buffer, owned_chunks
    from streamchat_b200.longva.model.builder import load_pretrained_model
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to STDOUT when the communicator is created (NCCL_DEBUG=VERSION/INFO in the
        # environment): create it here, with file descriptor 1 pointed at stderr, so that stdout carries the JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=device)
            dist.all_reduce(torch.zeros(1, device=device))
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    torch.cuda.set_device(device)
    _lib.load()

    tiny = args.model.endswith("tiny")
    n_frames = args.frames or (160 if tiny else 1000)
    strong = world > 1 and args.scaling == "strong"
    tokenizer, model, _, _ = load_pretrained_model(args.model, None, "llava_qwen", device_map=str(device))
    S = model.config.vision.image_size
    if args.text_encoder == "bert" and not tiny:
        # retrieval text encoder on the GPU (BERT-large shape, random init); texts are tokenised
        # by the word-hash stand-in tokenizer over the BERT vocabulary
        from streamchat_b200.longva.model.builder import ByteTokenizer
        from streamchat_b200.text_encoder import BertConfig, BertTextEncoder, random_bert_weights
        bcfg = BertConfig()
        embedder = BertTextEncoder(random_bert_weights(bcfg, device, seed=7), bcfg, device)
        embed_tok = ByteTokenizer(bcfg.vocab_size)
    else:
        embedder, embed_tok = HashTextEmbedder(1024, device), None
    cfg = StreamConfig(answer_tokens=args.answer_tokens)
    import streamchat_b200.memory as M
    M.CAPTION_TOKENS = args.caption_tokens

    def factory():
        np.random.seed(1234)                 # forgetting-curve sampling uses the global numpy RNG (H1)
        return StreamingSession(model, tokenizer, embedder, embed_tok, None, cfg)

    # the synthetic stream: strong scaling shards ONE stream (same seed on every rank, each rank keeps the frames of
    # its chunks); weak scaling / one GPU gives every rank its own stream
    chunk = cfg.chunk_size
    g = torch.Generator(device=device).manual_seed(1000 + (0 if strong else rank))
    frames_dev = torch.randn((n_frames, 3, S, S), generator=g, device=device, dtype=torch.float16)
    seg_slices, gather_buffer = [], None
    if strong:
        keep, off = [], 0
        for first in range(0, n_frames, args.segment):
            nseg = min(args.segment, n_frames - first)
            mine = owned_chunks(first // chunk, math.ceil(nseg / chunk), rank, world)
            idx = [f for c in mine for f in range(c * chunk, min((c + 1) * chunk, first + nseg))]
            keep += idx
            seg_slices.append((off, off + len(idx)))
            off += len(idx)
        frames_dev = frames_dev[torch.tensor(keep, dtype=torch.long, device=device)].contiguous() if keep \
            else frames_dev[:0].contiguous()
        gather_buffer = make_gather_buffer(model, chunk, n_frames, rank, world, peer=not args.no_peer)
    host_frames = torch.empty(frames_dev.shape, dtype=torch.float16, pin_memory=True)
    host_frames.copy_(frames_dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(e2e: bool, sharded: bool):
        if sharded:
            return run_stream_sharded(factory, gather_buffer, frames_dev, host_frames, seg_slices, n_frames, args.segment,
                                      world, rank, device, e2e)
        ids, a, b = run_stream(factory, None if e2e else frames_dev, host_frames, args.segment, device, e2e)
        return ids, a, b, 0

    def timed(e2e: bool, steps: int, sharded: bool):
        model.llm.profile = []
        STAGE_EVENTS.clear()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        r0 = model.llm.graph_launches
        e0.record()
        h2d = d2h = xb = 0
        for _ in range(steps):
            _, h2d, d2h, xb = one_step(e2e, sharded)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, float(h2d), float(d2h)], device=device, dtype=torch.float64)
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, h2d, d2h = float(mx[0]), int(t[1]), int(t[2])          # time: max over ranks; bytes: all ranks
        launches = (_lib.launch_count() - l0) + (model.llm.graph_launches - r0)
        return ms, h2d, d2h, launches, list(model.llm.profile), xb

    for _ in range(max(args.warmup, 0)):
        one_step(False, strong)
    sampler = ClockSampler(local) if rank == 0 else None     # one nvidia-smi poller per job (rank 0's GPU), not one per rank
    if sampler:
        sampler.start()
    ms, _, _, launches, prof, xbytes = timed(False, args.steps, strong)
    clocks = sampler.stop() if sampler else None
    breakdown = {}
    for name, a0, a1 in STAGE_EVENTS:
        breakdown[name] = breakdown.get(name, 0.0) + a0.elapsed_time(a1) / args.steps
    ms_e2e, h2d, d2h, _, _, _ = timed(True, args.steps, strong)

    streams = 1 if (strong or world == 1) else world
    total_frames = n_frames * streams * args.steps
    value = total_frames / (ms * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)

    weak = None
    if strong and not args.no_weak:
        # secondary: the round-1 formulation (one independent stream per rank, no data-path collective)
        g2 = torch.Generator(device=device).manual_seed(1000 + rank)
        frames_dev = None
        frames_w = torch.randn((n_frames, 3, S, S), generator=g2, device=device, dtype=torch.float16)
        run_stream(factory, frames_w, None, args.segment, device, False)
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        run_stream(factory, frames_w, None, args.segment, device, False)
        w1.record()
        barrier()
        t = torch.tensor([w0.elapsed_time(w1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        weak = {"value": n_frames * world / (float(t) * 1e-3), "unit": "frames/s", "scaling": "weak",
                "frames_per_rank": n_frames, "steps": 1, "ms_per_step": float(t),
                "note": "one independent stream per rank (the round-1 bench formulation), device-resident frames"}

    # roofline of the dominant kernel family: the weight-streaming decode step,
    # timed live (CUDA events on the launching stream around every decode loop of the timed region, rank 0)
    peak, peak_src = load_peaks()
    for p in prof:
        p["ms"] = p["events"][0].elapsed_time(p["events"][1])
    dec = [p for p in prof if p.get("kind", "decode") == "decode"]
    pre = [p for p in prof if p.get("kind") == "prefill"]
    dec_ms = sum(p["ms"] for p in dec)
    dec_steps = sum(p["steps"] for p in dec)
    dec_bytes = sum(p["bytes"] for p in dec)
    pre_ms = sum(p["ms"] for p in pre)
    if pre_ms > 0:   # tensor-bound half of the generation stage, reported beside the HBM-bound decode roofline
        breakdown["llm_prefill"] = pre_ms / args.steps
        breakdown["llm_prefill_tokens"] = sum(p["tokens"] for p in pre) / args.steps
        breakdown["llm_prefill_tflops"] = sum(p["flops"] for p in pre) / (pre_ms * 1e-3) / 1e12
    if dec_ms > 0:
        breakdown["llm_decode"] = dec_ms / args.steps
        breakdown["llm_decode_rows"] = {str(r): sum(p["steps"] for p in dec if p["rows"] == r) // args.steps
                                        for r in sorted({p["rows"] for p in dec})}
    if world > 1:
        breakdown["note"] = "stage times and decode profile are rank 0's; ms_per_step is the max over ranks"
        if strong:
            breakdown["exchange_bytes_received_per_rank"] = xbytes
            breakdown["gather"] = "scb_allgather_embed (NVLink peer memory)" if not isinstance(gather_buffer, torch.Tensor) \
                else "NCCL all_gather_into_tensor (in place)"
    roofline = None
    traffic = None
    for tname in ("r02_decode_mega_traffic.json", "r01_decode_mega_traffic.json"):
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", tname)
        if os.path.exists(tpath):  # DRAM bytes of ONE decode_mega_kernel launch (rows / ctx as recorded) from the ncu capture named there
            with open(tpath) as f:
                tj = json.load(f)
            traffic = {"bytes_per_decode_step": tj["dram_read_bytes"] + tj["dram_write_bytes"],
                       "algorithmic_bytes_same_step": tj["algorithmic_bytes"], "rows": tj["rows"], "ctx": tj["ctx"],
                       "source": "profiles/" + tname}
            break
    if dec_steps:
        achieved = dec_bytes / (dec_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "decode step: decode_mega_kernel (ONE persistent launch per step, <= 4 rows: TMA weight ring + "
                              "mma.sync consumers + grid barriers); linear_tcgen05 split-K + reduce + decode_attn_fused for wider "
                              "batches; bytes = weights + live KV per step",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src,
                    # DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed ncu capture
                    "traffic": traffic["bytes_per_decode_step"] if traffic else None, "traffic_detail": traffic,
                    "ms_per_decode_step": dec_ms / dec_steps, "decode_steps": dec_steps,
                    "share_of_step_time": dec_ms / (ms if ms else 1)}

    roofline_tensor = None
    if pre_ms > 0:
        tpeak, tsrc = load_tensor_peak()
        ach = breakdown["llm_prefill_tflops"]
        roofline_tensor = {"bound": "tensor", "kernel": "LLM prefill stage: linear_tcgen05_kernel<256, pair> (84 % of the stage in the ncu "
                           "launch list) + attention_tcgen05_kernel (13 %) + norms / rope; FLOPs = dense 2*6.525e9*T + causal attention",
                           "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "peak_source": tsrc,
                           "traffic": None, "share_of_step_time": pre_ms / (ms if ms else 1)}
    roofline_encode = None
    if breakdown.get("encode"):
        # the other tensor-bound stage: CLIP tower + projector + pool over this rank's frames (rank 0's stage time)
        tpeak, tsrc = load_tensor_peak()
        enc_frames = int(host_frames.shape[0])          # this rank's frames of the stream (frames_dev may be released by now)
        enc_flops = model.vision.flops_per_frame() * enc_frames
        ach = enc_flops / (breakdown["encode"] * 1e-3) / 1e12
        breakdown["encode_tflops"] = ach
        roofline_encode = {"bound": "tensor", "kernel": "encode stage: linear_tcgen05_kernel<256, pair> (75 % of a pass in the ncu launch "
                           "list) + attention_tcgen05_kernel<64> (17 %) + LayerNorm / im2col / pool; FLOPs = "
                           f"{model.vision.flops_per_frame() / 1e9:.2f} GF per frame (the tower layers feature_select needs, projector "
                           "Linear-2 after the pool)",
                           "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "peak_source": tsrc,
                           "traffic": None, "frames": enc_frames, "share_of_step_time": breakdown["encode"] * args.steps / (ms if ms else 1)}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if (strong or world == 1) else "weak",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": workload_config(args, n_frames, tiny, world),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "impl": "ours",
        "breakdown_ms_per_step": breakdown,
    }
    if roofline_tensor is not None:
        line["roofline_tensor"] = roofline_tensor      # the tensor-bound half of the step, beside the HBM-bound `roofline`
    if roofline_encode is not None:
        line["roofline_encode"] = roofline_encode
    if weak is not None:
        line["weak"] = weak
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n_frames, args, tiny)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
_REF_MODEL = {}


def cpu_baseline(n_frames: int, args, tiny: bool):
    """The reference's CPU path on a bounded sample of the same workload.  With the reference staged under oracle/_ref
    (oracle/stage_ref.py; done by build() wherever /root/reference exists) this is the REFERENCE'S OWN code —
    LlavaQwenForCausalLM.encode_images / generate_with_image_embedding over HF transformers and
    utiles.weighted_kmeans_feature — at the LongVA-7B widths with a reduced layer count, extrapolated to the full depth
    with per-layer timings (kind "reference"); otherwise the oracle port (kind "port")."""
    import contextlib
    from oracle import ref_runner as R
    if tiny or not R.available():
        return cpu_baseline_port(n_frames, args, tiny)
    with contextlib.redirect_stdout(sys.stderr):      # the reference prints while importing / loading: stdout carries ONE JSON line
        return _cpu_baseline_reference(n_frames, args, R)


def _cpu_baseline_reference(n_frames: int, args, R):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Lv, Ll = 2, 2
    if "m" not in _REF_MODEL:
        t0 = time.perf_counter()
        _REF_MODEL["m"] = R.build_reference_model(Lv, Ll)
        _REF_MODEL["build_s"] = time.perf_counter() - t0
    model = _REF_MODEL["m"]
    chunk_tokens, n_text = 40 * 144, 56
    enc = R.time_encode_images(model, 4)
    gen = R.time_generate(model, chunk_tokens, n_text, 8)
    km = R.time_weighted_kmeans(100, 144, 3584)
    # ---- extrapolation to the full depth: t = (time outside the layers) + L_full x (time per layer)
    t_frame = (enc["seconds"] - enc["layer_seconds"]) / enc["frames"] + R.VIT_LAYERS_FULL * enc["layer_seconds"] / (Lv * enc["frames"])
    pre_over, pre_layer = gen["prefill_seconds"] - gen["prefill_layer_seconds"], gen["prefill_layer_seconds"] / Ll
    dec_over = (gen["decode_seconds"] - gen["decode_layer_seconds"]) / gen["decode_steps"]
    dec_layer = gen["decode_layer_seconds"] / (Ll * gen["decode_steps"])
    t_prefill_chunk = pre_over + R.LLM_LAYERS_FULL * pre_layer                 # 5 816-token caption prompt
    t_token = dec_over + R.LLM_LAYERS_FULL * dec_layer
    n_chunks = math.ceil(n_frames / 40)
    n_merges = n_chunks // 10
    t_caption = t_prefill_chunk + (args.caption_tokens - 1) * t_token
    answer_tokens_in = 5 * 144 + (n_merges + min(1, n_chunks % 10)) * chunk_tokens + 120   # short memory + retrieved nodes
    t_answer = t_prefill_chunk * answer_tokens_in / gen["prefill_tokens"] + (args.answer_tokens - 1) * t_token
    t_summary = t_prefill_chunk * 1400 / gen["prefill_tokens"] + 255 * t_token          # ~1 400-token text prompt
    t_kmeans = n_merges * km["seconds"] * 400 / km["frames"]
    total = n_frames * t_frame + n_chunks * t_caption + n_merges * t_summary + t_answer + t_kmeans
    return {"value": n_frames / total, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": (f"the reference's own LlavaQwenForCausalLM (oracle/_ref, HF transformers) fp32 on {cores} threads at the "
                       f"LongVA-7B widths with {Lv} of 24 CLIP and {Ll} of 28 Qwen2 layers: encode_images on {enc['frames']} frames "
                       f"({enc['seconds']:.2f}s, {enc['layer_seconds']:.2f}s in layers), generate_with_image_embedding with one full "
                       f"{gen['prefill_tokens']}-token caption prompt ({gen['prefill_seconds']:.2f}s prefill, {gen['prefill_layer_seconds']:.2f}s in "
                       f"layers) + {gen['decode_steps']} cached decode steps ({gen['decode_seconds']:.2f}s, {gen['decode_layer_seconds']:.2f}s in layers), "
                       f"utiles.weighted_kmeans_feature on 100x144x3584 ({km['seconds']:.2f}s); extrapolated to full depth as "
                       f"outside-layer time + L_full x per-layer time, then to {n_frames} frames, {n_chunks} captions x "
                       f"({gen['prefill_tokens']} prefill + {args.caption_tokens} decode tokens), {n_merges} merges (k-means over 400 "
                       f"frames + 256-token summary) and one {args.answer_tokens}-token answer over ~{answer_tokens_in} prompt tokens "
                       f"(prefill scaled linearly in tokens: attention's quadratic term ignored, favours the CPU)"),
            "raw": {"encode": enc, "generate": gen, "kmeans": km, "model_build_seconds": _REF_MODEL["build_s"]},
            "per_unit_seconds": {"frame_encode": t_frame, "caption": t_caption, "decode_token": t_token, "summary": t_summary,
                                 "answer": t_answer, "kmeans_merges": t_kmeans},
            "extrapolated_seconds_per_stream": total}


def cpu_baseline_port(n_frames: int, args, tiny: bool):
    """Fallback when oracle/_ref is not staged: the oracle port (oracle/longva_model.py) timed on the host cores on a
    bounded sample of the same workload, extrapolated linearly in frames / layers / tokens."""
    from oracle import longva_model as om
    from streamchat_b200.longva.model.builder import SYNTHETIC
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = SYNTHETIC["tiny" if tiny else "longva7b"]
    v, l = cfg.vision, cfg.llm
    g = torch.Generator().manual_seed(0)

    def rnd(*shape):
        return torch.randn(shape, generator=g) * 0.02

    # ---- ViT: one frame through 2 of the needed layers at true width, extrapolated to all layers
    D, I = v.hidden_size, v.intermediate_size
    VT = om.VT
    W = {VT + "embeddings.class_embedding": rnd(D), VT + "embeddings.patch_embedding.weight": rnd(D, 3, v.patch_size, v.patch_size),
         VT + "embeddings.position_embedding.weight": rnd(v.tokens, D),
         VT + "pre_layrnorm.weight": torch.ones(D), VT + "pre_layrnorm.bias": torch.zeros(D)}
    vit_layers = 2
    for i in range(vit_layers):
        p = f"{VT}encoder.layers.{i}."
        for nme in ("q_proj", "k_proj", "v_proj", "out_proj"):
            W[p + f"self_attn.{nme}.weight"], W[p + f"self_attn.{nme}.bias"] = rnd(D, D), rnd(D)
        for nme in ("layer_norm1", "layer_norm2"):
            W[p + nme + ".weight"], W[p + nme + ".bias"] = torch.ones(D), torch.zeros(D)
        W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"] = rnd(I, D), rnd(I)
        W[p + "mlp.fc2.weight"], W[p + "mlp.fc2.bias"] = rnd(D, I), rnd(D)
    H = l.hidden_size
    W["model.mm_projector.0.weight"], W["model.mm_projector.0.bias"] = rnd(H, D), rnd(H)
    W["model.mm_projector.2.weight"], W["model.mm_projector.2.bias"] = rnd(H, H), rnd(H)
    px = torch.randn(1, 3, v.image_size, v.image_size, generator=g)
    with torch.no_grad():
        om.clip_tower(W, px, v.num_attention_heads, select_layer=vit_layers)          # warm-up
        t0 = time.perf_counter(); om.clip_tower(W, px, v.num_attention_heads, select_layer=vit_layers)
        t_vit2 = time.perf_counter() - t0
        feats = torch.randn(1, v.tokens - 1, D, generator=g)
        t0 = time.perf_counter()
        y = torch.nn.functional.gelu(om._lin(feats, W, "model.mm_projector.0")); om._lin(y, W, "model.mm_projector.2")
        t_proj = time.perf_counter() - t0
    t_frame = t_vit2 / vit_layers * v.layers_needed + t_proj

    # ---- Qwen2: 1 layer at true width; prefill sample of 256 tokens, 2 cached decode tokens, lm_head once
    LW = {}
    hd = l.head_dim
    p = "model.layers.0."
    LW[p + "self_attn.q_proj.weight"], LW[p + "self_attn.q_proj.bias"] = rnd(l.num_attention_heads * hd, H), rnd(l.num_attention_heads * hd)
    LW[p + "self_attn.k_proj.weight"], LW[p + "self_attn.k_proj.bias"] = rnd(l.num_key_value_heads * hd, H), rnd(l.num_key_value_heads * hd)
    LW[p + "self_attn.v_proj.weight"], LW[p + "self_attn.v_proj.bias"] = rnd(l.num_key_value_heads * hd, H), rnd(l.num_key_value_heads * hd)
    LW[p + "self_attn.o_proj.weight"] = rnd(H, H)
    LW[p + "mlp.gate_proj.weight"], LW[p + "mlp.up_proj.weight"] = rnd(l.intermediate_size, H), rnd(l.intermediate_size, H)
    LW[p + "mlp.down_proj.weight"] = rnd(H, l.intermediate_size)
    LW[p + "input_layernorm.weight"] = LW[p + "post_attention_layernorm.weight"] = LW["model.norm.weight"] = torch.ones(H)
    lm = rnd(l.vocab_size, H)
    Tp = 64 if tiny else 256
    x = torch.randn(Tp, H, generator=g)
    with torch.no_grad():
        cache = []
        t0 = time.perf_counter(); om.qwen2_forward(LW, x, l.num_attention_heads, l.num_key_value_heads, l.rope_theta, l.rms_norm_eps, cache, 0)
        t_prefill_layer_per_tok = (time.perf_counter() - t0) / Tp
        t0 = time.perf_counter()
        for s in range(2):
            om.qwen2_forward(LW, x[:1], l.num_attention_heads, l.num_key_value_heads, l.rope_theta, l.rms_norm_eps, cache, Tp + s)
        t_decode_layer = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter(); (x[:1] @ lm.t()).argmax(); t_lm = time.perf_counter() - t0
    L = l.num_hidden_layers
    chunk_tokens = 40 * ((v.grid // 2) ** 2) + 60
    n_chunks = math.ceil(n_frames / 40)
    t_caption = chunk_tokens * t_prefill_layer_per_tok * L + args.caption_tokens * (t_decode_layer * L + t_lm)
    t_answer = 3 * chunk_tokens * t_prefill_layer_per_tok * L + args.answer_tokens * (t_decode_layer * L + t_lm)
    total = n_frames * t_frame + n_chunks * t_caption + t_answer
    return {"value": n_frames / total, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": (f"oracle/longva_model.py fp32 on {cores} threads: 1 frame x {vit_layers} ViT layers + projector "
                       f"({t_vit2:.2f}s+{t_proj:.2f}s), 1 Qwen2 layer x {Tp}-token prefill ({t_prefill_layer_per_tok * Tp:.2f}s), "
                       f"2 cached decode tokens ({t_decode_layer:.3f}s/tok/layer), lm_head ({t_lm:.3f}s); extrapolated "
                       f"linearly to {v.layers_needed} ViT layers, {L} Qwen2 layers, {n_chunks} captions x ({chunk_tokens} prefill + "
                       f"{args.caption_tokens} decode tokens), 1 answer; attention's quadratic term ignored (favours the CPU)"),
            "extrapolated_seconds_per_stream": total}


def reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (the staged
    reference under oracle/_ref when present, else the oracle port), each step one bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    tiny = args.model.endswith("tiny")
    n_frames = args.frames or (160 if tiny else 1000)
    vals = []
    for _ in range(max(args.warmup, 0) + max(args.steps, 1)):
        vals.append(cpu_baseline(n_frames, args, tiny))
    timed = vals[max(args.warmup, 0):]
    best = dict(timed[-1])
    v = float(np.median([x["value"] for x in timed]))
    best["value"] = v
    best["values_per_step"] = [x["value"] for x in timed]
    print(json.dumps({
        "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_frames / v, "higher_is_better": True,
        "scaling": "strong" if args.scaling == "strong" or world == 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, n_frames, tiny, world),
        "cpu_baseline": best,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        reference(a)
    else:
        ours(a)
