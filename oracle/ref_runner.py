"""Run the REFERENCE's own implementation of the hot path on the host CPU (TEST / BASELINE INFRASTRUCTURE).

Loads the unmodified reference sources staged under oracle/_ref/ (oracle/stage_ref.py) — `longva` model classes on top
of the installed HF transformers, `utiles.weighted_kmeans_feature` (AST-extracted, as tests/golden/make_golden.py does),
`torch_kmeans` — with the two import shims of SURVEY.md §A.1, builds random-init models of the LongVA-7B widths with a
REDUCED layer count, and times the reference's public calls (`model.encode_images`, `model.generate_with_image_embedding`)
with per-layer forward hooks so that the bounded sample extrapolates to the full depth as

    t(full) = t(measured) + (L_full - L_sample) * mean per-layer time.

Used only by bench.py's `cpu_baseline` / `--impl reference` legs and tools/cpu_baselines.py (kind = "reference").
"""
from __future__ import annotations

import ast
import os
import sys
import tempfile
import time
import types
from typing import Dict, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

VIT_L14_336 = dict(hidden_size=1024, intermediate_size=4096, num_attention_heads=16, image_size=336, patch_size=14,
                   hidden_act="quick_gelu", layer_norm_eps=1e-5)
QWEN2_7B = dict(hidden_size=3584, intermediate_size=18944, num_attention_heads=28, num_key_value_heads=4,
                vocab_size=152064, max_position_embeddings=32768, rope_theta=1000000.0, rms_norm_eps=1e-6,
                tie_word_embeddings=False)
VIT_LAYERS_FULL, LLM_LAYERS_FULL = 24, 28      # the reference computes all 24 CLIP layers (it reads hidden_states[-2])


def available() -> bool:
    return os.path.exists(os.path.join(REF, "longva", "model", "language_model", "llava_qwen.py"))


_loaded = {}


def _import_reference():
    """`longva` from oracle/_ref with the shims of SURVEY §A.1 (qformer stub; ignore `config.rope_scaling = None`)."""
    if "cls" in _loaded:
        return _loaded["cls"]
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python -m oracle.stage_ref` where /root/reference exists")
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k == "longva" or k.startswith("longva.")]:
        del sys.modules[k]                      # a drop-in `longva` may be loaded: the reference's must win here
    sys.path.insert(0, REF)
    qf = types.ModuleType("longva.model.multimodal_resampler.qformer")
    qf.Qformer = object
    sys.modules["longva.model.multimodal_resampler.qformer"] = qf
    from longva.model.language_model.llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM
    from longva.utils import disable_torch_init
    _o = LlavaQwenConfig.__setattr__
    LlavaQwenConfig.__setattr__ = lambda s, k, v: None if (k == "rope_scaling" and v is None) else _o(s, k, v)
    disable_torch_init()                        # the driver calls it too (inference_streaming_longva_v2.py:688)
    _loaded["cls"] = (LlavaQwenConfig, LlavaQwenForCausalLM)
    return _loaded["cls"]


def build_reference_model(vit_layers: int, llm_layers: int, clip: Optional[dict] = None, llm: Optional[dict] = None):
    """The reference's LlavaQwenForCausalLM (fp32, eval, CPU), random init, LongVA-7B widths unless overridden."""
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel
    LlavaQwenConfig, LlavaQwenForCausalLM = _import_reference()
    clip = dict(clip or VIT_L14_336, num_hidden_layers=vit_layers)
    llm = dict(llm or QWEN2_7B, num_hidden_layers=llm_layers)
    tmp = tempfile.mkdtemp(prefix="scb_ref_tower_")
    CLIPVisionModel(CLIPVisionConfig(**clip)).save_pretrained(tmp)
    s = clip["image_size"]
    CLIPImageProcessor(size={"shortest_edge": s}, crop_size={"height": s, "width": s}).save_pretrained(tmp)
    cfg = LlavaQwenConfig(**llm, mm_vision_tower=tmp, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                          mm_projector_type="mlp2x_gelu", mm_hidden_size=clip["hidden_size"],
                          mm_resampler_type="spatial_pool", mm_spatial_pool_mode="average", mm_spatial_pool_stride=2,
                          mm_use_im_start_end=False, mm_patch_merge_type="flat")
    model = LlavaQwenForCausalLM(cfg)
    model.get_model().get_vision_tower().load_model()
    return model.float().eval()


class _LayerTimer:
    """Accumulates the wall time spent inside a list of modules (forward pre / post hooks; no change of code path)."""

    def __init__(self, modules):
        self.total, self.calls, self._t0 = 0.0, 0, 0.0
        self._h = []
        for m in modules:
            self._h.append(m.register_forward_pre_hook(self._pre))
            self._h.append(m.register_forward_hook(self._post))

    def _pre(self, *a):
        self._t0 = time.perf_counter()

    def _post(self, *a):
        self.total += time.perf_counter() - self._t0
        self.calls += 1

    def reset(self):
        self.total, self.calls = 0.0, 0

    def remove(self):
        for h in self._h:
            h.remove()


def time_encode_images(model, n_frames: int, seed: int = 0) -> Dict[str, float]:
    """`model.encode_images(images)` (llava_arch.py:179-184) on n_frames synthetic frames.
    Returns seconds: total, and the share spent inside the CLIP encoder layers (for the depth extrapolation)."""
    tower = model.get_model().get_vision_tower().vision_tower
    layers = list(tower.vision_model.encoder.layers)
    s = tower.config.image_size
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n_frames, 3, s, s, generator=g)
    tm = _LayerTimer(layers)
    with torch.no_grad():
        model.encode_images(images[:1])                         # warm-up (thread pool, allocator)
        tm.reset()
        t0 = time.perf_counter()
        out = model.encode_images(images)
        total = time.perf_counter() - t0
    tm.remove()
    return {"seconds": total, "layer_seconds": tm.total, "layers": len(layers), "frames": n_frames,
            "out_shape": tuple(out.shape)}


def time_generate(model, n_visual_tokens: int, n_text_tokens: int, max_new_tokens: int, seed: int = 0) -> Dict[str, float]:
    """`model.generate_with_image_embedding` (llava_qwen.py:137-155), greedy, no stop id: a prompt of n_text_tokens ids
    with one <image> placeholder spliced with n_visual_tokens feature rows, then max_new_tokens cached decode steps.
    Returns seconds split into prefill / decode and, for each, the share inside the decoder layers."""
    layers = list(model.model.layers)
    g = torch.Generator().manual_seed(seed)
    H = model.config.hidden_size
    ids = torch.randint(0, model.config.vocab_size, (1, n_text_tokens), generator=g)
    ids[0, n_text_tokens // 2] = -200
    emb = torch.randn(n_visual_tokens, H, generator=g)
    tm = _LayerTimer(layers)
    marks = []

    # the first call of layer 0 with more than one token is the prefill; later calls are decode steps
    def mark(mod, args, kwargs):
        marks.append((time.perf_counter(), tm.total))
    h = layers[0].register_forward_pre_hook(mark, with_kwargs=True)
    with torch.no_grad():
        t0 = time.perf_counter()
        out = model.generate_with_image_embedding(ids, image_embeddings=[emb], modalities=["video"], do_sample=False,
                                                  temperature=0, top_p=None, num_beams=1, max_new_tokens=max_new_tokens,
                                                  use_cache=False, eos_token_id=None, pad_token_id=0)
        t1 = time.perf_counter()
    h.remove()
    tm.remove()
    assert out.shape[1] == max_new_tokens and len(marks) == max_new_tokens, (out.shape, len(marks))
    # marks[0] = start of prefill layers, marks[1] = start of the first decode step's layers
    t_dec0, layer_at_dec0 = marks[1] if len(marks) > 1 else (t1, tm.total)
    prefill = t_dec0 - t0
    decode = t1 - t_dec0
    return {"prefill_seconds": prefill, "prefill_layer_seconds": layer_at_dec0, "prefill_tokens": n_visual_tokens + n_text_tokens - 1,
            "decode_seconds": decode, "decode_layer_seconds": tm.total - layer_at_dec0, "decode_steps": max_new_tokens - 1,
            "layers": len(layers)}


def reference_function(name: str):
    """A function of the reference's utiles.py, AST-extracted (importing the module needs matplotlib / gradio / langchain)."""
    if ("fn", name) in _loaded:
        return _loaded[("fn", name)]
    import math
    import random
    from collections import defaultdict

    import numpy as np
    import torch.nn.functional as F
    tree = ast.parse(open(os.path.join(REF, "utiles.py")).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name]
    if not body:
        raise KeyError(name)
    ns = {"torch": torch, "F": F, "np": np, "math": math, "random": random, "time": time, "defaultdict": defaultdict}
    exec(compile(ast.Module(body=body, type_ignores=[]), "utiles.py", "exec"), ns)
    _loaded[("fn", name)] = ns[name]
    return ns[name]


def time_weighted_kmeans(n_frames: int, tokens: int, dim: int, k: int = 5, seed: int = 0) -> Dict[str, float]:
    """utiles.weighted_kmeans_feature (utiles.py:291-330) on [n_frames, tokens, dim] fp32."""
    fn = reference_function("weighted_kmeans_feature")
    torch.manual_seed(seed)
    x = torch.randn(n_frames, tokens, dim)
    t0 = time.perf_counter()
    fn(x, k)
    return {"seconds": time.perf_counter() - t0, "frames": n_frames, "bytes_per_pass": x.numel() * 4}


def time_torch_kmeans(seed: int = 0) -> Dict[str, float]:
    """Config #1: torch_kmeans.KMeans(n_clusters=5) on [1, 64, 1024] fp32 (defaults: 8 restarts, <= 100 iterations)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from torch_kmeans import KMeans
    torch.manual_seed(seed)
    x = torch.randn(1, 64, 1024)
    km = KMeans(n_clusters=5, verbose=False)
    km(x)
    t0 = time.perf_counter()
    km(x)
    return {"seconds": time.perf_counter() - t0}
