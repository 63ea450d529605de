"""CPU oracle for the model part of the hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Plain-torch fp32 restatement, operator by operator, of what the reference executes through HF
transformers (arithmetic is third-party: transformers pinned 4.37.2 in requirements.txt:149,
5.5.0 installed here; operator order from SURVEY.md Appendix B):

  encode_images            longva/model/llava_arch.py:179-184  (tower -> projector -> 2x2 pool)
    CLIP vision tower      models/clip/modeling_clip.py:148-217,282-386,659-686;
                           feature_select hidden_states[-2][:, 1:] (clip_encoder.py:46-66)
    mm_projector           longva/model/multimodal_projector/builder.py:41-48 (Linear, GELU(erf), Linear)
    SpatialPool            longva/model/multimodal_resampler/spatial_pool.py:23-32
  prepare_inputs_embeds    longva/model/llava_arch.py:208-343 (batch 1, no padding)
  qwen2 forward / greedy   models/qwen2/modeling_qwen2.py:35-487; generate with inputs_embeds
                           returns only the new ids (llava_qwen.py:137-155)

Pinned against the reference itself: tests/golden/tiny_longva_ref.npz was produced by
longva's LlavaQwenForCausalLM (tests/golden/make_golden_models.py); tests/test_oracle_models.py
checks this restatement against it.  Weights are a dict keyed by HF state-dict names.

Two precisions: the default computes everything in fp32 (the mathematical function); `act_dtype=torch.float16`
additionally rounds every operator OUTPUT to fp16 at the points where the reference's own fp16 execution
(`model.half()` on the GPU: builder.py:213, inference_streaming_longva_v2.py:523) stores an fp16 tensor —
accumulation inside an operator stays fp32 as in cuBLAS / SDPA.  The second mode is what "within 1e-3 rel of
the reference" is measured against; the first bounds the distance to exact arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200
VT = "model.vision_tower.vision_tower.vision_model."


def _r(t: torch.Tensor, act_dtype) -> torch.Tensor:
    """Round to the activation storage type of the emulated execution (no-op for the fp32 oracle)."""
    return t if act_dtype is None else t.to(act_dtype).float()


def _lin(x, W, prefix, act_dtype=None):
    b = W.get(prefix + ".bias")
    return _r(F.linear(x, W[prefix + ".weight"].float(), None if b is None else b.float()), act_dtype)


def clip_tower(W: Dict[str, torch.Tensor], pixels: torch.Tensor, n_heads: int, eps: float = 1e-5,
               select_layer: int = -2, act_dtype=None) -> torch.Tensor:
    """[N,3,H,W] -> hidden_states[select_layer][:, 1:]  ([N, patches, hidden])."""
    r = lambda t: _r(t, act_dtype)
    x = pixels.float()
    pw = W[VT + "embeddings.patch_embedding.weight"].float()
    P = pw.shape[-1]
    pe = r(F.conv2d(x, pw, stride=P)).flatten(2).transpose(1, 2)              # :148-154,209-210
    cls = W[VT + "embeddings.class_embedding"].float().expand(x.shape[0], 1, -1)
    h = r(torch.cat([cls, pe], 1) + W[VT + "embeddings.position_embedding.weight"].float())  # :212-217
    D = h.shape[-1]
    h = r(F.layer_norm(h, (D,), W[VT + "pre_layrnorm.weight"].float(), W[VT + "pre_layrnorm.bias"].float(), eps))
    n_layers = 1 + max(int(k.split("encoder.layers.")[1].split(".")[0]) for k in W if "encoder.layers." in k and k.startswith(VT))
    hidden_states = [h]
    hd = D // n_heads
    for i in range(n_layers):
        p = f"{VT}encoder.layers.{i}."
        res = h
        y = r(F.layer_norm(h, (D,), W[p + "layer_norm1.weight"].float(), W[p + "layer_norm1.bias"].float(), eps))
        B, T, _ = y.shape
        q = _lin(y, W, p + "self_attn.q_proj", act_dtype).view(B, T, n_heads, hd).transpose(1, 2)
        k = _lin(y, W, p + "self_attn.k_proj", act_dtype).view(B, T, n_heads, hd).transpose(1, 2)
        v = _lin(y, W, p + "self_attn.v_proj", act_dtype).view(B, T, n_heads, hd).transpose(1, 2)
        a = r(torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v)  # :282-337 (SDPA: one fused op)
        a = a.transpose(1, 2).reshape(B, T, D)
        h = r(res + _lin(a, W, p + "self_attn.out_proj", act_dtype))
        res = h
        y = r(F.layer_norm(h, (D,), W[p + "layer_norm2.weight"].float(), W[p + "layer_norm2.bias"].float(), eps))
        y = _lin(y, W, p + "mlp.fc1", act_dtype)
        y = r(y * r(torch.sigmoid(r(1.702 * y))))                               # quick_gelu
        h = r(res + _lin(y, W, p + "mlp.fc2", act_dtype))
        hidden_states.append(h)
    return hidden_states[select_layer][:, 1:]                                  # clip_encoder.py:58,61


def encode_images(W: Dict[str, torch.Tensor], pixels: torch.Tensor, n_heads: int, act_dtype=None) -> torch.Tensor:
    feats = clip_tower(W, pixels, n_heads, act_dtype=act_dtype)
    y = _lin(feats, W, "model.mm_projector.0", act_dtype)
    y = _r(F.gelu(y), act_dtype)
    y = _lin(y, W, "model.mm_projector.2", act_dtype)
    B, T, Fd = y.shape
    g = int(math.sqrt(T))
    y = y.view(B, g, g, Fd).permute(0, 3, 1, 2)
    y = _r(F.avg_pool2d(y, 2, 2), act_dtype)
    return y.flatten(2).transpose(1, 2).contiguous()                           # [B, T/4, F]


def prepare_inputs_embeds(W, input_ids: torch.Tensor, image_features: Optional[List[torch.Tensor]],
                          max_length: Optional[int] = None) -> torch.Tensor:
    """Batch-1 restatement of llava_arch.py:208-343: splice image features at each -200."""
    emb = W["model.embed_tokens.weight"].float()
    ids = input_ids.reshape(-1)
    if image_features is None:
        return emb[ids]
    where = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
    if not where:                                                              # :246-254 (hazard H7)
        return emb[ids]
    parts, prev, img_i = [], -1, 0
    for w in where + [ids.numel()]:
        parts.append(emb[ids[prev + 1:w]])
        if w < ids.numel():
            parts.append(image_features[img_i].float())
            img_i += 1
        prev = w
    out = torch.cat(parts, 0)
    return out[:max_length] if max_length is not None else out


def _rms(x, w, eps, act_dtype=None):
    xf = x.float()
    # modeling_qwen2.py:249-263: normalise in fp32, cast back to the input dtype, THEN multiply by the weight
    return _r(w.float() * _r(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps), act_dtype), act_dtype)


def qwen2_forward(W, x: torch.Tensor, n_heads: int, n_kv: int, theta: float, eps: float,
                  cache: Optional[list] = None, pos0: int = 0, act_dtype=None):
    """x [T, H] -> final hidden [T, H]; cache (list of (k, v) per layer) is extended in place."""
    r = lambda t: _r(t, act_dtype)
    T, H = x.shape
    hd = H // n_heads
    n_layers = 1 + max(int(k.split("model.layers.")[1].split(".")[0]) for k in W if k.startswith("model.layers."))
    inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    pos = torch.arange(pos0, pos0 + T).float()
    ang = pos[:, None] * inv_freq[None, :]
    emb = torch.cat([ang, ang], -1)
    cos, sin = r(emb.cos())[:, None, :], r(emb.sin())[:, None, :]              # cast to the activation dtype (:51-114)

    def rope(t):
        half = hd // 2
        rot = torch.cat([-t[..., half:], t[..., :half]], -1)
        return r(r(t * cos) + r(rot * sin))

    h = x.float()
    for i in range(n_layers):
        p = f"model.layers.{i}."
        y = _rms(h, W[p + "input_layernorm.weight"], eps, act_dtype)
        q = rope(_lin(y, W, p + "self_attn.q_proj", act_dtype).view(T, n_heads, hd))
        k = rope(_lin(y, W, p + "self_attn.k_proj", act_dtype).view(T, n_kv, hd))
        v = _lin(y, W, p + "self_attn.v_proj", act_dtype).view(T, n_kv, hd)
        if cache is not None:
            if len(cache) <= i:
                cache.append((k, v))
            else:
                cache[i] = (torch.cat([cache[i][0], k], 0), torch.cat([cache[i][1], v], 0))
            k, v = cache[i]
        S = k.shape[0]
        g = n_heads // n_kv
        kk = k.repeat_interleave(g, dim=1).transpose(0, 1)                     # [Hq, S, hd]
        vv = v.repeat_interleave(g, dim=1).transpose(0, 1)
        s = q.transpose(0, 1) @ kk.transpose(-1, -2) * hd ** -0.5             # [Hq, T, S]
        mask = torch.arange(S)[None, :] > (torch.arange(T)[:, None] + (S - T))
        s = s.masked_fill(mask, float("-inf"))
        a = r(torch.softmax(s, -1) @ vv).transpose(0, 1).reshape(T, H)
        h = r(h + _lin(a, W, p + "self_attn.o_proj", act_dtype))
        y = _rms(h, W[p + "post_attention_layernorm.weight"], eps, act_dtype)
        y = r(r(F.silu(_lin(y, W, p + "mlp.gate_proj", act_dtype))) * _lin(y, W, p + "mlp.up_proj", act_dtype))
        h = r(h + _lin(y, W, p + "mlp.down_proj", act_dtype))
    return _rms(h, W["model.norm.weight"], eps, act_dtype)


def greedy_generate(W, inputs_embeds: torch.Tensor, max_new_tokens: int, n_heads: int, n_kv: int,
                    theta: float = 1e6, eps: float = 1e-6, eos_token_id: Optional[int] = None,
                    return_logits: bool = False, act_dtype=None):
    """Prefill + cached greedy decode; returns the NEW ids only (and per-step logits)."""
    cache: list = []
    lm = W["lm_head.weight"].float()
    emb = W["model.embed_tokens.weight"].float()
    h = qwen2_forward(W, inputs_embeds, n_heads, n_kv, theta, eps, cache, 0, act_dtype)
    pos = inputs_embeds.shape[0]
    ids, all_logits = [], []
    logits = _r(h[-1] @ lm.t(), act_dtype)
    for _ in range(max_new_tokens):
        all_logits.append(logits)
        nxt = int(torch.argmax(logits))
        ids.append(nxt)
        if eos_token_id is not None and nxt == eos_token_id:
            break
        h = qwen2_forward(W, emb[nxt][None, :], n_heads, n_kv, theta, eps, cache, pos, act_dtype)
        pos += 1
        logits = _r(h[-1] @ lm.t(), act_dtype)
    out = torch.tensor([ids], dtype=torch.long)
    return (out, torch.stack(all_logits)) if return_logits else out
