"""CLIP vision tower -> mm_projector -> 2x2 spatial pool on the sm_100a kernels.

Host-side orchestration of the reference's `encode_images` (longva/model/llava_arch.py:179-184):
  CLIPVisionTower.forward / feature_select   longva/model/multimodal_encoder/clip_encoder.py:46-79
  mm_projector (mlp2x_gelu)                   longva/model/multimodal_projector/builder.py:41-48
  SpatialPool (average, stride 2)             longva/model/multimodal_resampler/spatial_pool.py:23-32
Operator order of the HF CLIP math: SURVEY.md Appendix B.1 / B.2.

Per frame batch: im2col -> patch GEMM (tcgen05) -> token assembly + pre-LN -> L x { LN, fused QKV
GEMM, fused tcgen05 attention, out-proj GEMM (+residual), LN, fc1 GEMM (+quick-GELU), fc2 GEMM
(+residual) } -> projector GEMM-1 (+GELU) -> 2x2 average pool -> projector GEMM-2.
Only the layers that feed hidden_states[select_layer] are computed (23 of 24 for ViT-L/14), and
the 2x2 mean is taken before Linear-2 (it commutes with Linear-2 exactly in real arithmetic, B.2):
373.96 GFLOP per 336x336 frame instead of the reference's 385.06 + the unused 24th layer.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops

VT = "model.vision_tower.vision_tower.vision_model."


@dataclass
class ClipVisionConfig:
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    layer_norm_eps: float = 1e-5
    select_layer: int = -2            # mm_vision_select_layer
    num_channels: int = 3

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def layers_needed(self) -> int:
        # hidden_states has num_hidden_layers + 1 entries; entry i is the output of layer i
        idx = self.select_layer if self.select_layer >= 0 else self.num_hidden_layers + 1 + self.select_layer
        return idx


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class VisionEncoder:
    """Weights live on one device in fp16, pre-fused for the kernels (QKV concatenated, patch
    kernel flattened and K-padded to a TMA-legal row stride)."""

    def __init__(self, weights: Dict[str, torch.Tensor], cfg: ClipVisionConfig, device: torch.device,
                 max_frames_per_pass: int = 128):
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_frames = max_frames_per_pass
        D = cfg.hidden_size
        hd = D // cfg.num_attention_heads
        if hd not in (64, 128):
            raise ValueError(f"CLIP head_dim {hd} unsupported by the attention kernel (64 or 128)")

        def w(name):
            return weights[name].to(device=self.device, dtype=torch.float16).contiguous()

        kdim = cfg.num_channels * cfg.patch_size ** 2
        self.kpad = _round_up(kdim, 8)
        pw = torch.zeros((D, self.kpad), dtype=torch.float16, device=self.device)
        pw[:, :kdim] = w(VT + "embeddings.patch_embedding.weight").reshape(D, kdim)
        self.patch_w = pw
        self.cls = w(VT + "embeddings.class_embedding")
        self.pos = w(VT + "embeddings.position_embedding.weight")
        self.pre_g, self.pre_b = w(VT + "pre_layrnorm.weight"), w(VT + "pre_layrnorm.bias")
        self.layers = []
        for i in range(cfg.layers_needed):
            p = f"{VT}encoder.layers.{i}."
            self.layers.append(dict(
                ln1_g=w(p + "layer_norm1.weight"), ln1_b=w(p + "layer_norm1.bias"),
                qkv_w=torch.cat([w(p + "self_attn.q_proj.weight"), w(p + "self_attn.k_proj.weight"),
                                 w(p + "self_attn.v_proj.weight")], 0).contiguous(),
                qkv_b=torch.cat([w(p + "self_attn.q_proj.bias"), w(p + "self_attn.k_proj.bias"),
                                 w(p + "self_attn.v_proj.bias")], 0).contiguous(),
                out_w=w(p + "self_attn.out_proj.weight"), out_b=w(p + "self_attn.out_proj.bias"),
                ln2_g=w(p + "layer_norm2.weight"), ln2_b=w(p + "layer_norm2.bias"),
                fc1_w=w(p + "mlp.fc1.weight"), fc1_b=w(p + "mlp.fc1.bias"),
                fc2_w=w(p + "mlp.fc2.weight"), fc2_b=w(p + "mlp.fc2.bias")))
        self.proj1_w, self.proj1_b = w("model.mm_projector.0.weight"), w("model.mm_projector.0.bias")
        self.proj2_w, self.proj2_b = w("model.mm_projector.2.weight"), w("model.mm_projector.2.bias")
        self.out_dim = self.proj2_w.shape[0]
        # native entry point: the whole pass below as ONE C call (scb_vit_encode) over the same kernels
        self._native = ops.VitEncoder(
            layers=self.layers, patch_w=self.patch_w, cls=self.cls, pos=self.pos, pre_g=self.pre_g, pre_b=self.pre_b,
            proj1_w=self.proj1_w, proj1_b=self.proj1_b, proj2_w=self.proj2_w, proj2_b=self.proj2_b,
            n_heads=cfg.num_attention_heads, image_size=cfg.image_size, patch=cfg.patch_size,
            channels=cfg.num_channels, ln_eps=cfg.layer_norm_eps)

    # FLOPs of the formulation executed, per frame (for bench.py's roofline)
    def flops_per_frame(self) -> float:
        c = self.cfg
        D, I, T, Hh = c.hidden_size, c.intermediate_size, c.tokens, self.out_dim
        per_layer = 2 * T * D * 3 * D + 2 * T * D * D + 4 * T * T * D + 4 * T * D * I
        patch = 2 * (T - 1) * (c.num_channels * c.patch_size ** 2) * D
        proj = 2 * T * D * Hh + 2 * ((c.grid // 2) ** 2) * Hh * Hh
        return c.layers_needed * per_layer + patch + proj

    def tower(self, pixels: torch.Tensor) -> torch.Tensor:
        """[N,3,S,S] fp16 -> hidden_states[select_layer] as [N*tokens, D] (CLS rows included)."""
        c = self.cfg
        n = pixels.shape[0]
        D = c.hidden_size
        H = c.num_attention_heads
        hd = D // H
        T = c.tokens
        cols = ops.im2col_patch(pixels, c.patch_size, self.kpad)
        patch = ops.linear(cols, self.patch_w)
        x = ops.vit_embed_layernorm(patch, self.cls, self.pos, self.pre_g, self.pre_b, c.layer_norm_eps, n)
        scale = hd ** -0.5
        for ly in self.layers:
            y = ops.layernorm(x, ly["ln1_g"], ly["ln1_b"], c.layer_norm_eps)
            qkv = ops.linear(y, ly["qkv_w"], bias=ly["qkv_b"]).view(n, T, 3, H, hd)
            ctx = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale, causal=False)
            x = ops.linear(ctx.view(n * T, D), ly["out_w"], bias=ly["out_b"], residual=x)
            y = ops.layernorm(x, ly["ln2_g"], ly["ln2_b"], c.layer_norm_eps)
            y = ops.linear(y, ly["fc1_w"], bias=ly["fc1_b"], act=L.ACT_QUICK_GELU)
            x = ops.linear(y, ly["fc2_w"], bias=ly["fc2_b"], residual=x)
        return x

    def _encode_pass(self, pixels: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        n = pixels.shape[0]
        if os.environ.get("SCB_NATIVE_VIT", "1") != "0":
            # one native call issues the launch sequence below from C++ into a reused workspace (no per-op output
            # allocations); SCB_NATIVE_VIT=0 keeps the per-launch Python loop (same kernels, bit-identical results)
            return self._native.encode(pixels.contiguous())
        x = self.tower(pixels)
        y = ops.linear(x, self.proj1_w, bias=self.proj1_b, act=L.ACT_GELU_ERF)      # CLS rows ride along
        y = ops.avgpool2x2_tokens(y, n, c.tokens, 1, c.grid)                         # drops CLS, 2x2 mean
        y = ops.linear(y, self.proj2_w, bias=self.proj2_b)
        return y.view(n, (c.grid // 2) ** 2, self.out_dim)

    @torch.no_grad()
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        """images [N,3,S,S] (cuda; cast to fp16 like the reference's `.half()`) -> [N, (grid/2)^2, H] fp16."""
        if not images.is_cuda:
            raise L.ScbError("encode_images needs CUDA tensors: there is no CPU fallback")
        images = images.to(dtype=torch.float16)
        # even passes (a 200-frame segment runs as 2 x 100 frames, not 128 + 72): every GEMM of a pass then has the same,
        # large M (57 700 rows) — better tile reuse and no small ragged pass
        n = images.shape[0]
        n_pass = max(1, -(-n // self.max_frames))
        size = -(-n // n_pass)
        outs = [self._encode_pass(images[i:i + size]) for i in range(0, n, size)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
