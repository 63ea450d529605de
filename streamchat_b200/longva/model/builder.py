"""`load_pretrained_model` (reference longva/model/builder.py:27-285, qwen branch :202-213).

Returns (tokenizer, model, image_processor, context_len) like the reference.  Two sources:
  * a local HF checkpoint directory (config.json + *.safetensors [+ tokenizer files]) — weights are
    read with safetensors and handed to the sm_100a engines;
  * "synthetic:<name>" — random-init weights of a named architecture (there are no checkpoints or
    network in the build/bench environment; BASELINE configs are random-init), with a byte-level
    stand-in tokenizer.  Names: longva7b (CLIP ViT-L/14-336 + Qwen2-7B dims), tiny.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from ...qwen2 import Qwen2Config
from ...vision import VT, ClipVisionConfig
from .llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM

SYNTHETIC = {
    "longva7b": LlavaQwenConfig(),
    "tiny": LlavaQwenConfig(
        llm=Qwen2Config(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                        num_key_value_heads=1, vocab_size=320),
        vision=ClipVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                                num_attention_heads=2, image_size=56, patch_size=14),
        max_context=2048),
}


class ByteTokenizer:
    """Stand-in tokenizer for synthetic (random-init) checkpoints.  Whitespace-separated words are
    one token each: a word of the form t<id> maps back to that id (so decode -> encode is the
    identity on generated text, like a real tokenizer's ~1 token per sub-word), any other word is
    hashed into the vocabulary."""
    bos_token_id = None
    eos_token_id = None
    pad_token_id = 0

    def __init__(self, vocab_size: int):
        self.vocab_size = vocab_size

    def _word_id(self, word: str) -> int:
        if len(word) > 1 and word[0] == "t" and word[1:].isdigit() and int(word[1:]) < self.vocab_size:
            return int(word[1:])
        h = 2166136261
        for b in word.encode("utf-8"):
            h = ((h ^ b) * 16777619) & 0xFFFFFFFF
        return 3 + h % (self.vocab_size - 3)

    def __call__(self, text, **kw):
        class _Enc(dict):
            __getattr__ = dict.get
        ids = [self._word_id(w) for w in text.split()]
        if kw.get("return_tensors") == "pt":
            return _Enc(input_ids=torch.tensor([ids], dtype=torch.long),
                        attention_mask=torch.ones((1, len(ids)), dtype=torch.long))
        return _Enc(input_ids=ids, attention_mask=[1] * len(ids))

    def batch_decode(self, ids, skip_special_tokens=True):
        rows = ids.tolist() if hasattr(ids, "tolist") else ids
        return [" ".join(f"t{int(t)}" for t in row) for row in rows]

    def add_tokens(self, *a, **k):
        return 0


class SyntheticImageProcessor:
    """CLIPImageProcessor contract for already-sized frames: `.preprocess(img, return_tensors="pt")
    ["pixel_values"]` (utiles.py:77).  Real checkpoints use transformers.CLIPImageProcessor."""
    image_mean = [0.48145466, 0.4578275, 0.40821073]
    image_std = [0.26862954, 0.26130258, 0.27577711]

    def __init__(self, size: int):
        self.crop_size = {"height": size, "width": size}
        self.size = {"shortest_edge": size}

    def preprocess(self, image, return_tensors="pt"):
        import numpy as np
        arr = np.asarray(image, dtype=np.float32) / 255.0
        if arr.ndim == 2:
            arr = np.stack([arr] * 3, -1)
        s = self.crop_size["height"]
        if arr.shape[0] != s or arr.shape[1] != s:
            raise ValueError(f"SyntheticImageProcessor expects {s}x{s} frames")
        t = torch.from_numpy(arr).permute(2, 0, 1)
        mean = torch.tensor(self.image_mean).view(3, 1, 1)
        std = torch.tensor(self.image_std).view(3, 1, 1)
        return {"pixel_values": ((t - mean) / std).unsqueeze(0)}


def random_weights(cfg: LlavaQwenConfig, device, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """normal(0, std) fp16 matrices, LN/RMS gamma = 1, beta = 0 — generated on `device`
    (SURVEY.md §8d 'Concrete synthetic input')."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def mat(name, *shape):
        W[name] = (torch.randn(shape, generator=g, device=dev, dtype=torch.float32) * std).half()

    def ones(name, n):
        W[name] = torch.ones(n, dtype=torch.float16, device=dev)

    def zeros(name, n):
        W[name] = torch.zeros(n, dtype=torch.float16, device=dev)

    v, l = cfg.vision, cfg.llm
    D, I = v.hidden_size, v.intermediate_size
    mat(VT + "embeddings.class_embedding", D)
    mat(VT + "embeddings.patch_embedding.weight", D, v.num_channels, v.patch_size, v.patch_size)
    mat(VT + "embeddings.position_embedding.weight", v.tokens, D)
    ones(VT + "pre_layrnorm.weight", D); zeros(VT + "pre_layrnorm.bias", D)
    for i in range(v.num_hidden_layers):
        p = f"{VT}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            mat(p + f"self_attn.{n}.weight", D, D); mat(p + f"self_attn.{n}.bias", D)
        ones(p + "layer_norm1.weight", D); zeros(p + "layer_norm1.bias", D)
        ones(p + "layer_norm2.weight", D); zeros(p + "layer_norm2.bias", D)
        mat(p + "mlp.fc1.weight", I, D); mat(p + "mlp.fc1.bias", I)
        mat(p + "mlp.fc2.weight", D, I); mat(p + "mlp.fc2.bias", D)
    H, LI, hd = l.hidden_size, l.intermediate_size, l.head_dim
    mat("model.mm_projector.0.weight", H, D); mat("model.mm_projector.0.bias", H)
    mat("model.mm_projector.2.weight", H, H); mat("model.mm_projector.2.bias", H)
    mat("model.embed_tokens.weight", l.vocab_size, H)
    mat("lm_head.weight", l.vocab_size, H)
    ones("model.norm.weight", H)
    for i in range(l.num_hidden_layers):
        p = f"model.layers.{i}."
        mat(p + "self_attn.q_proj.weight", l.num_attention_heads * hd, H); mat(p + "self_attn.q_proj.bias", l.num_attention_heads * hd)
        mat(p + "self_attn.k_proj.weight", l.num_key_value_heads * hd, H); mat(p + "self_attn.k_proj.bias", l.num_key_value_heads * hd)
        mat(p + "self_attn.v_proj.weight", l.num_key_value_heads * hd, H); mat(p + "self_attn.v_proj.bias", l.num_key_value_heads * hd)
        mat(p + "self_attn.o_proj.weight", H, l.num_attention_heads * hd)
        mat(p + "mlp.gate_proj.weight", LI, H); mat(p + "mlp.up_proj.weight", LI, H); mat(p + "mlp.down_proj.weight", H, LI)
        ones(p + "input_layernorm.weight", H); ones(p + "post_attention_layernorm.weight", H)
    return W


def _tower_dir(model_path: str, c: dict) -> Optional[str]:
    """Directory of the CLIP tower named by `mm_vision_tower`: absolute, or relative to the checkpoint (a hub id
    such as "openai/clip-vit-large-patch14-336" resolves to neither and means: ViT-L/14-336 defaults, tower
    weights inside the main checkpoint — LongVA-7B was trained with the tower unfrozen)."""
    vt = c.get("mm_vision_tower")
    if not vt:
        return None
    for cand in (vt, os.path.join(model_path, vt)):
        if os.path.exists(os.path.join(cand, "config.json")):
            return cand
    return None


def _config_from_hf(path: str) -> LlavaQwenConfig:
    c = json.load(open(os.path.join(path, "config.json")))
    # transformers >= 5 stores theta under rope_parameters; 4.x (the reference's pin) at the top level
    theta = c.get("rope_theta", (c.get("rope_parameters") or {}).get("rope_theta", 1e6))
    llm = Qwen2Config(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
                      num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                      num_key_value_heads=c.get("num_key_value_heads", c["num_attention_heads"]),
                      vocab_size=c["vocab_size"], rope_theta=float(theta),
                      rms_norm_eps=float(c.get("rms_norm_eps", 1e-6)))
    vision = ClipVisionConfig(select_layer=c.get("mm_vision_select_layer", -2))
    vt = _tower_dir(path, c)
    if vt:
        vc = json.load(open(os.path.join(vt, "config.json")))
        vc = vc.get("vision_config", vc)
        vision = ClipVisionConfig(hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
                                  num_hidden_layers=vc["num_hidden_layers"],
                                  num_attention_heads=vc["num_attention_heads"], image_size=vc["image_size"],
                                  patch_size=vc["patch_size"], layer_norm_eps=vc.get("layer_norm_eps", 1e-5),
                                  select_layer=c.get("mm_vision_select_layer", -2))
    # stop ids: generation_config.json wins over config.json, as in HF generate() (the reference reaches it through
    # super().generate(), llava_qwen.py:155; for LongVA-7B-DPO this is <|im_end|> / <|endoftext|>)
    eos, pad = c.get("eos_token_id"), c.get("pad_token_id")
    gpath = os.path.join(path, "generation_config.json")
    if os.path.exists(gpath):
        gc = json.load(open(gpath))
        eos = gc.get("eos_token_id", eos)
        pad = gc.get("pad_token_id", pad)
    return LlavaQwenConfig(llm=llm, vision=vision, eos_token_id=eos, pad_token_id=pad,
                           mm_projector_type=c.get("mm_projector_type", "mlp2x_gelu"),
                           mm_resampler_type=c.get("mm_resampler_type", "spatial_pool"),
                           mm_spatial_pool_mode=c.get("mm_spatial_pool_mode", "average"),
                           mm_spatial_pool_stride=c.get("mm_spatial_pool_stride", 2),
                           mm_use_im_start_end=c.get("mm_use_im_start_end", False),
                           # the reference reads this flag with a default of True (builder.py:261)
                           mm_use_im_patch_token=c.get("mm_use_im_patch_token", True),
                           tokenizer_model_max_length=c.get("tokenizer_model_max_length"))


def _load_safetensors(path: str) -> Dict[str, torch.Tensor]:
    """Every tensor of the checkpoint's *.safetensors shards (model.safetensors or the files an index names)."""
    from safetensors import safe_open
    W: Dict[str, torch.Tensor] = {}
    idx = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(idx):
        files = sorted(set(json.load(open(idx))["weight_map"].values()))
    else:
        files = sorted(fn for fn in os.listdir(path) if fn.endswith(".safetensors"))
    for fn in files:
        with safe_open(os.path.join(path, fn), framework="pt", device="cpu") as f:
            for k in f.keys():
                W[k] = f.get_tensor(k)
    if not W:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    return W


def resize_token_embeddings(weights: Dict[str, torch.Tensor], llm: Qwen2Config, n_tokens: int) -> None:
    """`model.resize_token_embeddings(len(tokenizer))` of the reference loader (builder.py:266), applied to the weight
    dict in place.  Fewer tokens than rows: both tables are CUT to len(tokenizer) rows, exactly as HF does (the
    logits the reference arg-maxes over have len(tokenizer) entries).  More tokens than rows: rows are appended
    holding the mean of the existing rows (HF draws them at random: they belong to tokens no prompt on the path
    contains unless the caller types the literal <im_patch> / <im_start> markers).  The decode kernels stream the
    lm_head in 16-row tiles, so a cut table is padded back to a multiple of 16 with copies of row 0 — a copy can
    never win the lowest-index-wins arg-max over the original; `logit_vocab` records the true width."""
    emb = weights["model.embed_tokens.weight"]
    head = weights.get("lm_head.weight", emb)
    rows = emb.shape[0]
    if n_tokens == rows:
        return
    if n_tokens > rows:
        grow = n_tokens - rows
        emb = torch.cat([emb, emb.float().mean(0, keepdim=True).to(emb.dtype).expand(grow, -1)], 0)
        head = torch.cat([head, head.float().mean(0, keepdim=True).to(head.dtype).expand(grow, -1)], 0)
    else:
        emb, head = emb[:n_tokens], head[:n_tokens]
    pad = (-n_tokens) % 16
    if pad:
        head = torch.cat([head, head[:1].expand(pad, -1)], 0)
    weights["model.embed_tokens.weight"] = emb.contiguous()
    weights["lm_head.weight"] = head.contiguous()
    llm.vocab_size = head.shape[0]
    llm.logit_vocab = n_tokens


def _device_from_map(device_map, device) -> torch.device:
    if isinstance(device_map, str) and device_map.startswith("cuda"):
        return torch.device(device_map)
    if isinstance(device_map, dict) and device_map:
        return torch.device(next(iter(device_map.values())))
    return torch.device(device if device != "cuda" else "cuda:0")


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False,
                          device_map="auto", attn_implementation="flash_attention_2", customized_config=None,
                          overwrite_config=None, device="cuda", seed: int = 0, **kwargs):
    if load_8bit or load_4bit:
        raise NotImplementedError("quantised loading is not part of the hot path")
    if "qwen" not in model_name.lower():
        raise NotImplementedError(f"only the llava_qwen (LongVA) branch is implemented, got {model_name!r}")
    dev = _device_from_map(device_map, device)
    if str(model_path).startswith("synthetic:"):
        name = str(model_path).split(":", 1)[1]
        cfg = SYNTHETIC[name]
        weights = random_weights(cfg, dev, seed=seed)
        tokenizer = ByteTokenizer(cfg.llm.vocab_size)
        image_processor = SyntheticImageProcessor(cfg.vision.image_size)
    else:
        from transformers import AutoTokenizer, CLIPImageProcessor
        from ..constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN
        cfg = _config_from_hf(model_path)
        weights = _load_safetensors(model_path)
        hf_cfg = json.load(open(os.path.join(model_path, "config.json")))
        vt = _tower_dir(model_path, hf_cfg)
        if not any(k.startswith(VT) for k in weights):
            # frozen-tower checkpoints keep the CLIP weights in the tower's own directory (clip_encoder.py:40-41)
            if vt is None:
                raise FileNotFoundError(f"{model_path} holds no vision-tower weights and mm_vision_tower="
                                        f"{hf_cfg.get('mm_vision_tower')!r} is not a local directory")
            for k, v in _load_safetensors(vt).items():
                weights["model.vision_tower.vision_tower." + k] = v
        tokenizer = AutoTokenizer.from_pretrained(model_path)
        # builder.py:258-266: marker tokens are added, then the embedding tables follow len(tokenizer)
        if cfg.mm_use_im_patch_token:
            tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
        if cfg.mm_use_im_start_end:
            tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
        resize_token_embeddings(weights, cfg.llm, len(tokenizer))
        image_processor = CLIPImageProcessor.from_pretrained(vt if vt else hf_cfg.get("mm_vision_tower", model_path))
    if overwrite_config:
        for k, v in overwrite_config.items():
            setattr(cfg, k, v)
    model = LlavaQwenForCausalLM(cfg, weights, dev)
    del weights
    # builder.py:277-284 reports the checkpoint's max_sequence_length / max_position_embeddings; what bounds a
    # generation HERE is the KV-cache slot, so that is what callers get
    context_len = getattr(cfg, "max_context", 2048)
    return tokenizer, model, image_processor, context_len
