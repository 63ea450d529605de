"""LongVA (LLaVA-Qwen2) model object on the sm_100a engines — the call surface the streaming
driver uses:

  encode_images(images)                              longva/model/llava_arch.py:179-184
  prepare_inputs_embeddings_for_multimodal(...)      longva/model/llava_arch.py:208-343
  generate_with_image_embedding(inputs, image_embeddings, modalities, **kw)
                                                     longva/model/language_model/llava_qwen.py:137-155
  get_model().embed_tokens(ids), config.mm_use_im_start_end, device

No torch.nn arithmetic: the vision tower runs in streamchat_b200.vision.VisionEncoder and the
language model in streamchat_b200.qwen2.Qwen2Engine (hand-written CUDA behind the C-ABI).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import torch

from ... import ops
from ...qwen2 import Qwen2Config, Qwen2Engine
from ...vision import ClipVisionConfig, VisionEncoder
from ..constants import IMAGE_TOKEN_INDEX


@dataclass
class LlavaQwenConfig:
    llm: Qwen2Config = field(default_factory=Qwen2Config)
    vision: ClipVisionConfig = field(default_factory=ClipVisionConfig)
    model_type: str = "llava_qwen"
    mm_projector_type: str = "mlp2x_gelu"
    mm_resampler_type: str = "spatial_pool"
    mm_spatial_pool_mode: str = "average"
    mm_spatial_pool_stride: int = 2
    mm_use_im_start_end: bool = False
    mm_use_im_patch_token: bool = False
    tokenizer_model_max_length: Optional[int] = None
    max_context: int = 32768      # KV-cache slot length: an answer prompt carries 13-18 K tokens (short memory + 3 x 5 760-token nodes)
    max_batch: int = 32           # memory captions / summaries decoded together, one weight pass per step
    # generation stops at these ids (generation_config.json / config.json of the checkpoint, read by
    # load_pretrained_model) — what HF generate() does by default under the reference's super().generate(),
    # llava_qwen.py:155.  None (random-init synthetic models): every generation runs its full budget.
    eos_token_id: Optional[Union[int, List[int]]] = None
    pad_token_id: Optional[int] = None

    @property
    def hidden_size(self) -> int:
        return self.llm.hidden_size

    @property
    def vocab_size(self) -> int:
        return self.llm.vocab_size


class _Embedding:
    """`model.get_model().embed_tokens` (inference_streaming_longva_v2.py:177)."""

    def __init__(self, table: torch.Tensor):
        self.weight = table

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(device=self.weight.device, dtype=torch.int64)
        out = ops.gather_rows(self.weight, ids.reshape(-1).contiguous())
        return out.view(*ids.shape, self.weight.shape[1])


class _Inner:
    def __init__(self, outer: "LlavaQwenForCausalLM"):
        self._outer = outer
        self.embed_tokens = _Embedding(outer.llm.embed)

    def get_vision_tower(self):
        return self._outer.vision


class LlavaQwenForCausalLM:
    def __init__(self, config: LlavaQwenConfig, weights: Dict[str, torch.Tensor], device):
        if config.mm_projector_type != "mlp2x_gelu" or config.mm_resampler_type != "spatial_pool" \
                or config.mm_spatial_pool_mode != "average" or config.mm_spatial_pool_stride != 2:
            raise NotImplementedError("only LongVA's mlp2x_gelu projector + 2x2 average spatial pool are implemented")
        self.config = config
        self.device = torch.device(device)
        self.dtype = torch.float16
        self.vision = VisionEncoder(weights, config.vision, self.device)
        self.llm = Qwen2Engine(weights, config.llm, self.device, max_ctx=config.max_context,
                               max_batch=config.max_batch)
        self._inner = _Inner(self)

    # ---- reference surface ---------------------------------------------------------------
    def get_model(self):
        return self._inner

    def get_vision_tower(self):
        return self.vision

    def eval(self):
        return self

    @torch.no_grad()
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        return self.vision.encode_images(images.to(self.device))

    def prepare_inputs_embeddings_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values,
                                                 labels, image_features, modalities=["image"]):
        """Batch-1 splice of image features at each IMAGE_TOKEN_INDEX (llava_arch.py:208-343).
        Returns the reference's 6-tuple with attention_mask / position_ids None."""
        if input_ids.dim() == 2:
            if input_ids.shape[0] != 1:
                raise NotImplementedError("the streaming path generates with batch size 1")
            ids = input_ids[0]
        else:
            ids = input_ids
        ids = ids.to(self.device)
        embed = self._inner.embed_tokens
        where = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        if not where:
            # no <image> in the prompt: the visual features are silently dropped (hazard H7,
            # llava_arch.py:246-254)
            x = embed(ids)
        else:
            if len(image_features) < len(where):
                raise ValueError(f"{len(where)} image tokens but {len(image_features)} image feature tensors")
            text_ids = ids[ids != IMAGE_TOKEN_INDEX]
            text = embed(text_ids) if text_ids.numel() else torch.empty((0, self.config.hidden_size),
                                                                        dtype=torch.float16, device=self.device)
            parts, prev, used = [], -1, 0
            for n_img, w in enumerate(where + [ids.numel()]):
                span = w - prev - 1
                parts.append(text[used:used + span])
                used += span
                if w < ids.numel():
                    parts.append(image_features[n_img].to(device=self.device, dtype=torch.float16))
                prev = w
            x = torch.cat(parts, 0)
        max_len = self.config.tokenizer_model_max_length
        if max_len is not None:
            x = x[:max_len]
        return None, None, None, past_key_values, x.unsqueeze(0), None

    @torch.no_grad()
    def generate_with_image_embedding(self, inputs: Optional[torch.Tensor] = None,
                                      image_embeddings: Optional[List[torch.Tensor]] = None,
                                      modalities: Optional[List[str]] = ["image"], **kwargs) -> torch.Tensor:
        kwargs.pop("position_ids", None)
        kwargs.pop("attention_mask", None)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")
        if image_embeddings is not None:
            (_, _, _, _, inputs_embeds, _) = self.prepare_inputs_embeddings_for_multimodal(
                inputs, None, None, None, None, image_embeddings, modalities)
        else:
            inputs_embeds = self._inner.embed_tokens(inputs.to(self.device))
        kwargs.pop("use_cache", None)        # accepted and ignored: HF forces caching here (fact F6)
        num_beams = kwargs.pop("num_beams", 1)
        if num_beams not in (None, 1):
            raise NotImplementedError("beam search is not implemented (the driver uses num_beams=1)")
        do_sample = bool(kwargs.pop("do_sample", False))
        temperature = kwargs.pop("temperature", 1.0)
        temperature = 1.0 if temperature is None else float(temperature)
        return self.llm.generate(inputs_embeds[0], max_new_tokens=int(kwargs.pop("max_new_tokens", 20)),
                                 do_sample=do_sample, temperature=temperature,
                                 top_k=kwargs.pop("top_k", 50), top_p=kwargs.pop("top_p", None),
                                 eos_token_id=kwargs.pop("eos_token_id", self.config.eos_token_id))

    @torch.no_grad()
    def generate_requests(self, requests, modalities: Optional[List[str]] = ["image"], eos_token_id="config"):
        """Greedy generation for several independent requests `(input_ids, image_embeddings | None,
        max_new_tokens)` — e.g. the per-chunk captions of one memory update (utiles.py:539-559) plus a
        pending text summary (utiles.py:594-605).  Each request yields exactly the ids the
        one-at-a-time `generate_with_image_embedding` call yields; they are decoded in lock-step so every
        decode step streams the weights once for all of them."""
        if isinstance(eos_token_id, str):      # "config": the checkpoint's own stop ids (HF generate's default)
            eos_token_id = self.config.eos_token_id
        out: List[Optional[torch.Tensor]] = [None] * len(requests)
        n_batches = -(-len(requests) // self.llm.max_batch)
        B = -(-len(requests) // n_batches) if requests else 1          # even batches: 40 -> 20 + 20
        for i in range(0, len(requests), B):
            embeds, budgets = [], []
            for ids, img, max_new in requests[i:i + B]:
                if img is not None:
                    x = self.prepare_inputs_embeddings_for_multimodal(ids, None, None, None, None, img, modalities)[4][0]
                else:
                    x = self._inner.embed_tokens(ids.to(self.device))[0]
                embeds.append(x)
                budgets.append(int(max_new))
            res = self.llm.generate_batch(embeds, budgets, eos_token_id)
            out[i:i + len(res)] = res
        return out

    def generate_with_image_embedding_batch(self, inputs: torch.Tensor, image_embeddings_list: List[torch.Tensor],
                                            modalities: Optional[List[str]] = ["image"], max_new_tokens: int = 20,
                                            eos_token_id="config") -> List[torch.Tensor]:
        return self.generate_requests([(inputs, [e], max_new_tokens) for e in image_embeddings_list],
                                      modalities, eos_token_id)
