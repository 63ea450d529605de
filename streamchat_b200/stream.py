"""The per-video hot loop of the reference driver as a session object over tensors.

Mirrors, stage by stage, inference_streaming_longva_v2.py (the driver itself stays the
integration target; this class is what bench.py and the tests drive with synthetic frames):
  [A] ingest()         video_reader_thread_with_embedding  :454-531 (the encode tail :520-525:
                       torch.cat(frames).half() -> model.encode_images -> split into [1,144,H])
  [B] update_memory()  updating_memory_buffer              :267-378
  [C] answer()         longva_inference_with_embedding_multi_modal :164-264
Frame decode / PIL preprocessing ([A] head, a-V1) is the "next" row §8f-2 and is not part of this
class: frames arrive as normalised pixel tensors.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from . import memory as M
from .kmeans import weighted_kmeans_feature
from .longva.constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_TOKEN, IMAGE_TOKEN_INDEX
from .longva.conversation import conv_templates
from .longva.mm_utils import tokenizer_image_token

CAPTION_REQUEST = ("Please describe what you see in this video in as much detail as possible from a "
                   "first-person perspective, including the surrounding environment, what objects are there, etc.")


@dataclass
class StreamConfig:
    """Launch-script defaults (inference_streamchat_v0.3.sh:11-19)."""
    chunk_size: int = 40
    num_clusters: int = 5
    interval: int = 10
    short_window: int = 20
    remember_window: int = 5
    tau: float = 5
    conv_mode: str = "qwen_1_5_ego"
    temperature: float = 0.0          # 0 -> greedy (parity runs, hazard H2); the script uses 0.2
    answer_tokens: int = 256
    caption_do_sample: bool = False   # the reference samples captions at T=0.1 (utiles.py:551-552)
    lazy_text: bool = True            # defer greedy captions / summaries and generate them in batches


class HashTextEmbedder:
    """Synthetic stand-in for the BERT-large CLS embedder (mxbai-colbert-large-v1, 1024-d;
    utiles.py:704-708) when no weights exist: a deterministic vector per text with a common positive
    component (sentence embeddings of related captions correlate positively; the reference's search
    raises when no child has a positive similarity, utiles.py:746).
    Callable as `model(tokenizer, text)` — see memory.embed_text."""

    def __init__(self, dim: int = 1024, device="cuda:0"):
        self.dim = dim
        self.device = torch.device(device)

    def __call__(self, tokenizer, text: str) -> torch.Tensor:
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:7], "little")
        g = torch.Generator(device=self.device).manual_seed(seed)
        return torch.randn(self.dim, generator=g, device=self.device, dtype=torch.float32) + 1.0


class StreamingSession:
    def __init__(self, model, tokenizer, embedding_model, embedding_tokenizer=None, summarizer=None,
                 cfg: Optional[StreamConfig] = None):
        self.model = model
        self.summarizer = summarizer if summarizer is not None else model   # 2nd replica in the reference (:700)
        self.tokenizer = tokenizer
        self.embedding_model = embedding_model
        self.embedding_tokenizer = embedding_tokenizer
        self.cfg = cfg or StreamConfig()
        self.long_memory_tree: Optional[list] = None
        self.short_memory_buffer: List[torch.Tensor] = []
        self.frames_seen = 0
        self._caption_ids = self._captioning_prompt_ids()

    def _captioning_prompt_ids(self) -> torch.Tensor:
        # :288-299
        if self.summarizer.config.mm_use_im_start_end:
            qs = DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN + "\n" + CAPTION_REQUEST
        else:
            qs = DEFAULT_IMAGE_TOKEN + "\n" + CAPTION_REQUEST
        conv = conv_templates["qwen_1_5_ego"].copy()
        conv.append_message(conv.roles[0], qs)
        conv.append_message(conv.roles[1], None)
        ids = tokenizer_image_token(conv.get_prompt(), self.tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt")
        return ids.unsqueeze(0)

    # ---- [A] ---------------------------------------------------------------------------------
    @torch.no_grad()
    def ingest(self, frames: torch.Tensor) -> List[torch.Tensor]:
        """frames [n,3,S,S] on the model's device -> list of n tensors [1,144,H] (:520-525)."""
        feats = self.model.encode_images(frames.half())
        self.frames_seen += frames.shape[0]
        return list(torch.split(feats, 1))

    # ---- [B] ---------------------------------------------------------------------------------
    def update_memory(self, buffer_cache: List[torch.Tensor], kmeans_init_indices=None, chunk_nodes=None):
        """`chunk_nodes`: depth-0 nodes captioned elsewhere (one per chunk of `buffer_cache`) — the frame-sharded
        stream builds them on the owning rank and gathers them (distributed.ShardedStreamingSession)."""
        c = self.cfg
        short_window = c.short_window
        if len(buffer_cache) > short_window:                                   # :319-324
            wait_fifo = buffer_cache[-short_window:]
        else:
            short_window = len(buffer_cache)
            wait_fifo = buffer_cache
        remember = min(c.remember_window, len(wait_fifo))                      # :326-329
        probs = M.calculate_forgetting_probabilities(short_window, tau=c.tau)
        self.short_memory_buffer = M.select_data_without_replacement(wait_fifo, probs, remember)
        chunks = [buffer_cache[i:i + c.chunk_size] for i in range(0, len(buffer_cache), c.chunk_size)]
        if chunk_nodes is not None:
            kmeans_chunks = [n.centroids for n in chunk_nodes]                 # already stacked on the owning rank
        else:
            kmeans_chunks = [weighted_kmeans_feature(torch.cat(ch), c.num_clusters)[0] if len(ch) > c.chunk_size
                             else torch.cat(ch) for ch in chunks]              # :347 (guard never true, F4)
        self.long_memory_tree = M.fast_building_memory_tree_summarize_token(
            kmeans_chunks, c.num_clusters, c.interval, self.summarizer, self._caption_ids, self.tokenizer,
            chunks, self.long_memory_tree, caption_do_sample=c.caption_do_sample,
            kmeans_init_indices=kmeans_init_indices, conv_templates=conv_templates, lazy_text=c.lazy_text,
            chunk_nodes=chunk_nodes)
        assert len(self.short_memory_buffer) > 0, "No memory ?"
        return self.long_memory_tree, self.short_memory_buffer

    # ---- [C] ---------------------------------------------------------------------------------
    def build_question_inputs(self, question: str, history_prompt: Optional[str] = None):
        short = torch.cat(self.short_memory_buffer).view(-1, self.short_memory_buffer[0].shape[-1])
        if self.long_memory_tree is not None:
            feats, texts = M.fast_search_tree_multi_modal_with_embedding(
                self.long_memory_tree, question, short, self.embedding_model, self.embedding_tokenizer)
            long_emb = torch.cat(feats, dim=0).view(-1, feats[0].shape[-1])
            most_fine = texts[-1]
            image_embeddings = torch.cat([short, long_emb.to(short.device)], dim=0)
        else:
            image_embeddings, most_fine = short, None
        prm = ("In addition, the text caption memory information articles most relevant to the current problem is "
               "'{most_fine_grad_text}'.         The image information you currently see and recall in the "
               "{image_token} is equally important as the contextual information mentioned earlier.         "
               "Sometimes the contextual information does not contain a direct answer to the question.         "
               "You need to synthesize this information and give an answer to the following question:")
        notion = "DO NOT OUTPUT ANY EXPLANATORY TEXT THAT IS UNCERTAIN ABOUT THE CURRENT QUESTION."
        img = DEFAULT_IMAGE_TOKEN
        if self.model.config.mm_use_im_start_end:
            img = DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN
        if history_prompt is not None:
            if most_fine is not None:
                if self.model.config.mm_use_im_start_end:
                    qs = history_prompt + prm.format(most_fine_grad_text=most_fine, image_token="{image_token}") \
                        + img + "\n" + question + notion
                else:
                    qs = history_prompt + prm.format(most_fine_grad_text=most_fine, image_token=DEFAULT_IMAGE_TOKEN) \
                        + "\n" + question + notion
            else:
                # no <image> without im_start_end: the visual context is silently dropped (H7, :223)
                qs = history_prompt + ((img + "\n") if self.model.config.mm_use_im_start_end else "\n") + question + notion
        else:
            qs = img + "\n" + question + notion
        conv = conv_templates[self.cfg.conv_mode].copy()
        conv.append_message(conv.roles[0], qs)
        conv.append_message(conv.roles[1], None)
        input_ids = tokenizer_image_token(conv.get_prompt(), self.tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt")
        return input_ids.unsqueeze(0), image_embeddings

    @torch.no_grad()
    def answer(self, question: str, history_prompt: Optional[str] = None):
        input_ids, image_embeddings = self.build_question_inputs(question, history_prompt)
        t = self.cfg.temperature
        # Merge summaries of the tree's roots are never read by a search (memory.texts_read_by_search), so with lazy
        # texts they are still pending here; they are independent of the answer and decode in the SAME lock-step
        # batch — one weight pass per step for the answer and the summaries instead of two generation loops
        # (ids are those of the separate calls: rows of a batch are independent).
        ready = []
        if t == 0 and self.cfg.lazy_text and self.long_memory_tree is not None and self.summarizer is self.model \
                and hasattr(self.model, "generate_requests"):
            ready = M.ready_summaries(self.long_memory_tree)
        if ready:
            reqs = [(input_ids, [image_embeddings], self.cfg.answer_tokens)] + [M.summary_request(n) for n in ready]
            outs = self.model.generate_requests(reqs, modalities=["video"])
            output_ids = outs[0]
            for n, ids in zip(ready, outs[1:]):
                n.set_text(self.tokenizer.batch_decode(ids, skip_special_tokens=True)[0].strip())
        else:
            output_ids = self.model.generate_with_image_embedding(
                input_ids, image_embeddings=[image_embeddings], modalities=["video"],
                do_sample=True if t > 0 else False, temperature=t, top_p=None, num_beams=1,
                max_new_tokens=self.cfg.answer_tokens, use_cache=False)
        text = self.tokenizer.batch_decode(output_ids, skip_special_tokens=True)[0].strip()
        return text, output_ids
