"""Qwen2 decoder (LongVA's language model) on the sm_100a kernels: long-context prefill on
tcgen05 GEMM + fused attention, cached greedy decode as a CUDA-graph of weight-streaming GEMVs.

Replaces the third-party math under `generate_with_image_embedding`
(longva/model/language_model/llava_qwen.py:137-155 -> Qwen2ForCausalLM.generate with
inputs_embeds; prefill + DynamicCache decode because HF forces use_cache when inputs_embeds are
given, generation/utils.py:2425-2428).  Operator order: SURVEY.md Appendix B.3.

Data layout in HBM (per engine): fused QKV weight [(Hq+2Hkv)*hd, H] (+bias), o_proj [H, H], fused
gate|up [2I, H], down [H, I], all fp16 row-major (K contiguous); KV cache per layer
[Hkv, max_ctx, hd] fp16 for K (post-RoPE) and V.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops


@dataclass
class Qwen2Config:
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    vocab_size: int = 152064
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    logit_vocab: Optional[int] = None   # true logit width when the lm_head was padded to a 16-row multiple (builder.resize_token_embeddings)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


class Qwen2Engine:
    def __init__(self, weights: Dict[str, torch.Tensor], cfg: Qwen2Config, device: torch.device,
                 max_ctx: int = 16384, decode_splits: int = 37, max_batch: int = 1):
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_ctx = max_ctx
        self.decode_splits = decode_splits
        self.max_batch = max_batch
        self.mega_rows = int(os.environ.get("SCB_DECODE_MEGA_ROWS", "4"))   # widest batch the persistent kernel takes
        # batches this wide run the tcgen05 split-K chain; narrower ones the persistent kernel (<= mega_rows) or, where it
        # is unavailable, the GEMV chain.  Measured per step at LongVA-7B shapes (ctx 5 800): 5 / 7 / 8 rows 4.60 / 4.81 /
        # 4.64 ms on the split-K chain against 5.18 / 5.57 / 5.61 ms on the GEMV chain
        self.wide_min_rows = int(os.environ.get("SCB_WIDE_MIN_ROWS", "5"))
        self.prefill_group_tokens = int(os.environ.get("SCB_PREFILL_GROUP_TOKENS", "30000"))   # rows stacked per prefill pass
        if max_batch < 1 or max_batch > 32:
            raise ValueError('max_batch must be in 1..32 (sequences decoded in lock-step)')
        c = cfg
        if c.head_dim not in (64, 128):
            raise ValueError(f"head_dim {c.head_dim} unsupported (64 or 128)")

        def w(name):
            return weights[name].to(device=self.device, dtype=torch.float16).contiguous()

        self.embed = w("model.embed_tokens.weight")
        # checkpoints saved with tie_word_embeddings carry no lm_head.weight: the head IS the embedding table
        self.lm_head = w("lm_head.weight") if "lm_head.weight" in weights else self.embed
        self.final_norm = w("model.norm.weight")
        self.layers = []
        for i in range(c.num_hidden_layers):
            p = f"model.layers.{i}."
            self.layers.append(dict(
                ln1=w(p + "input_layernorm.weight"),
                qkv_w=torch.cat([w(p + "self_attn.q_proj.weight"), w(p + "self_attn.k_proj.weight"),
                                 w(p + "self_attn.v_proj.weight")], 0).contiguous(),
                qkv_b=torch.cat([w(p + "self_attn.q_proj.bias"), w(p + "self_attn.k_proj.bias"),
                                 w(p + "self_attn.v_proj.bias")], 0).contiguous(),
                o_w=w(p + "self_attn.o_proj.weight"),
                ln2=w(p + "post_attention_layernorm.weight"),
                # gate|up interleaved row-wise (2i = gate_i, 2i+1 = up_i) so the SiLU gate is a pair-wise
                # epilogue of the fused projection (prefill GEMM and decode GEMV alike)
                gu_w=torch.stack([w(p + "mlp.gate_proj.weight"), w(p + "mlp.up_proj.weight")], 1)
                .reshape(2 * c.intermediate_size, c.hidden_size).contiguous(),
                down_w=w(p + "mlp.down_proj.weight")))
        hd = c.head_dim
        # inv_freq exactly as HF computes it (modeling_qwen2.py:51-114): fp32 table
        self.inv_freq = (1.0 / (c.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).to(self.device)
        B = max_batch
        # one KV cache slot per concurrently decoded sequence: [B, Hkv, max_ctx, hd] per layer
        self.k_cache = [torch.zeros((B, c.num_key_value_heads, max_ctx, hd), dtype=torch.float16, device=self.device)
                        for _ in range(c.num_hidden_layers)]
        self.v_cache = [torch.zeros_like(k) for k in self.k_cache]
        self.ctx_lens = [0] * B
        # decode state (device-resident so the step graphs are replayable): [0] positions, [1] kv lengths
        self.state = torch.zeros((2, B), dtype=torch.int32, device=self.device)
        self.next_id = torch.zeros(B, dtype=torch.int64, device=self.device)
        qkv_dim = (c.num_attention_heads + 2 * c.num_key_value_heads) * hd
        self._x = torch.zeros((B, c.hidden_size), dtype=torch.float16, device=self.device)
        self._x2 = torch.zeros_like(self._x)
        self._nrm = torch.zeros_like(self._x)
        # split-K partials of the wide-batch path: [splits <= 8, rows, widest N]
        n_max = max(qkv_dim, c.hidden_size)
        self._part = torch.zeros((8, B, n_max), dtype=torch.float32, device=self.device)
        # SCB_SPLITK_FUSED=1: reduction + epilogue + RMSNorm inside the split-K GEMM launch (scb_linear_splitk) instead of
        # a reduce launch behind it — measured SLOWER (25 rows 5.40 -> 5.96 ms, 7 rows 4.69 -> 4.78 ms per step: the tail's
        # ticket / partial / row-counter round trips run while the weight stream keeps L2 busy, the separate reduce
        # kernel runs with programmatic dependent launch in a quiet memory system), so it stays opt-in
        self._tail_ws = (ops.splitk_tail_workspace(B, n_max, self.device)
                         if os.environ.get("SCB_SPLITK_FUSED", "0") == "1" else None)
        self._down_splits = next((s_ for s_ in (4, 8, 2, 1) if c.intermediate_size % (8 * s_) == 0), 1)
        self._qkv = torch.zeros((B, qkv_dim), dtype=torch.float16, device=self.device)
        self._attn = torch.zeros((B, c.num_attention_heads, hd), dtype=torch.float16, device=self.device)
        self._gu = torch.zeros((B, c.intermediate_size), dtype=torch.float16, device=self.device)
        self._logits = torch.zeros((B, c.vocab_size), dtype=torch.float16, device=self.device)
        self._attn_ws = ops.decode_attention_fused_workspace(self.device, c.num_attention_heads,
                                                             c.num_key_value_heads, hd, decode_splits, n_seqs=B)
        # persistent decode-step kernel (one launch per step) for narrow lock-step batches; shapes it does not
        # cover (and SCB_DECODE_MEGA=0) keep the per-layer kernel chain below — both are CUDA paths
        self._mega = None
        self.mega_unsupported: Optional[str] = None
        if os.environ.get("SCB_DECODE_MEGA", "1") != "0":
            try:
                self._mega = ops.Qwen2Decoder(
                    embed=self.embed, final_norm=self.final_norm, lm_head=self.lm_head, layers=self.layers,
                    k_cache=self.k_cache, v_cache=self.v_cache, n_q=c.num_attention_heads,
                    n_kv=c.num_key_value_heads, head_dim=hd, inter=c.intermediate_size, eps=c.rms_norm_eps,
                    theta=c.rope_theta, inv_freq=self.inv_freq, max_ctx=max_ctx)
            except L.ScbError as e:
                self.mega_unsupported = str(e)
        self._graphs = {}
        self.graph_replays = 0
        self.graph_launches = 0      # kernels launched through graph replays (launch_count() only sees eager calls)
        self.profile: Optional[list] = None   # bench.py: (start event, end event, steps, bytes) per decode loop

    # ---------------------------------------------------------------- weights / work accounting
    def weight_bytes(self) -> int:
        n = self.lm_head.numel() + self.final_norm.numel()
        for ly in self.layers:
            n += sum(t.numel() for t in ly.values())
        return 2 * n

    def kv_bytes_per_token(self, ctx: int) -> int:
        c = self.cfg
        return 2 * c.num_hidden_layers * c.num_key_value_heads * ctx * c.head_dim * 2

    def decode_bytes_per_token(self, ctx: int) -> int:
        return self.weight_bytes() + self.kv_bytes_per_token(ctx)

    def prefill_flops(self, T: int) -> float:
        c = self.cfg
        hd = c.head_dim
        per_tok = 2 * c.hidden_size * ((c.num_attention_heads + 2 * c.num_key_value_heads) * hd
                                       + c.hidden_size + 3 * c.intermediate_size)
        attn = 2.0 * T * T * c.num_attention_heads * hd   # causal: half of 4*T*T*H*hd
        return c.num_hidden_layers * (per_tok * T + attn)

    # ---------------------------------------------------------------- prefill
    @property
    def ctx_len(self) -> int:
        return self.ctx_lens[0]

    def reset(self, slot: Optional[int] = None) -> None:
        if slot is None:
            self.ctx_lens = [0] * self.max_batch
        else:
            self.ctx_lens[slot] = 0

    @torch.no_grad()
    def prefill(self, inputs_embeds: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """Append T tokens ([T, H] fp16) to the context of cache slot `slot`; returns the fp16 logits
        [1, V] of the last one (a view of row `slot` of the engine's logits buffer)."""
        c = self.cfg
        hd, Hq, Hkv = c.head_dim, c.num_attention_heads, c.num_key_value_heads
        x = inputs_embeds.to(device=self.device, dtype=torch.float16).contiguous()
        T = x.shape[0]
        pos0 = self.ctx_lens[slot]
        if pos0 + T > self.max_ctx:
            raise ValueError(f"context {pos0 + T} exceeds max_ctx {self.max_ctx}")
        if self._mega is not None and os.environ.get("SCB_NATIVE_PREFILL", "1") != "0":
            # ONE native call (scb_qwen2_prefill) issues the whole launch sequence below from C++ over the decoder
            # handle's pointer tables and a reused workspace (no per-op tensor allocations); SCB_NATIVE_PREFILL=0
            # keeps the per-launch Python loop (same kernels, bit-identical results — tests/test_gpu_fullwidth.py)
            self._mega.prefill(x, slot, pos0, self._logits[slot])
            self.ctx_lens[slot] = pos0 + T
            return self._logits[slot:slot + 1]
        scale = hd ** -0.5
        for i, ly in enumerate(self.layers):
            kc, vc = self.k_cache[i][slot], self.v_cache[i][slot]
            y = ops.rmsnorm(x, ly["ln1"], c.rms_norm_eps)
            qkv = ops.linear(y, ly["qkv_w"], bias=ly["qkv_b"])
            ops.rope_append(qkv, Hq, Hkv, hd, self.inv_freq, c.rope_theta, pos_base=pos0, k_cache=kc, v_cache=vc)
            q = qkv[:, :Hq * hd].view(1, T, Hq, hd)
            kv_len = pos0 + T
            k = kc[:, :kv_len].permute(1, 0, 2).unsqueeze(0)
            v = vc[:, :kv_len].permute(1, 0, 2).unsqueeze(0)
            ctx = ops.attention(q, k, v, scale, causal=True)
            x = ops.linear(ctx.view(T, Hq * hd), ly["o_w"], residual=x)
            y = ops.rmsnorm(x, ly["ln2"], c.rms_norm_eps)
            y = ops.linear(y, ly["gu_w"], act=L.ACT_SILU_PAIR)
            x = ops.linear(y, ly["down_w"], residual=x)
        self.ctx_lens[slot] = pos0 + T
        # final norm + lm_head on the last position only (logits_to_keep, modeling_qwen2.py:474-475)
        last = ops.rmsnorm(x[T - 1:T], self.final_norm, c.rms_norm_eps)
        return ops.gemv(last, self.lm_head, out=self._logits[slot:slot + 1])

    # ---------------------------------------------------------------- decode
    def _decode_step_launches(self, R: int = 1, ctx_cap: int = 0, ctx_sum: int = 0) -> None:
        """One cached decode step for the first R cache slots (5*layers + 2 launches): embedding
        gather (+ first RMSNorm), then per layer qkv GEMV -> fused RoPE/append/attention/combine ->
        o_proj GEMV (+residual, + post-attention RMSNorm by its last block) -> gate|up GEMV (SiLU-gate
        fused) -> down GEMV (+residual, + the next layer's input RMSNorm), then the lm_head GEMV with
        the greedy arg-max fused.  Every kernel prefetches the next one's weights into L2.  Leaves
        logits in self._logits[:R], the new tokens in next_id[:R] and advances the device-side
        [position, kv_len] state.  Rows are independent: a sequence decodes to the same ids whatever R is."""
        c = self.cfg
        hd, Hq, Hkv = c.head_dim, c.num_attention_heads, c.num_key_value_heads
        if self._use_mega(R):
            cap = ctx_cap if 0 < ctx_cap < self.max_ctx else self.max_ctx
            try:
                self._mega.step(R, self.next_id, self.state[1], self._logits, counters=self.state.view(-1), ctx_cap=cap,
                                ctx_sum=ctx_sum)
                return
            except L.ScbError as e:
                # e.g. a cooperative launch refused because the device is partitioned / shared, or a (rows, context)
                # combination whose activations leave no room for the weight ring: the per-layer chain (also CUDA)
                # takes over.  Launch errors surface before anything ran, so the step is simply redone there.
                if torch.cuda.is_current_stream_capturing():
                    raise
                self.mega_unsupported = str(e)
                self.mega_rows = R - 1
        scale = hd ** -0.5
        x, x2, nrm = self._x[:R], self._x2[:R], self._nrm[:R]
        eps = c.rms_norm_eps
        # KV splits: ~2 blocks per SM in ONE wave whatever the batch is (fewer partials to combine)
        G = Hq // Hkv
        cap = ctx_cap if 0 < ctx_cap < self.max_ctx else self.max_ctx
        smem_min = -(-cap * G * 4 // (48 * 1024))                # per-split score buffer: <= 48 KB per block
        splits = max(2, smem_min, min(self.decode_splits, 296 // (R * Hkv)))
        L0 = self.layers[0]
        ops.gather_rows_rmsnorm(self.embed, self.next_id[:R], L0["ln1"], eps, out=x, norm_out=nrm)
        n_layers = len(self.layers)
        wide = R >= self.wide_min_rows
        for i, ly in enumerate(self.layers):
            nxt_gamma = self.layers[i + 1]["ln1"] if i + 1 < n_layers else self.final_norm
            nxt_w = self.layers[i + 1]["qkv_w"] if i + 1 < n_layers else self.lm_head
            if wide:
                # wide batches: the A-fragment reads of the streaming GEMV (R x K per 8 weight rows) outgrow
                # L1, so every layer runs on the tcgen05 GEMM — split over K where N alone gives too few tiles
                ops.linear_splitk(nrm, ly["qkv_w"], 4, self._part, self._qkv[:R], bias=ly["qkv_b"], tail_ws=self._tail_ws)
            else:
                ops.gemv(nrm, ly["qkv_w"], bias=ly["qkv_b"], out=self._qkv[:R])
            ops.decode_attention_fused(self._qkv[:R], self.inv_freq, c.rope_theta, self.k_cache[i], self.v_cache[i],
                                       self.state[1, :R], Hq, scale, self._attn_ws, n_splits=splits,
                                       out=self._attn[:R], prefetch=None if wide else ly["o_w"], ctx_cap=cap)
            attn2d = self._attn[:R].view(R, Hq * hd)
            if wide:
                ops.linear_splitk(attn2d, ly["o_w"], 4, self._part, x2, residual=x, post_norm=(ly["ln2"], nrm, eps),
                                  tail_ws=self._tail_ws)
                ops.linear(nrm, ly["gu_w"], act=L.ACT_SILU_PAIR, out=self._gu[:R])
                ops.linear_splitk(self._gu[:R], ly["down_w"], self._down_splits, self._part, x, residual=x2,
                                  post_norm=(nxt_gamma, nrm, eps), tail_ws=self._tail_ws)
            else:
                ops.gemv(attn2d, ly["o_w"], residual=x, out=x2, post_norm=(ly["ln2"], nrm, eps), prefetch=ly["gu_w"])
                ops.gemv(nrm, ly["gu_w"], epilogue=L.EPI_SILU_PAIR, out=self._gu[:R], prefetch=ly["down_w"])
                ops.gemv(self._gu[:R], ly["down_w"], residual=x2, out=x, post_norm=(nxt_gamma, nrm, eps),
                         prefetch=nxt_w)
        if wide:
            ops.linear(nrm, self.lm_head, out=self._logits[:R])
            ops.argmax_rows(self._logits[:R], out=self.next_id[:R], counter=self.state.view(-1))
        else:
            ops.gemv(nrm, self.lm_head, out=self._logits[:R], argmax_out=self.next_id[:R],
                     counters=self.state.view(-1))

    def _use_mega(self, R: int) -> bool:
        return self._mega is not None and R <= min(self._mega.max_rows, self.mega_rows)

    def launches_per_decode_step(self, R: int = 1) -> int:
        if self._use_mega(R):
            return 1
        # narrow chain: gather+norm, 5 per layer, lm_head with fused arg-max; wide: split-K GEMM + reduce pairs
        wide = R >= self.wide_min_rows
        return 1 + self.cfg.num_hidden_layers * (8 if wide else 5) + (2 if wide else 1)

    def _check_context(self, prompt_lens, budgets) -> None:
        """A generation of `b` new tokens after a prompt of T tokens appends KV rows up to position T + b - 2 and
        attends over T + b - 1 of them: that has to fit the cache slot.  HF's DynamicCache grows without bound; here
        the slot is fixed (`max_context` of the model config), so the call fails BEFORE anything is launched rather
        than writing past the slot."""
        for T, b in zip(prompt_lens, budgets):
            if T + b - 1 > self.max_ctx:
                raise ValueError(f"prompt of {T} tokens + {b} new tokens needs a context of {T + b - 1} > max_ctx "
                                 f"{self.max_ctx}: raise LlavaQwenConfig.max_context (overwrite_config) or shorten the request")

    def _ctx_bucket(self, need: int) -> int:
        """Context capacity class of a decode graph (power of two, >= 2048, <= max_ctx)."""
        b = 2048
        while b < need and b < self.max_ctx:
            b *= 2
        return min(b, self.max_ctx)

    def _ensure_graph(self, R: int, ctx_cap: int = 0, ctx_sum: int = 0) -> "torch.cuda.CUDAGraph":
        """One graph per (rows, context-capacity class, bound of the rows' total length — the persistent kernel sizes
        its attention score buffer with it when a long context decodes next to short ones)."""
        ctx_cap = ctx_cap or self.max_ctx
        g = self._graphs.get((R, ctx_cap, ctx_sum))
        if g is not None:
            return g
        # warm-up on a side stream (loads modules, sets func attributes), then capture.  Both scribble on
        # cache rows 0.. of the first R slots, so graphs are built before any prefill of a generate call.
        saved_state, saved_id = self.state.clone(), self.next_id.clone()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.state[0].fill_(0)
            self.state[1].fill_(1)
            self._decode_step_launches(R, ctx_cap, ctx_sum)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.state[0].fill_(0)
        self.state[1].fill_(1)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._decode_step_launches(R, ctx_cap, ctx_sum)
        self._graphs[(R, ctx_cap, ctx_sum)] = g
        self.state.copy_(saved_state)
        self.next_id.copy_(saved_id)
        torch.cuda.synchronize(self.device)
        return g

    @torch.no_grad()
    def generate_batch(self, inputs_embeds_list: List[torch.Tensor], max_new_tokens, eos_token_id=None,
                       use_graph: bool = True) -> List[torch.Tensor]:
        """Greedy generation for up to `max_batch` independent prompts: each is prefilled into its own
        cache slot, then all decode in lock-step (one weight pass per step for all of them).
        `max_new_tokens` is an int or one int per prompt; sequences with a larger budget take the lower
        slots so that finished ones drop off the end and the step shrinks to the rows still running.
        Returns a list (in the caller's order) of [1, T_new] int64 tensors, each truncated after its
        first EOS.  Rows are independent: every sequence gets the ids it would get alone."""
        n = len(inputs_embeds_list)
        if n < 1 or n > self.max_batch:
            raise ValueError(f"{n} sequences but max_batch={self.max_batch}")
        budgets = [int(max_new_tokens)] * n if isinstance(max_new_tokens, int) else [int(b) for b in max_new_tokens]
        if len(budgets) != n or min(budgets) < 1:
            raise ValueError("max_new_tokens must be a positive int or one per sequence")
        self._check_context([int(x.shape[0]) for x in inputs_embeds_list], budgets)
        order = sorted(range(n), key=lambda i: -budgets[i])           # slot s holds caller's sequence order[s]
        cap = self._ctx_bucket(max(int(x.shape[0]) + b for x, b in zip(inputs_embeds_list, budgets)))
        # bound of the total context of the first R slots (prompt + budget each), rounded up to 1 K: graph key and the
        # persistent kernel's score-buffer size
        need = [min(cap, int(inputs_embeds_list[i].shape[0]) + budgets[i]) for i in order]

        def sum_bound(r: int) -> int:
            return -(-sum(need[:r]) // 1024) * 1024
        if os.environ.get("SCB_NO_GRAPH") == "1":      # profiling aid: every kernel as its own launch (ncu launch lists)
            use_graph = False
        with torch.cuda.device(self.device):
            if use_graph:
                for r in sorted({sum(1 for b in budgets if b > t) for t in [0] + budgets} - {0}):
                    self._ensure_graph(r, cap, sum_bound(r))            # before prefill (see _ensure_graph)
            self.reset()
            lens = []
            pf0 = pf1 = None
            if self.profile is not None:
                pf0, pf1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                pf0.record()
            if n > 1 and self._mega is not None and os.environ.get("SCB_NATIVE_PREFILL", "1") != "0" \
                    and os.environ.get("SCB_PREFILL_GROUP", "0") == "1":
                # opt-in (SCB_PREFILL_GROUP=1): prompts prefilled in GROUPS, their rows stacked, every linear layer once
                # over the stack (scb_qwen2_prefill_batch) — same arithmetic per row, fuller GEMM waves.  Measured on the
                # bench (same box, A/B/A): 2 189 / 2 205 ms of prefill per stream against 2 137 ms one prompt at a time —
                # the taller GEMMs push A out of L2 and the power-capped clock down, so the default stays one at a time
                xs = [inputs_embeds_list[i].to(device=self.device, dtype=torch.float16) for i in order]
                s0 = 0
                while s0 < n:
                    s1, tok = s0, 0
                    while s1 < n and (s1 == s0 or tok + xs[s1].shape[0] <= self.prefill_group_tokens):
                        tok += xs[s1].shape[0]
                        s1 += 1
                    for x_ in xs[s0:s1]:
                        if x_.shape[0] > self.max_ctx:
                            raise ValueError(f"context {x_.shape[0]} exceeds max_ctx {self.max_ctx}")
                    stacked = xs[s0] if s1 == s0 + 1 else torch.cat(xs[s0:s1], 0)
                    self._mega.prefill_batch(stacked.contiguous(), [int(x_.shape[0]) for x_ in xs[s0:s1]], list(range(s0, s1)),
                                             [0] * (s1 - s0), self._logits[s0:])
                    for s_ in range(s0, s1):
                        self.ctx_lens[s_] = int(xs[s_].shape[0])
                        lens.append(self.ctx_lens[s_])
                    s0 = s1
                ops.argmax_rows(self._logits[:n], out=self.next_id[:n])
            else:
                for s_, i in enumerate(order):
                    logits = self.prefill(inputs_embeds_list[i], slot=s_)
                    ops.argmax_rows(logits, out=self.next_id[s_:s_ + 1])
                    lens.append(self.ctx_lens[s_])
            if pf0 is not None:
                pf1.record()
                self.profile.append({"kind": "prefill", "events": (pf0, pf1), "tokens": sum(lens), "rows": n,
                                     "flops": sum(self.prefill_flops(t) for t in lens)})
            st = torch.zeros((2, self.max_batch), dtype=torch.int32)
            st[0, :n] = torch.tensor(lens, dtype=torch.int32)
            st[1, :n] = st[0, :n] + 1
            self.state.copy_(st, non_blocking=True)
            eos = set()
            if eos_token_id is not None:
                eos = set(eos_token_id) if isinstance(eos_token_id, (list, tuple, set)) else {int(eos_token_id)}
            slot_budget = [budgets[i] for i in order]
            total = slot_budget[0]
            out = torch.zeros((total, n), dtype=torch.int64, device=self.device)
            ev0 = ev1 = None
            if self.profile is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            n_steps, prof_bytes, prof_steps = 0, 0, 0
            for step in range(total):
                R = sum(1 for b in slot_budget if b > step)              # rows still running
                out[step, :R].copy_(self.next_id[:R], non_blocking=True)
                n_steps = step + 1
                if eos and (step % 16 == 15 or step == total - 1):
                    toks = out[:n_steps].t().tolist()
                    if all(any(t in eos for t in row[:slot_budget[s_]]) for s_, row in enumerate(toks)):
                        break
                if step == total - 1:
                    break
                Rn = sum(1 for b in slot_budget if b > step + 1)         # rows that need another token
                if use_graph:
                    self._graphs[(Rn, cap, sum_bound(Rn))].replay()
                    self.graph_replays += 1
                    self.graph_launches += self.launches_per_decode_step(Rn)
                else:
                    self._decode_step_launches(Rn, cap, sum_bound(Rn))
                prof_steps += 1
                prof_bytes += self.weight_bytes() + sum(self.kv_bytes_per_token(lens[s_] + step) for s_ in range(Rn))
            if ev0 is not None:
                ev1.record()
                self.profile.append({"kind": "decode", "events": (ev0, ev1), "steps": prof_steps, "rows": n,
                                     "bytes": prof_bytes})
            rows = out[:n_steps].t().contiguous()
            host = rows.tolist() if eos else None
            results: List[Optional[torch.Tensor]] = [None] * n
            for s_, i in enumerate(order):
                m = min(n_steps, slot_budget[s_])
                if eos:
                    for j, t in enumerate(host[s_][:m]):
                        if t in eos:
                            m = j + 1
                            break
                self.ctx_lens[s_] = lens[s_] + m - 1
                results[i] = rows[s_:s_ + 1, :m].clone()
            return results

    @torch.no_grad()
    def generate(self, inputs_embeds: torch.Tensor, max_new_tokens: int, do_sample: bool = False,
                 temperature: float = 1.0, top_k: Optional[int] = 50, top_p: Optional[float] = None,
                 eos_token_id=None, use_graph: bool = True) -> torch.Tensor:
        """Prefill + decode; returns the NEW token ids only, [1, T_new] int64 (HF generate with
        inputs_embeds).  Greedy (do_sample False or temperature 0) is the parity path."""
        if (not do_sample) or temperature == 0:
            return self.generate_batch([inputs_embeds], max_new_tokens, eos_token_id, use_graph)[0]
        # sampling path: same kernels, token drawn by torch.multinomial as HF does (hazard H2)
        self._check_context([int(inputs_embeds.shape[0])], [int(max_new_tokens)])
        with torch.cuda.device(self.device):
            self.reset()
            logits = self.prefill(inputs_embeds, slot=0)
            T = self.ctx_lens[0]
            st = torch.zeros((2, self.max_batch), dtype=torch.int32)
            st[0, 0], st[1, 0] = T, T + 1
            self.state.copy_(st)
            eos = set()
            if eos_token_id is not None:
                eos = set(eos_token_id) if isinstance(eos_token_id, (list, tuple, set)) else {int(eos_token_id)}
            toks = []
            lv = self.cfg.logit_vocab or self.cfg.vocab_size
            self.next_id[:1].copy_(self._sample(logits[:, :lv], temperature, top_k, top_p))
            for step in range(max_new_tokens):
                t = int(self.next_id[0])
                toks.append(t)
                if t in eos or step == max_new_tokens - 1:
                    break
                self._decode_step_launches(1)            # advances state, overwrites next_id with the arg-max
                self.next_id[:1].copy_(self._sample(self._logits[:1, :lv], temperature, top_k, top_p))
            self.ctx_lens[0] = T + len(toks) - 1
            return torch.tensor([toks], dtype=torch.int64, device=self.device)

    @staticmethod
    def _sample(logits: torch.Tensor, temperature: float, top_k: Optional[int], top_p: Optional[float]) -> torch.Tensor:
        """HF sampling warpers (temperature, top-k 50 default, optional top-p) on the [1,V] logits.
        Not a parity path (hazard H2): the reference's draws depend on the global torch RNG."""
        lg = logits.float() / max(temperature, 1e-6)
        if top_k:
            kth = torch.topk(lg, min(top_k, lg.shape[-1])).values[..., -1, None]
            lg = lg.masked_fill(lg < kth, float("-inf"))
        if top_p is not None and top_p < 1.0:
            sl, si = torch.sort(lg, descending=False)
            cp = sl.softmax(-1).cumsum(-1)
            rm = cp <= (1 - top_p)
            rm[..., -1:] = False
            lg = lg.masked_fill(rm.scatter(-1, si, rm), float("-inf"))
        return torch.multinomial(lg.softmax(-1), 1).view(1)
