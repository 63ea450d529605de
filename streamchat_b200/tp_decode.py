"""Tensor-parallel greedy decode of ONE sequence over a small group of ranks (the answer's serial tail of the
frame-sharded stream).

The strong-scaling limiter of the sharded stream is the answer: 255 one-row decode steps that stream all 14.1 GB of
weights on rank 0 while the other ranks idle.  Here T ranks (T divides the KV-head count: 2 or 4 for Qwen2-7B) split
every step Megatron-style on VIEWS of the weights each rank already holds:

    qkv      column-parallel   rank r: its q heads + its KV heads (a small fused per-rank copy: 1 152 x 3 584 at T = 4)
    attention                  rank r: its KV heads' caches (every rank of the group ran the same prefill, so the
                               caches are resident everywhere: no KV transfer)
    o_proj   row-parallel      o_w[:, r-th slice]      -> partial [H] fp32  -> exchange -> ordered sum + residual + RMSNorm
    gate|up  column-parallel   rows [2*I/T*r, 2*I/T*(r+1)) of the row-interleaved matrix (SiLU gate fused)
    down     row-parallel      down_w[:, r-th slice]   -> partial [H] fp32  -> exchange -> ordered sum + residual + RMSNorm
    lm_head  column-parallel   V/T logits per rank     -> exchange -> full row, arg-max on every rank

Each exchange is `peer_allgather_kernel` (scb_allgather_embed with the call number kept on the device) on a 14 KB slice
of a library-allocated, IPC-mapped buffer, followed by the existing ordered `scb_splitk_reduce` — partial sums are added
in rank order on every rank, so all ranks compute bit-identical activations and tokens.  One step = 9 launches per layer,
captured once in a CUDA graph per context-capacity class.  HBM bytes per rank per step: 1/T of the weights + 1/T of the
KV cache.  The reference has no counterpart (it decodes on one GPU: llava_qwen.py:155).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from . import ops


class TensorParallelDecoder:
    """Construction is a collective over `group` (every rank of the group calls it with its own engine; the engines
    hold identical weights).  `generate` must be called by every rank of the group with identical inputs."""

    def __init__(self, engine, tp_rank: int, tp_size: int, group=None):
        c = engine.cfg
        T, r = tp_size, tp_rank
        Hq, Hkv, hd, H, I, V = (c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.hidden_size,
                                c.intermediate_size, engine.lm_head.shape[0])
        if T < 2 or Hkv % T or Hq % T or I % T or V % T or (I // T) % 32 or (Hq // T * hd) % 32 or (V // T) % 8:
            raise L.ScbError(f"tensor-parallel decode: {T} ranks do not divide heads {Hq}/{Hkv}, inter {I}, vocab {V}")
        self.engine, self.rank, self.T, self.group = engine, r, T, group
        self.device = engine.device
        self.qh, self.kvh, self.Ir, self.Vr = Hq // T, Hkv // T, I // T, V // T
        qh, kvh, Ir = self.qh, self.kvh, self.Ir
        self.layers = []
        for ly in engine.layers:
            w, b = ly["qkv_w"], ly["qkv_b"]
            q0, k0, v0 = 0, Hq * hd, (Hq + Hkv) * hd
            rows = torch.cat([torch.arange(q0 + qh * hd * r, q0 + qh * hd * (r + 1)),
                              torch.arange(k0 + kvh * hd * r, k0 + kvh * hd * (r + 1)),
                              torch.arange(v0 + kvh * hd * r, v0 + kvh * hd * (r + 1))]).to(self.device)
            self.layers.append(dict(
                qkv_w=w[rows].contiguous(), qkv_b=b[rows].contiguous(),
                o_w=ly["o_w"][:, qh * hd * r: qh * hd * (r + 1)],                # view, row stride H
                gu_w=ly["gu_w"][2 * Ir * r: 2 * Ir * (r + 1)],                   # view (gate_i / up_i rows stay adjacent)
                down_w=ly["down_w"][:, Ir * r: Ir * (r + 1)],                    # view, row stride I
                ln1=ly["ln1"], ln2=ly["ln2"]))
        self.lm_w = engine.lm_head[self.Vr * r: self.Vr * (r + 1)]
        # exchange buffers (library-allocated, mapped by every peer); o_proj and down_proj alternate between two so that
        # a fast rank's next push cannot land in a buffer a slow rank is still summing
        self.px_o = ops.PeerGather((T, 1, H), torch.float32, r, T, self.device, group, device_epoch=True)
        self.px_d = ops.PeerGather((T, 1, H), torch.float32, r, T, self.device, group, device_epoch=True)
        self.px_l = ops.PeerGather((T, self.Vr), torch.float16, r, T, self.device, group, device_epoch=True)
        dev = self.device
        self._x = torch.zeros((1, H), dtype=torch.float16, device=dev)
        self._x2 = torch.zeros_like(self._x)
        self._nrm = torch.zeros_like(self._x)
        self._qkv = torch.zeros((1, (qh + 2 * kvh) * hd), dtype=torch.float16, device=dev)
        self._attn = torch.zeros((1, qh, hd), dtype=torch.float16, device=dev)
        self._gu = torch.zeros((1, Ir), dtype=torch.float16, device=dev)
        self.splits = min(64, max(2, (2 * 148) // kvh))
        self._attn_ws = ops.decode_attention_fused_workspace(dev, qh, kvh, hd, self.splits, n_seqs=1)
        self._graphs = {}
        self.launches_per_step = 1 + 9 * len(self.layers) + 3

    # ------------------------------------------------------------------------------------------------------
    def _step(self, ctx_cap: int) -> None:
        """One decode step for cache slot 0: consumes engine.next_id[0], leaves the greedy next token there and
        advances the device-side (position, kv_len) state — the launch sequence every rank of the group issues."""
        e = self.engine
        c = e.cfg
        hd = c.head_dim
        eps = c.rms_norm_eps
        scale = hd ** -0.5
        r, kvh = self.rank, self.kvh
        x, x2, nrm = self._x, self._x2, self._nrm
        ops.gather_rows_rmsnorm(e.embed, e.next_id[:1], self.layers[0]["ln1"], eps, out=x, norm_out=nrm)
        n_layers = len(self.layers)
        for i, ly in enumerate(self.layers):
            nxt_gamma = self.layers[i + 1]["ln1"] if i + 1 < n_layers else e.final_norm
            ops.gemv(nrm, ly["qkv_w"], bias=ly["qkv_b"], out=self._qkv)
            kc = e.k_cache[i][0, kvh * r: kvh * (r + 1)]
            vc = e.v_cache[i][0, kvh * r: kvh * (r + 1)]
            ops.decode_attention_fused(self._qkv, e.inv_freq, c.rope_theta, kc, vc, e.state[1, :1], self.qh, scale,
                                       self._attn_ws, n_splits=self.splits, out=self._attn, ctx_cap=ctx_cap)
            ops.gemv(self._attn.view(1, -1), ly["o_w"], out=self.px_o.buffer[r], out_dtype=torch.float32)
            self.px_o.allgather()
            ops.splitk_reduce(self.px_o.buffer, x2, residual=x, post_norm=(ly["ln2"], nrm, eps))
            ops.gemv(nrm, ly["gu_w"], epilogue=L.EPI_SILU_PAIR, out=self._gu)
            ops.gemv(self._gu, ly["down_w"], out=self.px_d.buffer[r], out_dtype=torch.float32)
            self.px_d.allgather()
            ops.splitk_reduce(self.px_d.buffer, x, residual=x2, post_norm=(nxt_gamma, nrm, eps))
        ops.gemv(nrm, self.lm_w, out=self.px_l.buffer[r: r + 1])
        self.px_l.allgather()
        ops.argmax_rows(self.px_l.buffer.view(1, -1), out=e.next_id[:1], counter=e.state.view(-1))

    def _graph(self, ctx_cap: int) -> "torch.cuda.CUDAGraph":
        g = self._graphs.get(ctx_cap)
        if g is not None:
            return g
        e = self.engine
        saved_state, saved_id = e.state.clone(), e.next_id.clone()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):                      # warm-up (every rank: the exchanges are collective)
            e.state[0].fill_(0)
            e.state[1].fill_(1)
            self._step(ctx_cap)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        e.state[0].fill_(0)
        e.state[1].fill_(1)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step(ctx_cap)
        self._graphs[ctx_cap] = g
        e.state.copy_(saved_state)
        e.next_id.copy_(saved_id)
        torch.cuda.synchronize(self.device)
        return g

    @torch.no_grad()
    def generate(self, inputs_embeds: torch.Tensor, max_new_tokens: int, eos_token_id=None) -> torch.Tensor:
        """Prefill (every rank, identically) + tensor-parallel cached greedy decode.  Returns the NEW ids [1, T_new],
        identical on every rank of the group."""
        e = self.engine
        T_in = int(inputs_embeds.shape[0])
        e._check_context([T_in], [int(max_new_tokens)])
        cap = e._ctx_bucket(T_in + int(max_new_tokens))
        with torch.cuda.device(self.device):
            g = self._graph(cap)                        # before the prefill: capture scribbles on cache row 0
            e.reset()
            logits = e.prefill(inputs_embeds, slot=0)
            ops.argmax_rows(logits, out=e.next_id[:1])
            st = torch.zeros((2, e.max_batch), dtype=torch.int32)
            st[0, 0], st[1, 0] = T_in, T_in + 1
            e.state.copy_(st, non_blocking=True)
            eos = set()
            if eos_token_id is not None:
                eos = set(eos_token_id) if isinstance(eos_token_id, (list, tuple, set)) else {int(eos_token_id)}
            out = torch.zeros((max_new_tokens,), dtype=torch.int64, device=self.device)
            ev0 = ev1 = None
            if e.profile is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            n_steps, prof_bytes = 0, 0
            for step in range(max_new_tokens):
                out[step: step + 1].copy_(e.next_id[:1], non_blocking=True)
                n_steps = step + 1
                if eos and (step % 16 == 15 or step == max_new_tokens - 1):
                    if any(t in eos for t in out[:n_steps].tolist()):
                        break
                if step == max_new_tokens - 1:
                    break
                g.replay()
                e.graph_replays += 1
                e.graph_launches += self.launches_per_step
                prof_bytes += (e.weight_bytes() + e.kv_bytes_per_token(T_in + step)) // self.T
            if ev0 is not None:
                ev1.record()
                e.profile.append({"kind": "decode", "events": (ev0, ev1), "steps": n_steps - 1, "rows": 1,
                                  "bytes": prof_bytes, "tp": self.T})
            ids = out[:n_steps]
            if eos:
                host = ids.tolist()
                for j, t in enumerate(host):
                    if t in eos:
                        ids = ids[: j + 1]
                        break
            e.ctx_lens[0] = T_in + int(ids.numel()) - 1
            return ids.view(1, -1).clone()

    def close(self) -> None:
        for p in (self.px_o, self.px_d, self.px_l):
            p.close()
