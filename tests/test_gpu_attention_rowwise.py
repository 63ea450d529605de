"""GPU parity of the fused attention, row-wise glue and decode kernels (through the C-ABI)
against plain PyTorch fp32 references of the same ops."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sdpa_ref(q, k, v, scale, causal):
    # q [B,Tq,Hq,D], k/v [B,Tk,Hkv,D] -> [B,Tq,Hq,D] in fp64
    B, Tq, Hq, D = q.shape
    Hkv = k.shape[2]
    g = Hq // Hkv
    qd = q.double().permute(0, 2, 1, 3)
    kd = k.double().permute(0, 2, 1, 3).repeat_interleave(g, dim=1)
    vd = v.double().permute(0, 2, 1, 3).repeat_interleave(g, dim=1)
    s = qd @ kd.transpose(-1, -2) * scale
    if causal:
        Tk = k.shape[1]
        mask = torch.arange(Tk)[None, :] > (torch.arange(Tq)[:, None] + (Tk - Tq))
        s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vd).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,Tq,Tk,Hq,Hkv,D,causal", [
    (2, 577, 577, 16, 16, 64, False),     # CLIP ViT-L/14-336 self-attention
    (1, 128, 128, 1, 1, 64, False),
    (1, 100, 300, 2, 2, 64, False),       # ragged, several kv blocks
    (1, 512, 512, 4, 2, 128, True),       # GQA causal
    (1, 700, 700, 7, 1, 128, True),       # Qwen2 group size 7, ragged
    (1, 130, 500, 7, 1, 128, True),       # chunked prefill: causal with offset
    (1, 1, 257, 2, 1, 128, True),
])
def test_attention_matches_reference(cuda, B, Tq, Tk, Hq, Hkv, D, causal):
    from streamchat_b200 import ops
    torch.manual_seed(B * 1000 + Tq + Tk)
    q = torch.randn(B, Tq, Hq, D).half()
    k = torch.randn(B, Tk, Hkv, D).half()
    v = torch.randn(B, Tk, Hkv, D).half()
    scale = 1.0 / math.sqrt(D)
    ref = _sdpa_ref(q, k, v, scale, causal)
    out = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), scale, causal)
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < 4e-3, err


def test_attention_fused_qkv_views_and_cache_layout(cuda):
    """q/k/v as strided slices of a fused [T, 3*H*D] buffer (ViT) and k/v in the [Hkv, T, D] cache
    layout permuted to [1, T, Hkv, D] (Qwen2)."""
    from streamchat_b200 import ops
    torch.manual_seed(4)
    T, H, D = 300, 4, 64
    qkv = torch.randn(2, T, 3 * H * D).half().to(cuda)
    q = qkv[:, :, :H * D].view(2, T, H, D)
    k = qkv[:, :, H * D:2 * H * D].view(2, T, H, D)
    v = qkv[:, :, 2 * H * D:].view(2, T, H, D)
    out = ops.attention(q, k, v, 0.125, False)
    ref = _sdpa_ref(q.cpu(), k.cpu(), v.cpu(), 0.125, False)
    assert (out.double().cpu() - ref).abs().max().item() < 4e-3
    # cache layout with a max_pos larger than the live context and garbage (NaN) beyond it
    Hkv, maxp, Dk, ctx = 2, 640, 128, 333
    kc = torch.full((Hkv, maxp, Dk), float("nan")).half()
    vc = torch.full((Hkv, maxp, Dk), float("nan")).half()
    kc[:, :ctx] = torch.randn(Hkv, ctx, Dk).half()
    vc[:, :ctx] = torch.randn(Hkv, ctx, Dk).half()
    qq = torch.randn(1, ctx, 4, Dk).half()
    kcd, vcd = kc.to(cuda), vc.to(cuda)
    kview = kcd[:, :ctx].permute(1, 0, 2).unsqueeze(0)
    vview = vcd[:, :ctx].permute(1, 0, 2).unsqueeze(0)
    out = ops.attention(qq.to(cuda), kview, vview, 0.0884, True)
    ref = _sdpa_ref(qq, kc[:, :ctx].permute(1, 0, 2).unsqueeze(0), vc[:, :ctx].permute(1, 0, 2).unsqueeze(0),
                    0.0884, True)
    assert torch.isfinite(out).all()
    assert (out.double().cpu() - ref).abs().max().item() < 4e-3


def test_layernorm_rmsnorm_embed(cuda):
    from streamchat_b200 import ops
    torch.manual_seed(0)
    x = (torch.randn(300, 1024) * 2 + 0.5).half()
    g, b = torch.randn(1024).half(), torch.randn(1024).half()
    ref = torch.nn.functional.layer_norm(x.float(), (1024,), g.float(), b.float(), 1e-5)
    out = ops.layernorm(x.to(cuda), g.to(cuda), b.to(cuda), 1e-5).cpu().float()
    assert (out - ref).abs().max().item() < 1e-2 and (out - ref).abs().mean().item() < 1e-3
    # RMSNorm in HF order (modeling_qwen2.py:249-263)
    x2 = torch.randn(17, 3584).half()
    w = (1 + 0.1 * torch.randn(3584)).half()
    xf = x2.float()
    nrm = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).half()
    ref2 = (w * nrm).float()
    out2 = ops.rmsnorm(x2.to(cuda), w.to(cuda), 1e-6).cpu().float()
    assert (out2 - ref2).abs().max().item() < 4e-3
    # token assembly + pre-LN
    n, T, D = 3, 10, 64
    patch = torch.randn(n * (T - 1), D).half()
    cls, pos = torch.randn(D).half(), torch.randn(T, D).half()
    gg, bb = torch.randn(D).half(), torch.randn(D).half()
    emb = torch.cat([cls.expand(n, 1, D), patch.view(n, T - 1, D)], 1) + pos
    ref3 = torch.nn.functional.layer_norm(emb.float(), (D,), gg.float(), bb.float(), 1e-5).view(-1, D)
    out3 = ops.vit_embed_layernorm(patch.to(cuda), cls.to(cuda), pos.to(cuda), gg.to(cuda), bb.to(cuda),
                                   1e-5, n).cpu().float()
    assert (out3 - ref3).abs().max().item() < 1e-2


def test_im2col_pool_silu_gather_argmax(cuda):
    from streamchat_b200 import ops
    torch.manual_seed(1)
    pix = torch.randn(2, 3, 28, 42).half()
    cols = ops.im2col_patch(pix.to(cuda), 14, 592).cpu()
    ref = torch.nn.functional.unfold(pix.float(), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(cols[:, :588].float(), ref) and (cols[:, 588:] == 0).all()
    # pool: [n, 1+g*g, D] with CLS skipped
    n, g, D = 2, 6, 64
    x = torch.randn(n, 1 + g * g, D).half()
    ref = torch.nn.functional.avg_pool2d(x[:, 1:].float().view(n, g, g, D).permute(0, 3, 1, 2), 2)
    ref = ref.flatten(2).transpose(1, 2).reshape(-1, D)
    out = ops.avgpool2x2_tokens(x.view(-1, D).to(cuda), n, 1 + g * g, 1, g).cpu().float()
    assert (out - ref).abs().max().item() < 2e-3
    # silu * up
    gu = torch.randn(5, 2 * 256).half()
    ref = (torch.nn.functional.silu(gu[:, :256].float()).half().float() * gu[:, 256:].float())
    out = ops.silu_mul(gu.to(cuda)).cpu().float()
    assert (out - ref).abs().max().item() < 4e-3
    # gather
    table = torch.randn(50, 128).half()
    ids = torch.tensor([3, 49, 0, 3], dtype=torch.int64)
    assert torch.equal(ops.gather_rows(table.to(cuda), ids.to(cuda)).cpu(), table[ids])
    # argmax: fp16 logits with ties -> lowest index; counter increments
    lg = torch.randn(3, 152064).half()
    lg[0, 777] = 9.0
    lg[0, 90000] = 9.0
    lg[1, 152063] = 10.0
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    out = ops.argmax_rows(lg.to(cuda), counter=cnt).cpu()
    assert out.tolist() == torch.argmax(lg.float(), dim=1).tolist() and out[0] == 777
    assert int(cnt) == 1
    out32 = ops.argmax_rows(lg.float().to(cuda)).cpu()
    assert out32.tolist() == out.tolist()


def _rope_ref(x, pos, inv_freq):
    # HF apply_rotary_pos_emb in fp16 (modeling_qwen2.py:116-146)
    ang = pos[:, None].float() * inv_freq[None, :]
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos().half(), emb.sin().half()
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return (x * cos[:, None, :]) + (rot * sin[:, None, :])


def test_rope_append(cuda):
    from streamchat_b200 import ops
    torch.manual_seed(2)
    rows, nq, nkv, hd, maxp = 9, 4, 2, 128, 64
    qkv = torch.randn(rows, (nq + 2 * nkv) * hd).half()
    inv_freq = 1.0 / (1000000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    pos = torch.arange(rows) + 5
    q_ref = _rope_ref(qkv[:, :nq * hd].view(rows, nq, hd), pos, inv_freq)
    k_ref = _rope_ref(qkv[:, nq * hd:(nq + nkv) * hd].view(rows, nkv, hd), pos, inv_freq)
    v_ref = qkv[:, (nq + nkv) * hd:].view(rows, nkv, hd)
    d = qkv.to(cuda)
    kc = torch.zeros(nkv, maxp, hd, dtype=torch.float16, device=cuda)
    vc = torch.zeros_like(kc)
    ops.rope_append(d, nq, nkv, hd, inv_freq.to(cuda), 1e6, pos_base=5, k_cache=kc, v_cache=vc)
    got_q = d[:, :nq * hd].view(rows, nq, hd).cpu().float()
    got_k = d[:, nq * hd:(nq + nkv) * hd].view(rows, nkv, hd).cpu().float()
    assert (got_q - q_ref.float()).abs().max().item() < 4e-3
    assert (got_k - k_ref.float()).abs().max().item() < 4e-3
    assert torch.equal(kc[:, 5:5 + rows].permute(1, 0, 2).cpu().float(), got_k)
    assert torch.equal(vc[:, 5:5 + rows].permute(1, 0, 2).cpu(), v_ref)
    assert (kc[:, :5] == 0).all() and (kc[:, 5 + rows:] == 0).all()
    # device-side positions (graph-replayable decode step)
    d2 = qkv[:1].clone().to(cuda)
    p = torch.tensor([20], dtype=torch.int32, device=cuda)
    ops.rope_append(d2, nq, nkv, hd, inv_freq.to(cuda), 1e6, positions=p, k_cache=kc, v_cache=vc)
    k20 = _rope_ref(qkv[:1, nq * hd:(nq + nkv) * hd].view(1, nkv, hd), torch.tensor([20]), inv_freq)
    assert (kc[:, 20].cpu().float() - k20[0].float()).abs().max().item() < 4e-3


@pytest.mark.parametrize("rows,n,k,pro", [(1, 4608, 3584, 1), (1, 3584, 3584, 0), (1, 3584, 18944, 2),
                                          (2, 1000, 512, 0), (4, 300, 1024, 1), (1, 152064, 3584, 1),
                                          (5, 3584, 18944, 0), (8, 4608, 3584, 0), (3, 37888, 3584, 0),
                                          (13, 3584, 18944, 0), (16, 4608, 3584, 0), (9, 1000, 512, 0),
                                          (4, 3584, 18944, 2), (2, 4608, 3584, 1),
                                          (7, 1003, 96, 0), (1, 77, 32, 0), (2, 50, 40, 0)])
def test_gemv(cuda, rows, n, k, pro):
    from streamchat_b200 import ops
    torch.manual_seed(rows + n + k)
    w = (torch.randn(n, k) * 0.02).half()
    bias = torch.randn(n).half() if n < 10000 else None
    res = torch.randn(rows, n).half() if pro != 1 else None
    if pro == 2:
        a = torch.randn(rows, 2 * k).half()
        eff = (torch.nn.functional.silu(a[:, :k].float()).half().float() * a[:, k:].float()).half()
        gamma = None
    elif pro == 1:
        a = torch.randn(rows, k).half()
        gamma = (1 + 0.1 * torch.randn(k)).half()
        af = a.float()
        eff = (gamma * (af * torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + 1e-6)).half())
    else:
        a = torch.randn(rows, k).half()
        eff, gamma = a, None
    ref = eff.double() @ w.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if res is not None:
        ref = ref + res.double()
    out = ops.gemv(a.to(cuda), w.to(cuda), bias=None if bias is None else bias.to(cuda),
                   residual=None if res is None else res.to(cuda), prologue=pro,
                   gamma=None if gamma is None else gamma.to(cuda), eps=1e-6, out_dtype=torch.float32)
    err = (out.double().cpu() - ref).abs()
    assert (err <= 2e-3 * (ref.abs() + 1)).all(), err.max().item()


@pytest.mark.parametrize("hq,hkv,d,ctx,splits", [(28, 4, 128, 4700, 37), (28, 4, 128, 5, 37), (4, 4, 64, 300, 8)])
def test_decode_attention(cuda, hq, hkv, d, ctx, splits):
    from streamchat_b200 import ops
    torch.manual_seed(ctx)
    maxp = ctx + 100
    q = torch.randn(hq, d).half()
    kc = torch.randn(hkv, maxp, d).half()
    vc = torch.randn(hkv, maxp, d).half()
    scale = 1 / math.sqrt(d)
    ref = _sdpa_ref(q.view(1, 1, hq, d), kc[:, :ctx].permute(1, 0, 2).unsqueeze(0),
                    vc[:, :ctx].permute(1, 0, 2).unsqueeze(0), scale, False)[0, 0]
    out = ops.decode_attention(q.to(cuda), kc.to(cuda), vc.to(cuda), scale, kv_len=ctx, n_splits=splits)
    assert (out.double().cpu() - ref).abs().max().item() < 3e-3
    lens = torch.tensor([ctx], dtype=torch.int32, device=cuda)
    out2 = ops.decode_attention(q.to(cuda), kc.to(cuda), vc.to(cuda), scale, kv_len_dev=lens, n_splits=splits)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows", [1, 5, 16])
def test_gemv_fused_argmax(cuda, rows):
    from streamchat_b200 import ops, _lib
    torch.manual_seed(rows)
    n, k = 152064, 3584
    w = (torch.randn(n, k, device=cuda) * 0.02).half()
    a = torch.randn(rows, k, device=cuda).half()
    gamma = (1 + 0.1 * torch.randn(k, device=cuda)).half()
    ids = torch.full((rows,), -1, dtype=torch.int64, device=cuda)
    cnt = torch.tensor([10, 11], dtype=torch.int32, device=cuda)
    a = ops.rmsnorm(a, gamma, 1e-6)
    out = ops.gemv(a, w, argmax_out=ids, counters=cnt)
    assert out.dtype == torch.float16
    assert ids.tolist() == torch.argmax(out.float(), dim=1).tolist()      # over the stored fp16 logits
    assert cnt.tolist() == [11, 12]
    # ties -> lowest index: duplicate the winning weight row at a lower and a higher index
    best = int(ids[0])
    w2 = w.clone()
    lo, hi = 5, n - 3
    w2[lo] = w[best]
    w2[hi] = w[best]
    out2 = ops.gemv(a, w2, argmax_out=ids, counters=cnt)
    assert int(ids[0]) == min(lo, best) and ids.tolist() == torch.argmax(out2.float(), dim=1).tolist()
    assert cnt.tolist() == [12, 13]


@pytest.mark.parametrize("hq,hkv,d,ctx,splits,maxp", [(28, 4, 128, 4701, 37, 8192), (28, 4, 128, 1, 37, 8192),
                                                      (28, 4, 128, 38, 37, 32768), (4, 2, 64, 300, 8, 512),
                                                      (2, 1, 128, 77, 37, 2048)])
def test_decode_attention_fused(cuda, hq, hkv, d, ctx, splits, maxp):
    """ctx counts the token being decoded: its k/v come from the fused qkv row, not from the cache."""
    from streamchat_b200 import ops
    torch.manual_seed(ctx + hq)
    pos = ctx - 1
    qkv = torch.randn(1, (hq + 2 * hkv) * d).half()
    inv_freq = 1.0 / (1000000.0 ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    kc = torch.zeros(hkv, maxp, d).half()
    vc = torch.zeros(hkv, maxp, d).half()
    kc[:, :pos] = torch.randn(hkv, pos, d).half()
    vc[:, :pos] = torch.randn(hkv, pos, d).half()
    scale = 1 / math.sqrt(d)
    p = torch.tensor([pos])
    q_rot = _rope_ref(qkv[:, :hq * d].view(1, hq, d), p, inv_freq)
    k_new = _rope_ref(qkv[:, hq * d:(hq + hkv) * d].view(1, hkv, d), p, inv_freq)
    v_new = qkv[:, (hq + hkv) * d:].view(1, hkv, d)
    k_all = torch.cat([kc[:, :pos], k_new.permute(1, 0, 2)], 1)
    v_all = torch.cat([vc[:, :pos], v_new.permute(1, 0, 2)], 1)
    ref = _sdpa_ref(q_rot.view(1, 1, hq, d), k_all.permute(1, 0, 2).unsqueeze(0), v_all.permute(1, 0, 2).unsqueeze(0),
                    scale, False)[0, 0]
    kcd, vcd = kc.to(cuda), vc.to(cuda)
    ws = ops.decode_attention_fused_workspace(cuda, hq, hkv, d, splits)
    lens = torch.tensor([ctx], dtype=torch.int32, device=cuda)
    out = ops.decode_attention_fused(qkv.to(cuda), inv_freq.to(cuda), 1e6, kcd, vcd, lens, hq, scale, ws, splits)
    assert (out.double().cpu() - ref).abs().max().item() < 4e-3
    assert (kcd[:, pos].cpu().float() - k_new[0].float()).abs().max().item() < 4e-3
    assert torch.equal(vcd[:, pos].cpu(), v_new[0])
    assert (ws[:256] == 0).all()                                        # tickets left clean
    out2 = ops.decode_attention_fused(qkv.to(cuda), inv_freq.to(cuda), 1e6, kcd, vcd, lens, hq, scale, ws, splits)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows", [1, 5, 12])
def test_gemv_and_linear_silu_pair_epilogue(cuda, rows):
    """gate|up interleaved weights: out[r,i] = fp16(silu(fp16 gate_i)) * fp16 up_i — decode GEMV and the
    tcgen05 prefill GEMM must agree with each other bit for bit on the same rows and with the eager
    formulation (modeling_qwen2.py:41-48)."""
    from streamchat_b200 import ops, _lib
    torch.manual_seed(rows)
    k, inter = 512, 1536
    gate = (torch.randn(inter, k) * 0.05).half()
    up = (torch.randn(inter, k) * 0.05).half()
    a = torch.randn(rows, k).half()
    w = torch.stack([gate, up], 1).reshape(2 * inter, k).contiguous()
    g = (a.double() @ gate.double().t()).float().half().float()
    u = (a.double() @ up.double().t()).float().half().float()
    ref = torch.nn.functional.silu(g).half().float() * u
    out_v = ops.gemv(a.to(cuda), w.to(cuda), epilogue=_lib.EPI_SILU_PAIR)
    assert out_v.shape == (rows, inter)
    tol = 2e-3 * (ref.abs() + 1)          # fp16 output: half an ulp of the value
    assert ((out_v.float().cpu() - ref).abs() <= tol).all()
    out_m = ops.linear(a.to(cuda), w.to(cuda), act=_lib.ACT_SILU_PAIR)
    assert out_m.shape == (rows, inter)
    assert ((out_m.float().cpu() - ref).abs() <= tol).all()


def test_decode_attention_fused_multi_sequence(cuda):
    """Batched captions: S sequences with different lengths in one launch == S single launches."""
    from streamchat_b200 import ops
    torch.manual_seed(3)
    S, hq, hkv, d, maxp, splits = 5, 28, 4, 128, 1024, 37
    lens = [700, 1, 333, 1000, 64]
    qkv = torch.randn(S, (hq + 2 * hkv) * d, device=cuda).half()
    inv_freq = (1.0 / (1000000.0 ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))).to(cuda)
    kc = torch.randn(S, hkv, maxp, d, device=cuda).half()
    vc = torch.randn(S, hkv, maxp, d, device=cuda).half()
    kc1, vc1 = kc.clone(), vc.clone()
    scale = 1 / math.sqrt(d)
    ws = ops.decode_attention_fused_workspace(cuda, hq, hkv, d, splits, n_seqs=S)
    ld = torch.tensor(lens, dtype=torch.int32, device=cuda)
    out = ops.decode_attention_fused(qkv, inv_freq, 1e6, kc, vc, ld, hq, scale, ws, splits)
    ws1 = ops.decode_attention_fused_workspace(cuda, hq, hkv, d, splits)
    for s_ in range(S):
        o1 = ops.decode_attention_fused(qkv[s_:s_ + 1], inv_freq, 1e6, kc1[s_], vc1[s_], ld[s_:s_ + 1], hq, scale,
                                        ws1, splits)
        assert torch.equal(out[s_], o1[0]), s_
    assert torch.equal(kc, kc1) and torch.equal(vc, vc1)


def test_gemv_post_norm_prefetch_and_gather_norm(cuda):
    """Decode-chain fusions: RMSNorm of the finished GEMV output by the last block, L2 prefetch of the
    next weights (no functional effect), embedding gather fused with the first norm."""
    from streamchat_b200 import ops
    torch.manual_seed(8)
    for rows, n, k in ((1, 3584, 3584), (5, 3584, 18944), (3, 256, 512)):
        a = torch.randn(rows, k, device=cuda).half()
        w = (torch.randn(n, k, device=cuda) * 0.02).half()
        res = torch.randn(rows, n, device=cuda).half()
        gamma = (1 + 0.1 * torch.randn(n, device=cuda)).half()
        nxt = torch.randn(1000, 64, device=cuda).half()
        nrm = torch.zeros(rows, n, device=cuda, dtype=torch.float16)
        out = ops.gemv(a, w, residual=res, post_norm=(gamma, nrm, 1e-6), prefetch=nxt)
        plain = ops.gemv(a, w, residual=res)
        assert torch.equal(out, plain)
        assert torch.equal(nrm, ops.rmsnorm(out, gamma, 1e-6))
    table = torch.randn(50, 256, device=cuda).half()
    ids = torch.tensor([3, 49, 0], dtype=torch.int64, device=cuda)
    gamma = (1 + 0.1 * torch.randn(256, device=cuda)).half()
    x = torch.zeros(3, 256, device=cuda, dtype=torch.float16)
    nrm = torch.zeros_like(x)
    ops.gather_rows_rmsnorm(table, ids, gamma, 1e-6, out=x, norm_out=nrm)
    assert torch.equal(x, table[ids]) and torch.equal(nrm, ops.rmsnorm(x, gamma, 1e-6))


@pytest.mark.parametrize("rows,n,k,splits", [(13, 3584, 18944, 4), (9, 4608, 3584, 4), (16, 3584, 3584, 4), (3, 256, 512, 2)])
def test_linear_splitk_matches_gemv(cuda, rows, n, k, splits):
    """Wide decode batches: K-sliced tcgen05 GEMM + ordered reduce == the streaming GEMV path (same
    fp16 inputs, fp32 accumulation) within fp16 output rounding, epilogues included."""
    from streamchat_b200 import ops
    torch.manual_seed(rows + n)
    a = torch.randn(rows, k, device=cuda).half()
    w = (torch.randn(n, k, device=cuda) * 0.02).half()
    res = torch.randn(rows, n, device=cuda).half()
    bias = torch.randn(n, device=cuda).half()
    gamma = (1 + 0.1 * torch.randn(n, device=cuda)).half()
    part = torch.zeros(8, 16, max(n, 4608), device=cuda, dtype=torch.float32)
    out = torch.zeros(rows, n, device=cuda, dtype=torch.float16)
    nrm = torch.zeros_like(out)
    ops.linear_splitk(a, w, splits, part, out, bias=bias, residual=res, post_norm=(gamma, nrm, 1e-6))
    ref = (a.double() @ w.double().t() + bias.double() + res.double()).cpu()
    assert ((out.double().cpu() - ref).abs() <= 2e-3 * (ref.abs() + 1)).all()
    # same rounding points as the standalone norm; the sum of squares is ordered differently (1 fp16 ulp)
    ref_n = ops.rmsnorm(out, gamma, 1e-6).float()
    assert ((nrm.float() - ref_n).abs() <= 2e-3 * (ref_n.abs() + 1e-3)).all()
    ref_v = ops.gemv(a, w, bias=bias, residual=res)
    assert (out.float() - ref_v.float()).abs().max().item() <= 2e-3 * (ref.abs().max().item() + 1)


@pytest.mark.parametrize("rows,n,k,splits,with_norm", [(25, 3584, 18944, 4, True), (25, 4608, 3584, 4, False),
                                                       (7, 3584, 3584, 4, True), (3, 256, 512, 2, True),
                                                       (128, 3584, 3584, 4, True), (25, 200, 512, 2, True)])
def test_linear_splitk_fused_tail(cuda, rows, n, k, splits, with_norm):
    """scb_linear_splitk (reduction + epilogue + RMSNorm inside the GEMM launch) against the two-launch path: the
    finished rows are BIT-identical (same slices summed in the same order), the normalised rows agree to the ordering
    of the sum of squares (1 fp16 ulp); the ticket workspace comes back zero, so repeated calls on one buffer — the
    qkv / o / down chain of a decode step, different N each — keep working."""
    from streamchat_b200 import ops
    torch.manual_seed(rows + n + 1)
    a = torch.randn(rows, k, device=cuda).half()
    w = (torch.randn(n, k, device=cuda) * 0.02).half()
    res = torch.randn(rows, n, device=cuda).half()
    bias = torch.randn(n, device=cuda).half()
    gamma = (1 + 0.1 * torch.randn(n, device=cuda)).half()
    part = torch.zeros(8, 128, max(n, 4608), device=cuda, dtype=torch.float32)
    ws = ops.splitk_tail_workspace(128, 4608, cuda)
    assert ops.linear_splitk_fits_one_launch(rows, n, splits, cuda)
    pn = lambda t: (gamma, t, 1e-6) if with_norm else None
    out0, nrm0 = torch.zeros(rows, n, device=cuda, dtype=torch.float16), torch.zeros(rows, n, device=cuda, dtype=torch.float16)
    ops.linear_splitk(a, w, splits, part, out0, bias=bias, residual=res, post_norm=pn(nrm0))
    for rep in range(3):
        out1, nrm1 = torch.full_like(out0, 7.0), torch.full_like(out0, 7.0)
        part.fill_(float("nan"))
        ops.linear_splitk(a, w, splits, part, out1, bias=bias, residual=res, post_norm=pn(nrm1), tail_ws=ws)
        assert torch.equal(out1, out0), f"rep {rep}"
        if with_norm:
            assert torch.equal(nrm1, ops.rmsnorm(out1, gamma, 1e-6)) or \
                ((nrm1.float() - nrm0.float()).abs() <= 2e-3 * (nrm0.float().abs() + 1e-3)).all()
        assert int(ws[:768].abs().sum().item()) == 0, "tickets must be left zero"
    # no residual / no bias
    out2 = torch.zeros_like(out0)
    ops.linear_splitk(a, w, splits, part, out2, tail_ws=ws)
    ref = (a.double() @ w.double().t()).cpu()
    assert ((out2.double().cpu() - ref).abs() <= 2e-3 * (ref.abs() + 1)).all()


def test_linear_splitk_too_many_tiles_takes_two_launches(cuda):
    """More (n tile, K slice) pairs than SMs: the wrapper runs GEMM + reduce (the fused tail needs every CTA resident)."""
    from streamchat_b200 import ops
    torch.manual_seed(5)
    rows, n, k, splits = 4, 4608, 1024, 8
    assert not ops.linear_splitk_fits_one_launch(rows, n, splits, cuda)
    a = torch.randn(rows, k, device=cuda).half()
    w = (torch.randn(n, k, device=cuda) * 0.02).half()
    part = torch.zeros(8, 16, n, device=cuda, dtype=torch.float32)
    ws = ops.splitk_tail_workspace(16, n, cuda)
    out = torch.zeros(rows, n, device=cuda, dtype=torch.float16)
    ops.linear_splitk(a, w, splits, part, out, tail_ws=ws)
    ref = (a.double() @ w.double().t()).cpu()
    assert ((out.double().cpu() - ref).abs() <= 2e-3 * (ref.abs() + 1)).all()


def test_linear_splitk_fused_shapes_share_one_workspace(cuda):
    """qkv (N = 4608) -> o (N = 3584, post-norm) -> down (K = 18944, post-norm) on ONE ticket buffer, twice over, as
    in a decode step: the sums of squares one shape leaves behind must not be read as another shape's tickets."""
    from streamchat_b200 import ops
    torch.manual_seed(11)
    rows = 25
    shapes = [(4608, 3584, False), (3584, 3584, True), (3584, 18944, True)]
    ws = ops.splitk_tail_workspace(32, 4608, cuda)
    part = torch.zeros(8, 32, 4608, device=cuda, dtype=torch.float32)
    data = []
    for n, k, norm in shapes:
        a = torch.randn(rows, k, device=cuda).half()
        w = (torch.randn(n, k, device=cuda) * 0.02).half()
        gamma = (1 + 0.1 * torch.randn(n, device=cuda)).half()
        ref, ref_n = torch.zeros(rows, n, device=cuda, dtype=torch.float16), torch.zeros(rows, n, device=cuda, dtype=torch.float16)
        ops.linear_splitk(a, w, 4, part, ref, post_norm=(gamma, ref_n, 1e-6) if norm else None)
        data.append((a, w, gamma, norm, ref, ref_n))
    for rep in range(2):
        for a, w, gamma, norm, ref, ref_n in data:
            out, nrm = torch.zeros_like(ref), torch.zeros_like(ref)
            ops.linear_splitk(a, w, 4, part, out, post_norm=(gamma, nrm, 1e-6) if norm else None, tail_ws=ws)
            assert torch.equal(out, ref), (rep, tuple(w.shape))
            if norm:
                assert ((nrm.float() - ref_n.float()).abs() <= 2e-3 * (ref_n.float().abs() + 1e-3)).all()
