"""GPU parity at the BASELINE widths (CLIP ViT-L/14-336 = hidden 1024 / 16 heads / MLP 4096 / 577
tokens; Qwen2-7B = hidden 3584 / 28 q + 4 kv heads / MLP 18944 / vocab 152064), depth cut to 2 layers so the CPU
oracle finishes in seconds.  These are the shapes the bench runs: the K-split of o_proj / down_proj and the
152 064-row arg-max of the persistent decode kernel, the 25-row split-K chain, the CTA-pair GEMM tiles and the
D=64 / D=128 attention kernels all see their production dimensions here.

fp tolerance: tests/parity.py (1e-3 relative L2 where fp16 execution itself holds 1e-3, else 1.5 x the error of
the oracle's fp16-rounded execution on the same inputs).  Greedy ids: exact, a divergence is tolerated only at a
step where the fp32 oracle's own top-2 logit margin is below fp16 resolution of the logit scale."""
import pytest
import torch

from oracle import longva_model as om
from parity import elem_rel, record, rel_l2, rel_max, tolerance

pytestmark = pytest.mark.gpu


def _check_ids(got, want, ref_logits, what):
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            top2 = torch.topk(ref_logits[i], 2).values
            margin = float(top2[0] - top2[1])
            assert margin < 4e-3 * float(ref_logits[i].abs().max()), (what, i, got, want, margin)
            return i
    return None


def test_vit_l14_336_width_vs_oracle(cuda):
    """2 frames through patch-embed + 2 full-width CLIP layers + projector + 2x2 pool (encode_images) vs the oracle."""
    from streamchat_b200.longva.model.builder import random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig
    from streamchat_b200.qwen2 import Qwen2Config
    from streamchat_b200.vision import ClipVisionConfig, VisionEncoder
    vcfg = ClipVisionConfig(num_hidden_layers=3)                 # select_layer -2 -> layers 0 and 1 feed the output
    cfg = LlavaQwenConfig(llm=Qwen2Config(num_hidden_layers=0, vocab_size=16), vision=vcfg)
    W = random_weights(cfg, cuda, seed=21, std=0.02)
    enc = VisionEncoder(W, vcfg, cuda)
    torch.manual_seed(0)
    images = torch.randn(2, 3, 336, 336).half()
    Wc = {k: v.float().cpu() for k, v in W.items()}
    o32 = om.encode_images(Wc, images.float(), n_heads=16)
    o16 = om.encode_images(Wc, images.float(), n_heads=16, act_dtype=torch.float16)
    t32 = om.clip_tower(Wc, images.float(), n_heads=16)
    t16 = om.clip_tower(Wc, images.float(), n_heads=16, act_dtype=torch.float16)
    tower = enc.tower(images.to(cuda)).view(2, 577, 1024)[:, 1:].float().cpu()
    feats = enc.encode_images(images.to(cuda))
    assert feats.shape == (2, 144, 3584) and feats.dtype == torch.float16
    feats = feats.float().cpu()
    for name, got, a32, a16 in (("vit_l14_tower", tower, t32, t16), ("vit_l14_encode_images", feats, o32, o16)):
        tol = tolerance(a16, a32)
        record(name, rel_l2=rel_l2(got, a32), rel_max=rel_max(got, a32), elem_rel_p999=elem_rel(got, a32),
               fp16_exec_rel_l2=rel_l2(a16, a32), fp16_exec_rel_max=rel_max(a16, a32), tol=tol)
        assert rel_l2(got, a32) < tol, (name, rel_l2(got, a32), tol)
        assert rel_max(got, a32) < max(2e-3, 2.0 * rel_max(a16, a32)), (name, rel_max(got, a32))


@pytest.fixture(scope="module")
def qwen7b_2layers(cuda):
    from streamchat_b200.longva.model.builder import random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig
    from streamchat_b200.qwen2 import Qwen2Config, Qwen2Engine
    from streamchat_b200.vision import ClipVisionConfig
    lcfg = Qwen2Config(num_hidden_layers=2)                       # 3584 / 18944 / 28q / 4kv / 152064
    full = LlavaQwenConfig(llm=lcfg, vision=ClipVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1,
                                                             num_attention_heads=2, image_size=56, patch_size=14))
    W = {k: v for k, v in random_weights(full, cuda, seed=31, std=0.02).items()
         if not k.startswith("model.vision_tower") and "mm_projector" not in k}
    eng = Qwen2Engine(W, lcfg, cuda, max_ctx=1024, max_batch=25)
    assert eng._mega is not None, eng.mega_unsupported
    Wc = {k: v.float().cpu() for k, v in W.items()}
    return lcfg, W, Wc, eng


def _prompt(Wc, n_text, n_feat, seed):
    """Rows like a real prompt: embedding-table rows (std 0.02) around a block of visual features (std 0.5)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, Wc["model.embed_tokens.weight"].shape[0], (n_text,), generator=g)
    text = Wc["model.embed_tokens.weight"][ids]
    feat = torch.randn(n_feat, text.shape[1], generator=g) * 0.5
    return torch.cat([text[: n_text // 2], feat, text[n_text // 2:]], 0).half()


def test_qwen2_7b_width_prefill_and_decode_mega_vs_oracle(cuda, qwen7b_2layers):
    """512-token prefill (CTA-pair GEMM tiles, D=128 causal attention over the cache) + 8 greedy tokens through the
    persistent decode kernel at full width (1 row, then 2 rows in lock-step) vs oracle.greedy_generate."""
    lcfg, W, Wc, eng = qwen7b_2layers
    x1, x2 = _prompt(Wc, 64, 448, 1), _prompt(Wc, 40, 260, 2)
    kw = dict(n_heads=28, n_kv=4, theta=lcfg.rope_theta, eps=lcfg.rms_norm_eps)
    ids1, lg1 = om.greedy_generate(Wc, x1.float(), 8, return_logits=True, **kw)
    ids2, lg2 = om.greedy_generate(Wc, x2.float(), 8, return_logits=True, **kw)
    h16 = om.qwen2_forward(Wc, x1.float(), 28, 4, lcfg.rope_theta, lcfg.rms_norm_eps, [], 0, torch.float16)
    l16 = (h16[-1] @ Wc["lm_head.weight"].t()).half().float()
    # prefill logits of the last position (152 064-row lm_head)
    eng.reset()
    first = eng.prefill(x1.to(cuda)).float().cpu()[0]
    eng.reset()
    tol = tolerance(l16, lg1[0])
    record("qwen2_7b_width_prefill_logits", rel_l2=rel_l2(first, lg1[0]), rel_max=rel_max(first, lg1[0]),
           elem_rel_p999=elem_rel(first, lg1[0]), fp16_exec_rel_l2=rel_l2(l16, lg1[0]), tol=tol)
    assert rel_l2(first, lg1[0]) < tol, (rel_l2(first, lg1[0]), tol)
    assert int(first.argmax()) == int(ids1[0, 0]) or \
        float(torch.topk(lg1[0], 2).values.diff().abs()) < 4e-3 * float(lg1[0].abs().max())
    # one row through decode_mega_kernel
    assert eng._use_mega(1) and eng._use_mega(2)
    a = eng.generate(x1.to(cuda), 8)
    div = _check_ids(a[0].tolist(), ids1[0].tolist(), lg1, "mega 1 row")
    if div is None:
        last = eng._logits[0].float().cpu()
        record("qwen2_7b_width_mega_step8_logits", rel_l2=rel_l2(last, lg1[-1]), rel_max=rel_max(last, lg1[-1]))
        assert rel_l2(last, lg1[-1]) < 2.0 * tol, rel_l2(last, lg1[-1])
    # graph replay == eager launches
    assert torch.equal(eng.generate(x1.to(cuda), 8, use_graph=False), a)
    # two rows in lock-step
    b = eng.generate_batch([x1.to(cuda), x2.to(cuda)], 8)
    assert torch.equal(b[0], a)
    _check_ids(b[1][0].tolist(), ids2[0].tolist(), lg2, "mega 2 rows")


def test_qwen2_7b_width_gemv_chain_vs_oracle(cuda, qwen7b_2layers):
    """The per-layer GEMV chain (the 5..8-row path) at full width gives the persistent kernel's ids."""
    lcfg, W, Wc, eng = qwen7b_2layers
    x1 = _prompt(Wc, 64, 448, 1)
    ids1, lg1 = om.greedy_generate(Wc, x1.float(), 8, n_heads=28, n_kv=4, theta=lcfg.rope_theta, eps=lcfg.rms_norm_eps,
                                   return_logits=True)
    mega = eng._mega
    try:
        eng._mega = None
        eng._graphs.clear()
        c = eng.generate(x1.to(cuda), 8)
    finally:
        eng._mega = mega
        eng._graphs.clear()
    _check_ids(c[0].tolist(), ids1[0].tolist(), lg1, "gemv chain")


def test_qwen2_7b_width_25_row_splitk_chain_vs_oracle(cuda, qwen7b_2layers):
    """25 sequences in lock-step (the caption batch of the bench): tcgen05 split-K GEMMs + ordered reduce + batched
    decode attention over 25 caches at full width; every row vs its own oracle run."""
    lcfg, W, Wc, eng = qwen7b_2layers
    xs = [_prompt(Wc, 16, 24 + 2 * i, 100 + i) for i in range(25)]
    outs = eng.generate_batch([x.to(cuda) for x in xs], 6)
    assert len(outs) == 25 and not eng._use_mega(25)
    n_div = 0
    for i, (x, o) in enumerate(zip(xs, outs)):
        ids, lg = om.greedy_generate(Wc, x.float(), 6, n_heads=28, n_kv=4, theta=lcfg.rope_theta, eps=lcfg.rms_norm_eps,
                                     return_logits=True)
        n_div += _check_ids(o[0].tolist(), ids[0].tolist(), lg, f"row {i}") is not None
    record("qwen2_7b_width_25_rows", rows=25, tokens_per_row=6, rows_with_a_sub_resolution_divergence=n_div)
    assert n_div <= 5, n_div        # ties below fp16 resolution are rare events, not the norm


def test_fused_splitk_chain_gives_the_same_ids_at_7b_width(cuda, qwen7b_2layers):
    """Opt-in scb_linear_splitk (reduction inside the GEMM launch) under the 25-row chain: the finished rows are
    bit-identical to the two-launch path, the normalised rows differ at most in the order of the sum of squares —
    same greedy ids on every row."""
    from streamchat_b200 import ops
    lcfg, W, Wc, eng = qwen7b_2layers
    xs = [_prompt(Wc, 16, 24 + 2 * i, 100 + i).to(cuda) for i in range(25)]
    base = eng.generate_batch(xs, 6)
    assert eng._tail_ws is None
    eng._tail_ws = ops.splitk_tail_workspace(eng.max_batch, eng._part.shape[2], cuda)
    eng._graphs.clear()
    try:
        fused = eng.generate_batch(xs, 6)
    finally:
        eng._tail_ws = None
        eng._graphs.clear()
    same = sum(a[0].tolist() == b[0].tolist() for a, b in zip(base, fused))
    assert same >= 24, same      # a 1-ulp difference of a normalised value may flip a sub-resolution tie


def test_native_prefill_equals_per_launch_prefill_at_7b_width(cuda, qwen7b_2layers, monkeypatch):
    """scb_qwen2_prefill (one native call) and the per-launch Python prefill issue the same kernels: bit-identical
    logits and KV rows at full width, including a continuation (pos0 > 0)."""
    lcfg, W, Wc, eng = qwen7b_2layers
    x1, x2 = _prompt(Wc, 32, 300, 7).to(cuda), _prompt(Wc, 10, 40, 8).to(cuda)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SCB_NATIVE_PREFILL", mode)
        eng.reset()
        la = eng.prefill(x1, slot=1).clone()
        lb = eng.prefill(x2, slot=1).clone()
        n = x1.shape[0] + x2.shape[0]
        res[mode] = (la, lb, [k[1, :, :n].clone() for k in eng.k_cache], [v[1, :, :n].clone() for v in eng.v_cache])
    eng.reset()
    assert torch.equal(res["1"][0], res["0"][0]) and torch.equal(res["1"][1], res["0"][1])
    for ka, kb in zip(res["1"][2] + res["1"][3], res["0"][2] + res["0"][3]):
        assert torch.equal(ka, kb)


def test_stacked_prefill_equals_one_prompt_at_a_time_at_7b_width(cuda, qwen7b_2layers, monkeypatch):
    """scb_qwen2_prefill_batch (several prompts stacked, every GEMM once over the stack) leaves the logits and KV rows
    of the one-prompt-at-a-time prefill, bit for bit, and generate_batch gives the same ids either way."""
    lcfg, W, Wc, eng = qwen7b_2layers
    xs = [_prompt(Wc, 12, n, 40 + n).to(cuda) for n in (300, 37, 140)]
    eng.reset()
    want = []
    for q, x in enumerate(xs):
        lg = eng.prefill(x, slot=q + 2).clone()
        want.append((lg, [k[q + 2, :, : x.shape[0]].clone() for k in eng.k_cache], [v[q + 2, :, : x.shape[0]].clone() for v in eng.v_cache]))
    eng.reset()
    logits = torch.zeros((3, lcfg.vocab_size), dtype=torch.float16, device=cuda)
    eng._mega.prefill_batch(torch.cat(xs, 0), [int(x.shape[0]) for x in xs], [5, 6, 7], [0, 0, 0], logits)
    for q, x in enumerate(xs):
        assert torch.equal(logits[q:q + 1], want[q][0])
        for l in range(len(eng.k_cache)):
            assert torch.equal(eng.k_cache[l][5 + q, :, : x.shape[0]], want[q][1][l])
            assert torch.equal(eng.v_cache[l][5 + q, :, : x.shape[0]], want[q][2][l])
    monkeypatch.setenv("SCB_PREFILL_GROUP", "1")
    a = eng.generate_batch(xs, 5)
    monkeypatch.setenv("SCB_PREFILL_GROUP", "0")
    b = eng.generate_batch(xs, 5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # argument validation: a cache slot twice in one batch
    from streamchat_b200 import _lib as L
    with pytest.raises(L.ScbError, match="twice"):
        eng._mega.prefill_batch(torch.cat(xs[:2], 0), [int(xs[0].shape[0]), int(xs[1].shape[0])], [1, 1], [0, 0], logits)
