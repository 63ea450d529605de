"""GPU parity of the persistent decode-step kernel (scb_qwen2_decode_step, csrc/decode_mega.cu) through
the C-ABI: against the CPU oracle (oracle/longva_model.py, pinned on the reference's own golden) and
against the per-layer kernel chain on the same weights.  Token ids exact (or diverging only where the
oracle's own top-2 margin is below fp16 resolution); logits within 2e-3 of the logit scale."""
import pytest
import torch

from oracle import longva_model as om

pytestmark = pytest.mark.gpu


def _engine_pair(cuda, cfg, seed, std, max_ctx=1024, max_batch=4):
    from streamchat_b200.longva.model.builder import random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig
    from streamchat_b200.qwen2 import Qwen2Engine
    from streamchat_b200.vision import ClipVisionConfig
    full = LlavaQwenConfig(llm=cfg, vision=ClipVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1,
                                                            num_attention_heads=2, image_size=56, patch_size=14),
                           max_context=max_ctx)
    W = random_weights(full, cuda, seed=seed, std=std)
    Wl = {k: v for k, v in W.items() if not k.startswith("model.vision_tower") and "mm_projector" not in k}
    mega = Qwen2Engine(Wl, cfg, cuda, max_ctx=max_ctx, max_batch=max_batch)
    assert mega._mega is not None, mega.mega_unsupported
    chain = Qwen2Engine(Wl, cfg, cuda, max_ctx=max_ctx, max_batch=max_batch)
    chain._mega = None
    return {k: v.cpu() for k, v in Wl.items()}, mega, chain


def _check_ids(got, want, ref_logits):
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            top2 = torch.topk(ref_logits[i], 2).values
            margin = float(top2[0] - top2[1])
            assert margin < 4e-3 * float(ref_logits[i].abs().max()), (i, got, want, margin)
            return


@pytest.mark.parametrize("shape", ["tile", "split"])
def test_mega_step_vs_oracle_and_chain(cuda, shape):
    """'tile': every matrix is distributed as whole 16-row tiles (fewer tiles than SMs, K tails shorter than
    a 512-k unit).  'split': o_proj / down_proj have more tiles than SMs and are cut along K across CTAs
    (two partial slots per tile), K = 6912 leaves a ragged last unit."""
    from streamchat_b200.qwen2 import Qwen2Config
    if shape == "tile":
        cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                          num_key_value_heads=2, vocab_size=1008)
        std, T, steps = 0.05, 37, 12
    else:
        cfg = Qwen2Config(hidden_size=3072, intermediate_size=6912, num_hidden_layers=2, num_attention_heads=24,
                          num_key_value_heads=3, vocab_size=4096)
        std, T, steps = 0.02, 300, 8
    Wc, mega, chain = _engine_pair(cuda, cfg, seed=11, std=std)
    torch.manual_seed(1)
    x = (torch.randn(T, cfg.hidden_size) * 0.5).half()
    ref_ids, ref_logits = om.greedy_generate(Wc, x.float(), steps, n_heads=cfg.num_attention_heads,
                                             n_kv=cfg.num_key_value_heads, return_logits=True)
    a = mega.generate(x.to(cuda), steps, use_graph=False)
    b = chain.generate(x.to(cuda), steps, use_graph=False)
    _check_ids(a[0].tolist(), ref_ids[0].tolist(), ref_logits)
    _check_ids(b[0].tolist(), ref_ids[0].tolist(), ref_logits)
    # same prefix -> the stored logits of the last step agree between the two CUDA paths and with the oracle
    if a[0].tolist() == b[0].tolist() == ref_ids[0].tolist():
        la, lb = mega._logits[0].float().cpu(), chain._logits[0].float().cpu()
        scale = float(ref_logits[-1].abs().max())
        assert (la - lb).abs().max().item() < 2e-3 * scale + 2e-3
        assert (la - ref_logits[-1]).abs().max().item() < 4e-3 * scale + 4e-3
    # graph replay == eager launches, bit for bit
    g = mega.generate(x.to(cuda), steps, use_graph=True)
    assert torch.equal(g, a)


def test_mega_rows_are_independent(cuda):
    """1..4 sequences in lock-step through the persistent kernel: every row gets the ids (and KV cache rows)
    of its one-at-a-time run, whatever the batch width and whichever cache slot it lives in."""
    from streamchat_b200.qwen2 import Qwen2Config
    cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=1008)
    _, mega, _ = _engine_pair(cuda, cfg, seed=5, std=0.05)
    torch.manual_seed(2)
    xs = [(torch.randn(n, 512) * 0.5).half().to(cuda) for n in (9, 130, 41, 66)]
    solo = [mega.generate(x, 10) for x in xs]
    for width in (2, 3, 4):
        bat = mega.generate_batch(xs[:width], 10)
        for s, b in zip(solo, bat):
            assert torch.equal(s, b)
    # mixed budgets: rows drop off the end, the step narrows
    bat = mega.generate_batch(xs, [10, 4, 7, 2])
    for s, b, n in zip(solo, bat, (10, 4, 7, 2)):
        assert torch.equal(s[:, :n], b)


def test_mega_abi_errors(cuda):
    from streamchat_b200 import _lib as L
    from streamchat_b200.qwen2 import Qwen2Config
    cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=1, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=1008)
    _, mega, _ = _engine_pair(cuda, cfg, seed=5, std=0.05, max_batch=1)
    with pytest.raises(L.ScbError):
        mega._mega.step(5, mega.next_id, mega.state[1], mega._logits)          # rows out of range
    bad = L.Qwen2DecodeStepArgs()
    assert L.load().scb_qwen2_decode_step(bad, None) != 0                       # NULL handle
    assert b"handle" in L.load().scb_last_error()


def test_native_prefill_equals_per_launch_prefill(cuda, monkeypatch):
    """One native call (scb_qwen2_prefill) issues exactly the launches of Qwen2Engine.prefill: logits of the last
    position and the KV cache rows must be bit-identical, including a chunked continuation (pos0 > 0)."""
    from streamchat_b200.qwen2 import Qwen2Config
    cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=1008)
    _, a, b = _engine_pair(cuda, cfg, seed=9, std=0.05, max_batch=2)
    torch.manual_seed(4)
    x1 = (torch.randn(150, 512) * 0.5).half().to(cuda)
    x2 = (torch.randn(37, 512) * 0.5).half().to(cuda)
    monkeypatch.setenv("SCB_NATIVE_PREFILL", "1")
    la1 = a.prefill(x1, slot=1).clone()
    la2 = a.prefill(x2, slot=1).clone()
    monkeypatch.setenv("SCB_NATIVE_PREFILL", "0")
    lb1 = b.prefill(x1, slot=1).clone()
    lb2 = b.prefill(x2, slot=1).clone()
    assert torch.equal(la1, lb1) and torch.equal(la2, lb2)
    for ka, kb, va, vb in zip(a.k_cache, b.k_cache, a.v_cache, b.v_cache):
        assert torch.equal(ka[1, :, :187], kb[1, :, :187]) and torch.equal(va[1, :, :187], vb[1, :, :187])
