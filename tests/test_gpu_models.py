"""GPU parity of the model path: CUDA engines (through the drop-in LongVA surface) vs
(1) golden vectors produced by the reference's own LlavaQwenForCausalLM and (2) the CPU oracle.
Tolerances (north_star): fp features within 1e-3 relative (of the feature scale), greedy ids exact."""
import os

import numpy as np
import pytest
import torch

from oracle import longva_model as om

pytestmark = pytest.mark.gpu


def _tiny(golden_dir, cuda):
    from streamchat_b200.longva.model.builder import SYNTHETIC
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenForCausalLM
    g = np.load(os.path.join(golden_dir, "tiny_longva_ref.npz"))
    W = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    model = LlavaQwenForCausalLM(SYNTHETIC["tiny"], W, cuda)
    return g, W, model


def _rel_err(out, ref):
    scale = ref.abs().max().item()
    return (out - ref).abs().max().item() / scale


def test_encode_images_matches_reference_golden(cuda, golden_dir):
    g, W, model = _tiny(golden_dir, cuda)
    images = torch.from_numpy(g["images"]).to(cuda)
    tower = model.vision.tower(images).view(3, 17, 128)[:, 1:].float().cpu()
    assert _rel_err(tower, torch.from_numpy(g["tower_out"])) < 2e-3
    feats = model.encode_images(images)
    assert feats.shape == (3, 4, 256) and feats.dtype == torch.float16
    assert _rel_err(feats.float().cpu(), torch.from_numpy(g["encode_images"])) < 2e-3


def test_generate_matches_reference_golden(cuda, golden_dir):
    g, W, model = _tiny(golden_dir, cuda)
    ids = torch.from_numpy(g["prompt_ids"])
    emb = torch.from_numpy(g["image_embeddings"]).to(cuda)
    (_, _, _, _, x, _) = model.prepare_inputs_embeddings_for_multimodal(ids, None, None, None, None, [emb], ["video"])
    assert torch.equal(x[0].cpu(), torch.from_numpy(g["inputs_embeds"]).half())
    logits = model.llm.prefill(x[0]).float().cpu()[0]
    model.llm.reset()
    ref = torch.from_numpy(g["first_logits"])
    assert (logits - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 2e-3
    for use_cache in (False, True):       # accepted and ignored (fact F6)
        out = model.generate_with_image_embedding(ids, image_embeddings=[emb], modalities=["video"],
                                                  do_sample=False, temperature=0, top_p=None, num_beams=1,
                                                  max_new_tokens=12, use_cache=use_cache)
        np.testing.assert_array_equal(out.cpu().numpy(), g["generated_ids"])
    # text-only path
    tout = model.generate_with_image_embedding(torch.from_numpy(g["text_prompt_ids"]), image_embeddings=None,
                                               modalities=["video"], do_sample=False, max_new_tokens=6,
                                               use_cache=False)
    np.testing.assert_array_equal(tout.cpu().numpy(), g["text_generated_ids"])
    # graph replay and eager launches agree bit for bit
    x0 = model.prepare_inputs_embeddings_for_multimodal(ids, None, None, None, None, [emb], ["video"])[4][0]
    a = model.llm.generate(x0, 12, use_graph=True)
    b = model.llm.generate(x0, 12, use_graph=False)
    assert torch.equal(a, b)
    # sampling path runs and returns valid ids (not a parity path, hazard H2)
    s = model.generate_with_image_embedding(ids, image_embeddings=[emb], do_sample=True, temperature=0.1,
                                            max_new_tokens=5, use_cache=False)
    assert s.shape == (1, 5) and int(s.min()) >= 0 and int(s.max()) < 320


def test_no_image_token_drops_features_h7(cuda, golden_dir):
    g, W, model = _tiny(golden_dir, cuda)
    emb = torch.from_numpy(g["image_embeddings"]).to(cuda)
    ids = torch.tensor([[12, 13, 14, 15, 16]])
    x = model.prepare_inputs_embeddings_for_multimodal(ids, None, None, None, None, [emb], ["video"])[4]
    assert x.shape == (1, 5, 256)
    ref = om.prepare_inputs_embeds(W, ids, [emb.cpu().float()])
    assert torch.equal(x[0].cpu().float(), ref)


def test_wider_model_vs_oracle(cuda):
    """Wider random model (CLIP 256-d/4 heads, Qwen2 512-d, GQA 4q/2kv... head_dim 128) against the
    CPU oracle on the same seeded weights: features within 2e-3 of the feature scale, greedy ids
    exact or diverging only where the oracle's own top-2 logit margin is below fp16 resolution."""
    from streamchat_b200.longva.model.builder import random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM
    from streamchat_b200.qwen2 import Qwen2Config
    from streamchat_b200.vision import ClipVisionConfig
    cfg = LlavaQwenConfig(
        llm=Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=1000),
        vision=ClipVisionConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=4,
                                num_attention_heads=4, image_size=112, patch_size=14),
        max_context=1024)
    W = random_weights(cfg, cuda, seed=3, std=0.05)
    model = LlavaQwenForCausalLM(cfg, W, cuda)
    Wc = {k: v.cpu() for k, v in W.items()}
    torch.manual_seed(0)
    images = torch.randn(2, 3, 112, 112).half()
    ref = om.encode_images(Wc, images.float(), n_heads=4)
    feats = model.encode_images(images.to(cuda))
    assert feats.shape == ref.shape == (2, 16, 512)
    assert _rel_err(feats.float().cpu(), ref) < 2e-3
    ids = torch.tensor([[7, 8, -200, 9, 10, 11]])
    emb = feats.reshape(-1, 512)
    x = om.prepare_inputs_embeds(Wc, ids, [emb.cpu().float()])
    ref_ids, ref_logits = om.greedy_generate(Wc, x, 10, n_heads=4, n_kv=2, return_logits=True)
    out = model.generate_with_image_embedding(ids, image_embeddings=[emb], do_sample=False, max_new_tokens=10)
    got, want = out[0].tolist(), ref_ids[0].tolist()
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            top2 = torch.topk(ref_logits[i], 2).values
            margin = float(top2[0] - top2[1])
            assert margin < 4e-3 * float(ref_logits[i].abs().max()), (i, got, want, margin)
            break


def test_batched_generation_equals_sequential(cuda, golden_dir):
    """Chunk captions decoded together (one weight pass per step) give, per sequence, exactly the ids
    of the one-at-a-time calls — the precondition for batching the reference's caption loop."""
    g, W, model = _tiny(golden_dir, cuda)
    torch.manual_seed(5)
    ids = torch.from_numpy(g["prompt_ids"])
    embs = [(torch.randn(n, 256) * 0.5).half().to(cuda) for n in (12, 40, 7, 25, 33)]
    seq = [model.generate_with_image_embedding(ids, image_embeddings=[e], do_sample=False, max_new_tokens=9)
           for e in embs]
    bat = model.generate_with_image_embedding_batch(ids, embs, modalities=["video"], max_new_tokens=9)
    assert len(bat) == 5
    for a, b in zip(seq, bat):
        assert torch.equal(a, b)
    # EOS truncation per sequence
    eos = int(seq[1][0, 3])
    bat2 = model.llm.generate_batch(
        [model.prepare_inputs_embeddings_for_multimodal(ids, None, None, None, None, [e], ["video"])[4][0] for e in embs],
        9, eos_token_id=eos)
    for a, b in zip(seq, bat2):
        row = a[0].tolist()
        n = row.index(eos) + 1 if eos in row else len(row)
        assert b[0].tolist() == row[:n]


def test_mixed_budget_requests_equal_sequential(cuda, golden_dir):
    """Captions (short budget, with image features) and a deferred text summary (long budget, text only)
    decoded in one batch: rows drop out as they finish; ids equal the one-at-a-time calls."""
    g, W, model = _tiny(golden_dir, cuda)
    torch.manual_seed(6)
    ids = torch.from_numpy(g["prompt_ids"])
    tids = torch.tensor([[12, 13, 14, 15, 16, 17, 18]])
    embs = [(torch.randn(n, 256) * 0.5).half().to(cuda) for n in (12, 40, 7)]
    seq = [model.generate_with_image_embedding(tids, image_embeddings=None, do_sample=False, max_new_tokens=14)]
    seq += [model.generate_with_image_embedding(ids, image_embeddings=[e], do_sample=False, max_new_tokens=6) for e in embs]
    reqs = [(ids, [embs[0]], 6), (tids, None, 14), (ids, [embs[1]], 6), (ids, [embs[2]], 6)]
    bat = model.generate_requests(reqs, modalities=["video"])
    assert [tuple(b.shape) for b in bat] == [(1, 6), (1, 14), (1, 6), (1, 6)]
    assert torch.equal(bat[1], seq[0])
    for got, want in zip([bat[0], bat[2], bat[3]], seq[1:]):
        assert torch.equal(got, want)


def test_wide_batch_generation_equals_sequential(cuda, golden_dir):
    """More than 8 sequences in lock-step take the tcgen05 path (K-sliced GEMMs + ordered reduce, batched
    decode attention over 20 caches): every row still gets the ids of its one-at-a-time run."""
    g, W, model = _tiny(golden_dir, cuda)
    torch.manual_seed(11)
    ids = torch.from_numpy(g["prompt_ids"])
    embs = [(torch.randn(5 + 3 * i, 256) * 0.5).half().to(cuda) for i in range(20)]
    seq = [model.generate_with_image_embedding(ids, image_embeddings=[e], do_sample=False, max_new_tokens=8)
           for e in embs]
    budgets = [8] * 18 + [5, 3]
    reqs = [(ids, [e], b) for e, b in zip(embs, budgets)]
    bat = model.generate_requests(reqs, modalities=["video"])
    assert len(bat) == 20
    for a, b, n in zip(seq, bat, budgets):
        assert b.shape == (1, n)
        assert torch.equal(a[:, :n], b)
