"""The drop-in proof (GPU): the reference driver's OWN functions — `updating_memory_buffer`
(inference_streaming_longva_v2.py:267-378) and `longva_inference_with_embedding_multi_modal` (:164-264) — are
AST-extracted from the unmodified driver source (staged, git-ignored, under oracle/_ref/ by oracle/stage_ref.py) and
executed over dropin/: every name they use resolves through the reference's own import paths
(`from longva.constants import ...`, `from utiles import ...`) to streamchat_b200.  The tiny model comes from a
`save_pretrained` checkpoint written by the reference's model class (tests/golden/tiny_ckpt), loaded through the
drop-in `load_pretrained_model`; its encode_images / greedy ids must equal the reference golden."""
import ast
import importlib
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "inference_streaming_longva_v2.py")


@pytest.fixture(scope="module")
def dropin_modules():
    """dropin/ first on sys.path for this module only (the shadow packages must not leak into other tests)."""
    shadow = ("longva", "utiles", "memory_bank", "torch_kmeans")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in shadow}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        yield
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for k in [k for k in sys.modules if k.split(".")[0] in shadow]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.fixture(scope="module")
def tiny_from_checkpoint(cuda, golden_dir, dropin_modules):
    from longva.model.builder import load_pretrained_model          # the reference's import path, served by dropin/
    return load_pretrained_model(os.path.join(golden_dir, "tiny_ckpt"), None, "llava_qwen", device_map="cuda:0",
                                 overwrite_config={"max_context": 2048, "max_batch": 8})


def test_load_pretrained_model_reads_a_save_pretrained_checkpoint(cuda, golden_dir, tiny_from_checkpoint):
    """Row a-L5: config.json / generation_config.json / model.safetensors / tokenizer / vision_tower/ of a checkpoint the
    reference's LlavaQwenForCausalLM.save_pretrained wrote (ref builder.py:202-213,258-285)."""
    import streamchat_b200
    tokenizer, model, image_processor, context_len = tiny_from_checkpoint
    assert type(model).__module__.startswith("streamchat_b200")
    g = np.load(os.path.join(golden_dir, "tiny_longva_ref.npz"))
    assert len(tokenizer) == 320 and tokenizer.convert_tokens_to_ids("<im_patch>") == 319     # builder.py:261-266
    assert model.config.eos_token_id == [317, 318] and model.config.pad_token_id == 318        # generation_config.json
    assert model.config.llm.rope_theta == 1e6 and model.config.vision.image_size == 56 and context_len == 2048
    assert tuple(image_processor.preprocess(np.zeros((80, 64, 3), np.uint8), return_tensors="pt")["pixel_values"].shape) == (1, 3, 56, 56)
    images = torch.from_numpy(g["images"]).to(cuda)
    feats = model.encode_images(images).float().cpu()
    ref = torch.from_numpy(g["encode_images"])
    assert float((feats - ref).norm() / ref.norm()) < 1.5e-3
    ids = torch.from_numpy(g["prompt_ids"])
    emb = torch.from_numpy(g["image_embeddings"]).to(cuda)
    kw = dict(image_embeddings=[emb], modalities=["video"], do_sample=False, temperature=0, top_p=None, num_beams=1,
              max_new_tokens=12, use_cache=False)
    out = model.generate_with_image_embedding(ids, eos_token_id=None, **kw)          # the golden was made without a stop id
    np.testing.assert_array_equal(out.cpu().numpy(), g["generated_ids"])
    # default: the checkpoint's own stop ids end the generation, the stop token included (HF generate under
    # llava_qwen.py:155) — make a token of the golden continuation the stop id
    want = g["generated_ids"][0].tolist()
    stop = want[5]
    model.config.eos_token_id = [stop, 318]
    try:
        out = model.generate_with_image_embedding(ids, **kw)
        assert out[0].tolist() == want[:want.index(stop) + 1]
        outs = model.generate_requests([(ids, [emb], 12), (ids, [emb], 12)], modalities=["video"])
        assert all(o[0].tolist() == want[:want.index(stop) + 1] for o in outs)
    finally:
        model.config.eos_token_id = [317, 318]
    # a generation that cannot fit the KV-cache slot fails before anything is launched
    with pytest.raises(ValueError, match="max_ctx"):
        model.generate_with_image_embedding(ids, image_embeddings=[emb], max_new_tokens=4096, do_sample=False)


def _extract(names):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref not staged (python -m oracle.stage_ref needs the reference tree)")
    tree = ast.parse(open(DRIVER).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in body) == sorted(names)
    return ast.Module(body=body, type_ignores=[])


def test_reference_driver_functions_run_over_dropin(cuda, tiny_from_checkpoint, capsys):
    """updating_memory_buffer + longva_inference_with_embedding_multi_modal, verbatim, over dropin/ — and they agree
    with streamchat_b200.stream.StreamingSession (the restatement bench.py drives) on the same inputs."""
    tokenizer, model, _, _ = tiny_from_checkpoint
    import streamchat_b200.memory as M
    from streamchat_b200.stream import HashTextEmbedder, StreamConfig, StreamingSession
    ns = {"torch": torch, "time": __import__("time"), "np": np,
          "args": types.SimpleNamespace(temperature=0, top_p=None, num_beams=1)}
    # the driver's own import lines for the names the two functions use (:13-26)
    exec("from longva.constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN\n"
         "from longva.conversation import conv_templates, SeparatorStyle\n"
         "from longva.mm_utils import tokenizer_image_token\n"
         "from utiles import fast_search_tree_multi_modal_with_embedding, calculate_forgetting_probabilities, "
         "select_data_without_replacement, weighted_kmeans_feature, fast_building_memory_tree_summarize_token, "
         "count_nodes_by_depth, RED, RESET, BLUE, GREEN\n", ns)
    exec(compile(_extract(["updating_memory_buffer", "longva_inference_with_embedding_multi_modal"]), DRIVER, "exec"), ns)
    assert ns["fast_building_memory_tree_summarize_token"] is M.fast_building_memory_tree_summarize_token

    g = torch.Generator(device=cuda).manual_seed(0)
    frames = torch.randn(88, 3, 56, 56, device=cuda, dtype=torch.float16, generator=g)
    embedder = HashTextEmbedder(1024, cuda)
    question = "w5 w17 w200 w9"
    saved = (M.CAPTION_TOKENS, M.SUMMARY_TOKENS, M.CAPTION_DO_SAMPLE)
    M.CAPTION_TOKENS, M.SUMMARY_TOKENS, M.CAPTION_DO_SAMPLE = 8, 8, False      # greedy override of parity runs (H2)
    eos = model.config.eos_token_id
    model.config.eos_token_id = None                                           # equal work in both runs
    try:
        def seed():
            torch.manual_seed(3); np.random.seed(7); random.seed(5)
        # ---- the driver's functions
        seed()
        tree, short = None, None
        for s0 in range(0, 88, 22):
            feats = list(torch.split(model.encode_images(frames[s0:s0 + 22].half()), 1))      # :520-525
            tree, short = ns["updating_memory_buffer"](feats, tree, model, tokenizer, True, short_window=20,
                                                       remember_window=5, tau=5, chunk_size=8, num_clusters=5, interval=3)
        outputs, _, _ = ns["longva_inference_with_embedding_multi_modal"](
            question, 5, "qwen_1_5_ego", model, embedder, tokenizer, None, None, short, tree)
        texts_a = [n.text for n in tree]
        # ---- the session object bench.py drives
        seed()
        sess = StreamingSession(model, tokenizer, embedder, None, None,
                                StreamConfig(chunk_size=8, interval=3, answer_tokens=256))
        for s0 in range(0, 88, 22):
            sess.update_memory(sess.ingest(frames[s0:s0 + 22]))
        text_b, ids_b = sess.answer(question)
    finally:
        M.CAPTION_TOKENS, M.SUMMARY_TOKENS, M.CAPTION_DO_SAMPLE = saved
        model.config.eos_token_id = eos
    depths = dict(M.count_nodes_by_depth(tree))
    assert max(depths) >= 1 and sum(depths.values()) >= 11                     # merges happened
    assert texts_a == [n.text for n in sess.long_memory_tree]
    for a, b in zip(tree, sess.long_memory_tree):
        assert torch.equal(a.centroids, b.centroids)
    assert len(short) == 5 and all(torch.equal(a, b) for a, b in zip(short, sess.short_memory_buffer))
    assert isinstance(outputs, str) and outputs == text_b and ids_b.shape == (1, 256)
    assert "memory update finish" in capsys.readouterr().out                   # the driver's own prints ran
