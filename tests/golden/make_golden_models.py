"""Golden vectors for the model part of the hot path, produced by the REFERENCE's own classes
(longva.model.language_model.llava_qwen.LlavaQwenForCausalLM: encode_images +
generate_with_image_embedding) on top of the installed HF transformers, run on CPU in fp32 with
fp16-representable random weights.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_models.py

Import shims as in SURVEY.md §A.1 (qformer stub; ignore `config.rope_scaling = None`).
Writes tiny_longva_ref.npz: all weights (fp16), inputs and the reference outputs.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

qf = types.ModuleType("longva.model.multimodal_resampler.qformer")
qf.Qformer = object
sys.modules["longva.model.multimodal_resampler.qformer"] = qf
import transformers  # noqa: E402
from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel  # noqa: E402
from longva.model.language_model.llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM  # noqa: E402

_o = LlavaQwenConfig.__setattr__
LlavaQwenConfig.__setattr__ = lambda s, k, v: None if (k == "rope_scaling" and v is None) else _o(s, k, v)

# tiny architecture that still satisfies the kernels' shape rules (head_dim 64 / 128)
CLIP = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
            image_size=56, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
LLM = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
           num_key_value_heads=1, vocab_size=320, max_position_embeddings=4096, rope_theta=1000000.0,
           rms_norm_eps=1e-6, tie_word_embeddings=False)


def main():
    torch.manual_seed(0)
    tmp = tempfile.mkdtemp()
    vcfg = CLIPVisionConfig(**CLIP)
    CLIPVisionModel(vcfg).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56}).save_pretrained(tmp)
    cfg = LlavaQwenConfig(**LLM, mm_vision_tower=tmp, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                          mm_projector_type="mlp2x_gelu", mm_hidden_size=CLIP["hidden_size"],
                          mm_resampler_type="spatial_pool", mm_spatial_pool_mode="average",
                          mm_spatial_pool_stride=2, mm_use_im_start_end=False, mm_patch_merge_type="flat")
    model = LlavaQwenForCausalLM(cfg)
    model.get_model().get_vision_tower().load_model()
    model = model.float().eval()
    # random init with a bit more spread than HF's default so logits are not flat; biases non-zero
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).half().float())
            elif "norm" in name and "weight" in name:
                p.copy_((1 + 0.1 * torch.randn(p.shape, generator=g)).half().float())
            else:
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).half().float())
    out = {"transformers_version": np.array(transformers.__version__),
           "clip_cfg": np.array(repr(CLIP)), "llm_cfg": np.array(repr(LLM))}
    for name, p in model.state_dict().items():
        if p.dtype.is_floating_point:
            out["w:" + name] = p.half().numpy()
    images = torch.randn(3, 3, 56, 56, generator=g).half().float()
    with torch.no_grad():
        tower_out = model.get_model().get_vision_tower()(images)
        feats = model.encode_images(images)                      # [3, 4, 256]
    out["images"] = images.half().numpy()
    out["tower_out"] = tower_out.numpy()
    out["encode_images"] = feats.numpy()
    ids = torch.tensor([[5, 6, 7, -200, 8, 9, 10, 11]])
    emb = feats.reshape(-1, feats.shape[-1]).half().float()
    gen = {}
    for uc in (False, True):
        with torch.no_grad():
            gen[uc] = model.generate_with_image_embedding(
                ids, image_embeddings=[emb], modalities=["video"], do_sample=False, temperature=0,
                top_p=None, num_beams=1, max_new_tokens=12, use_cache=uc, eos_token_id=None, pad_token_id=0)
    assert torch.equal(gen[False], gen[True])                    # fact F6
    out["prompt_ids"] = ids.numpy()
    out["image_embeddings"] = emb.half().numpy()
    out["generated_ids"] = gen[True].numpy()
    # first-step logits for tolerance checks
    with torch.no_grad():
        (_, _, _, _, inputs_embeds, _) = model.prepare_inputs_embeddings_for_multimodal(
            ids, None, None, None, None, [emb], ["video"])
        logits = model(inputs_embeds=inputs_embeds).logits[0, -1]
    out["inputs_embeds"] = inputs_embeds[0].numpy()
    out["first_logits"] = logits.numpy()
    # text-only path (image_embeddings=None, utiles.py:594-605)
    tids = torch.tensor([[12, 13, 14, 15, 16]])
    with torch.no_grad():
        tgen = model.generate_with_image_embedding(tids, image_embeddings=None, modalities=["video"],
                                                   do_sample=False, max_new_tokens=6, use_cache=False,
                                                   eos_token_id=None, pad_token_id=0)
    out["text_prompt_ids"] = tids.numpy()
    out["text_generated_ids"] = tgen.numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_longva_ref.npz"), **out)
    print("encode_images", feats.shape, "generated", gen[True].tolist(), "text", tgen.tolist())
    top2 = torch.topk(logits, 2).values
    print("first-step top-2 logit margin:", float(top2[0] - top2[1]))


if __name__ == "__main__":
    main()
