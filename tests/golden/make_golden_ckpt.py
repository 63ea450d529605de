"""A tiny LongVA checkpoint directory written by the REFERENCE's own model class through HF `save_pretrained`
(what `load_pretrained_model` reads: longva/model/builder.py:202-213,258-285), with the weights of
tiny_longva_ref.npz — so that loading it through streamchat_b200's builder must reproduce that golden's
encode_images / generated ids.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ckpt.py

Writes tests/golden/tiny_ckpt/: config.json, generation_config.json, model.safetensors (fp16), tokenizer files
(a 319-word WordLevel vocabulary: with the `<im_patch>` token the builder adds, len(tokenizer) == vocab_size 320),
and vision_tower/ (CLIP vision config + preprocessor_config.json; the tower's weights live in the main checkpoint,
as in LongVA-7B whose tower was unfrozen during training).
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tiny_ckpt")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

qf = types.ModuleType("longva.model.multimodal_resampler.qformer")
qf.Qformer = object
sys.modules["longva.model.multimodal_resampler.qformer"] = qf
import transformers  # noqa: E402
from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel, PreTrainedTokenizerFast  # noqa: E402
from longva.model.language_model.llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM  # noqa: E402

_o = LlavaQwenConfig.__setattr__
LlavaQwenConfig.__setattr__ = lambda s, k, v: None if (k == "rope_scaling" and v is None) else _o(s, k, v)


def main():
    g = np.load(os.path.join(HERE, "tiny_longva_ref.npz"))
    clip, llm = eval(str(g["clip_cfg"])), eval(str(g["llm_cfg"]))
    tmp = tempfile.mkdtemp()
    CLIPVisionModel(CLIPVisionConfig(**clip)).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56}).save_pretrained(tmp)
    cfg = LlavaQwenConfig(**llm, mm_vision_tower=tmp, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                          mm_projector_type="mlp2x_gelu", mm_hidden_size=clip["hidden_size"],
                          mm_resampler_type="spatial_pool", mm_spatial_pool_mode="average",
                          mm_spatial_pool_stride=2, mm_use_im_start_end=False, mm_patch_merge_type="flat",
                          unfreeze_mm_vision_tower=True)
    model = LlavaQwenForCausalLM(cfg)
    model.get_model().get_vision_tower().load_model()
    sd = model.state_dict()
    for k in list(sd):
        if "w:" + k in g.files:
            sd[k] = torch.from_numpy(g["w:" + k]).float()
    model.load_state_dict(sd)
    model = model.half().eval()
    # stop ids as a real checkpoint carries them (LongVA-7B-DPO: <|im_end|> 151645 / <|endoftext|> 151643)
    model.generation_config.eos_token_id = [317, 318]
    model.generation_config.pad_token_id = 318
    model.config.eos_token_id = 318
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    model.save_pretrained(OUT)
    # the tower directory travels with the checkpoint: make the reference relative
    vt = os.path.join(OUT, "vision_tower")
    os.makedirs(vt)
    for fn in ("config.json", "preprocessor_config.json"):
        shutil.copy(os.path.join(tmp, fn), os.path.join(vt, fn))
    c = json.load(open(os.path.join(OUT, "config.json")))
    c["mm_vision_tower"] = "vision_tower"
    json.dump(c, open(os.path.join(OUT, "config.json"), "w"), indent=2, sort_keys=True)
    # tokenizer: 319 words + <unk>; ids 317 / 318 are the stop tokens
    from tokenizers import Tokenizer, models, pre_tokenizers
    vocab = {f"w{i}": i for i in range(317)}
    vocab.update({"<|im_end|>": 317, "<|endoftext|>": 318})
    tk = Tokenizer(models.WordLevel(vocab, unk_token="w0"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<|endoftext|>", pad_token="<|endoftext|>", unk_token="w0")
    fast.save_pretrained(OUT)
    sizes = {fn: os.path.getsize(os.path.join(OUT, fn)) for fn in sorted(os.listdir(OUT)) if os.path.isfile(os.path.join(OUT, fn))}
    print("wrote", OUT, sizes, "len(tokenizer)", len(fast), "transformers", transformers.__version__)


if __name__ == "__main__":
    main()
