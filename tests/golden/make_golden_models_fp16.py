"""The reference's own fp16 execution (`model.half()`, as the driver runs it on the GPU:
longva/model/builder.py:213, inference_streaming_longva_v2.py:523) of the tiny golden model, run here on the
CPU: encode_images and the first-step logits.  Pins the ERROR SCALE of fp16 execution against the fp32 golden —
the yardstick the GPU parity tolerances are calibrated with (tests/test_oracle_models.py,
tests/test_gpu_fullwidth.py) — and the rounding points of the oracle's `act_dtype=torch.float16` mode.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_models_fp16.py

Reads tiny_longva_ref.npz (weights + inputs), writes tiny_longva_ref_fp16.npz.
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


def main():
    g = np.load(os.path.join(HERE, "tiny_longva_ref.npz"))
    clip, llm = eval(str(g["clip_cfg"])), eval(str(g["llm_cfg"]))
    tmp = tempfile.mkdtemp()
    CLIPVisionModel(CLIPVisionConfig(**clip)).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56}).save_pretrained(tmp)
    cfg = LlavaQwenConfig(**llm, mm_vision_tower=tmp, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                          mm_projector_type="mlp2x_gelu", mm_hidden_size=clip["hidden_size"],
                          mm_resampler_type="spatial_pool", mm_spatial_pool_mode="average",
                          mm_spatial_pool_stride=2, mm_use_im_start_end=False, mm_patch_merge_type="flat")
    model = LlavaQwenForCausalLM(cfg)
    model.get_model().get_vision_tower().load_model()
    sd = model.state_dict()
    for k in list(sd):
        if "w:" + k in g.files:
            sd[k] = torch.from_numpy(g["w:" + k]).float()
    model.load_state_dict(sd)
    model = model.half().eval()
    with torch.no_grad():
        feats = model.encode_images(torch.from_numpy(g["images"]).half())
        logits = model(inputs_embeds=torch.from_numpy(g["inputs_embeds"])[None].half()).logits[0, -1]
    np.savez_compressed(os.path.join(HERE, "tiny_longva_ref_fp16.npz"),
                        transformers_version=np.array(transformers.__version__), torch_version=np.array(torch.__version__),
                        encode_images_fp16=feats.numpy(), first_logits_fp16=logits.numpy())
    print("wrote tiny_longva_ref_fp16.npz", feats.dtype, logits.dtype)


if __name__ == "__main__":
    main()
