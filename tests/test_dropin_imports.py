"""The documented drop-in run puts dropin/ first on PYTHONPATH; the import block of the
reference driver (inference_streaming_longva_v2.py:13-33) must then resolve completely — hot-path names to
streamchat_b200, the rest to the reference's own modules or to call-time placeholders.  Run in a subprocess so
the shadow packages do not leak into the test session.  (Third-party modules of the driver that are neither
in the reference tree nor here — torchaudio, decord — are not dropin's business.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the same names, module by module, as the driver's import statements
DRIVER_IMPORTS = r'''
from longva.constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN
from longva.conversation import conv_templates, SeparatorStyle
from longva.model.builder import load_pretrained_model
from longva.utils import disable_torch_init
from longva.mm_utils import tokenizer_image_token, process_images, get_model_name_from_path
from utiles import compute_gradients, Optical_flow, SSIM, long_short_memory_update, \
    long_short_memory_update_with_summarize, search_tree, search_tree_multi_modal, \
    fast_search_tree_multi_modal_with_embedding, MultimodalTreeNode, TreeNode, \
    build_prompt_with_search_memory_only_related, convert_to_markdown, visualize_memory_feature_with_PCA, \
    calculate_forgetting_probabilities, select_data_without_replacement, compress_spatial_features, \
    weighted_kmeans_feature, fast_building_memory_tree_summarize_token, count_nodes_by_depth, RED, RESET, BLUE, GREEN
from utiles import process_images as process_images_ours
from memory_bank.memory_retrieval.local_doc_qa import LocalMemoryRetrieval
from memory_bank.memory_utils import summarize_memory_event_personality, enter_name, save_local_memory
from memory_bank.summarize_memory import LLMClientLLaMA3
from memory_bank.prompt_utils import *
from torch_kmeans import KMeans

import streamchat_b200.memory as M, streamchat_b200.kmeans as K
assert fast_building_memory_tree_summarize_token is M.fast_building_memory_tree_summarize_token
assert weighted_kmeans_feature is K.weighted_kmeans_feature and KMeans is K.KMeans
assert callable(only_related_prompt_dict_ego) and RED.startswith("\033[")
disable_torch_init()
try:                                   # off-path names import fine and fail only when called
    compute_gradients(None)
except NotImplementedError as e:
    assert "outside the hot path" in str(e)
else:
    raise SystemExit("placeholder did not raise")
import numpy as np
from PIL import Image
class P:                               # CLIPImageProcessor contract (utiles.py:77)
    def preprocess(self, im, return_tensors="pt"):
        import torch
        return {"pixel_values": torch.zeros(1, 3, 4, 4)}
one = process_images_ours([Image.fromarray(np.zeros((4, 4, 3), np.uint8))], P(), None)
two = process_images_ours([Image.fromarray(np.zeros((4, 4, 3), np.uint8))] * 2, P(), None)
assert tuple(one.shape) == (1, 3, 4, 4) and tuple(two.shape) == (1, 2, 3, 4, 4)
print("DROPIN-IMPORTS-OK")
'''


def test_driver_import_block_resolves_against_dropin():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT])
    r = subprocess.run([sys.executable, "-c", DRIVER_IMPORTS], env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.join(ROOT, "tests"))
    assert r.returncode == 0 and "DROPIN-IMPORTS-OK" in r.stdout, r.stdout + r.stderr
