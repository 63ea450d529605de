"""Every name inference_streaming_longva_v2.py:19-27 imports from the reference's utiles.py.
Hot-path functions come from streamchat_b200 (sm_100a kernels behind them); names outside the hot path resolve
lazily to the reference's own utiles.py when it is importable, else to placeholders that raise when called
(streamchat_b200/_offpath.py)."""
from streamchat_b200._offpath import resolver as _resolver
from streamchat_b200.kmeans import weighted_kmeans_feature  # noqa: F401
from streamchat_b200.memory import (MultimodalTreeNode, calculate_forgetting_probabilities,  # noqa: F401
                                    count_nodes_by_depth, fast_building_memory_tree_summarize_token,
                                    fast_search_tree_multi_modal_with_embedding,
                                    select_data_without_replacement)
from streamchat_b200.memory_bank import build_prompt_with_search_memory_only_related  # noqa: F401

# terminal colours the driver prints with (utiles.py:31-38)
RED, GREEN, YELLOW, BLUE = "\033[31m", "\033[32m", "\033[33m", "\033[34m"
PURPLE, CYAN, WHITE, RESET = "\033[35m", "\033[36m", "\033[37m", "\033[0m"


class TreeNode:
    """utiles.py:40-45 (node of the non-multimodal tree; only constructed by code off the hot path)."""

    def __init__(self, centroids, labels=None, depth=0):
        self.centroids, self.labels, self.children, self.depth = centroids, labels, [], depth


def process_images(images, image_processor, model_cfg):
    """utiles.py:71-87 (imported by the driver as `process_images_ours`, :27,511): per-image
    `image_processor.preprocess(...)["pixel_values"][0]`; several images are stacked into ONE list entry, and
    equal shapes are stacked again — so one frame gives [1,3,S,S], n > 1 frames give [1,n,3,S,S]."""
    import torch
    out = [image_processor.preprocess(im, return_tensors="pt")["pixel_values"][0] for im in images]
    if len(images) > 1:
        out = [torch.stack(out, dim=0)]
    if all(x.shape == out[0].shape for x in out):
        out = torch.stack(out, dim=0)
    return out


__getattr__ = _resolver("utiles", __file__, {
    "compute_gradients": "utiles.py:89", "Optical_flow": "utiles.py", "SSIM": "utiles.py",
    "long_short_memory_update": "utiles.py", "long_short_memory_update_with_summarize": "utiles.py",
    "search_tree": "utiles.py:909-935", "search_tree_multi_modal": "utiles.py",
    "convert_to_markdown": "utiles.py", "visualize_memory_feature_with_PCA": "utiles.py",
    "compress_spatial_features": "utiles.py:264-289", "building_memory_tree": "utiles.py",
    "buildingd_memory_tree_buttom_up": "utiles.py",
})
