"""The functions inference_streaming_longva_v2.py:19-26 imports from the reference's utiles.py."""
from streamchat_b200.kmeans import weighted_kmeans_feature  # noqa: F401
from streamchat_b200.memory import (MultimodalTreeNode, calculate_forgetting_probabilities,  # noqa: F401
                                    count_nodes_by_depth, fast_building_memory_tree_summarize_token,
                                    fast_search_tree_multi_modal_with_embedding,
                                    select_data_without_replacement)
from streamchat_b200.memory_bank import build_prompt_with_search_memory_only_related  # noqa: F401


def process_images(images, image_processor, model_cfg):
    """utiles.py:71-87 ('pad' aspect branch omitted: LongVA checkpoints use the default branch)."""
    import torch
    out = [image_processor.preprocess(im, return_tensors="pt")["pixel_values"][0] for im in images]
    return torch.stack(out, dim=0) if all(x.shape == out[0].shape for x in out) else out
