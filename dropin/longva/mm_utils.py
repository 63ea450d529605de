"""longva/mm_utils.py names the driver imports (inference_streaming_longva_v2.py:18)."""
from streamchat_b200._offpath import resolver as _resolver
from streamchat_b200.longva.mm_utils import get_model_name_from_path, process_images, tokenizer_image_token  # noqa: F401

__getattr__ = _resolver("longva.mm_utils", __file__, {
    "process_highres_image": "longva/mm_utils.py", "process_anyres_image": "longva/mm_utils.py",
    "process_highres_image_crop_split": "longva/mm_utils.py", "expand2square": "longva/mm_utils.py",
    "KeywordsStoppingCriteria": "longva/mm_utils.py", "load_image_from_base64": "longva/mm_utils.py",
    "get_anyres_image_grid_shape": "longva/mm_utils.py",
})
