from streamchat_b200.longva.mm_utils import get_model_name_from_path, tokenizer_image_token  # noqa: F401
