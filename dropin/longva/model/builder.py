from streamchat_b200.longva.model.builder import load_pretrained_model  # noqa: F401
