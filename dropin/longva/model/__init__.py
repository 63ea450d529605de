from streamchat_b200.longva.model import LlavaQwenConfig, LlavaQwenForCausalLM  # noqa: F401
