from .llava_qwen import LlavaQwenConfig, LlavaQwenForCausalLM  # noqa: F401
