"""memory_bank/prompt_utils.py — the driver does `from memory_bank.prompt_utils import *` (:33) and uses
`only_related_prompt_dict_ego` (:785)."""
from streamchat_b200._offpath import resolver as _resolver
from streamchat_b200.memory_bank import only_related_prompt_dict_ego  # noqa: F401

boot_name_dict = {"en": "AI Companion", "cn": "AI伴侣"}             # prompt_utils.py:2
boot_actual_name_dict = {"en": "SiliconFriend", "cn": "硅基朋友"}    # prompt_utils.py:3

__all__ = ["only_related_prompt_dict_ego", "boot_name_dict", "boot_actual_name_dict"]

__getattr__ = _resolver("memory_bank.prompt_utils", __file__, {
    "output_prompt": "memory_bank/prompt_utils.py:4", "generate_user_keyword": "memory_bank/prompt_utils.py:103",
    "generate_ai_keyword": "memory_bank/prompt_utils.py:106",
})
