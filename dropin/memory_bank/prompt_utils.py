from streamchat_b200.memory_bank import only_related_prompt_dict_ego  # noqa: F401
