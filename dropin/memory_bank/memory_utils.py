from streamchat_b200.memory_bank import enter_name, save_local_memory  # noqa: F401
