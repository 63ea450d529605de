from streamchat_b200.memory_bank import LocalMemoryRetrieval  # noqa: F401
