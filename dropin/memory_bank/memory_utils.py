"""memory_bank/memory_utils.py names the driver imports (inference_streaming_longva_v2.py:29)."""
from streamchat_b200._offpath import resolver as _resolver
from streamchat_b200.memory_bank import enter_name, save_local_memory  # noqa: F401

__getattr__ = _resolver("memory_bank.memory_utils", __file__, {
    # LLM-written personality / event summaries of the dialogue history: an offline step, not on the per-frame path
    "summarize_memory_event_personality": "memory_bank/memory_utils.py:30-58",
    "get_docs_with_score": "memory_bank/memory_utils.py",
})
