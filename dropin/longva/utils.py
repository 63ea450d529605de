"""longva/utils.py names the driver imports (inference_streaming_longva_v2.py:17,688)."""
from streamchat_b200._offpath import resolver as _resolver


def disable_torch_init():
    """longva/utils.py:133-140 skips torch.nn's default weight initialisation to speed up HF model construction.
    The sm_100a engines build no torch.nn modules, so there is nothing to disable; kept for the call at :688."""
    import torch
    setattr(torch.nn.Linear, "reset_parameters", lambda self: None)
    setattr(torch.nn.LayerNorm, "reset_parameters", lambda self: None)


__getattr__ = _resolver("longva.utils", __file__, {
    "build_logger": "longva/utils.py", "violates_moderation": "longva/utils.py:143",
    "pretty_print_semaphore": "longva/utils.py", "rank0_print": "longva/utils.py",
})
