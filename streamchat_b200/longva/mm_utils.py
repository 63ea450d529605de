"""Prompt tokenisation helpers (reference longva/mm_utils.py:341-360)."""
from __future__ import annotations

import torch

from .constants import IMAGE_TOKEN_INDEX


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """Tokenise `prompt`, replacing every literal "<image>" by one `image_token_index` entry.
    A leading BOS produced by the tokenizer is kept once (and stripped from later chunks)."""
    chunks = [tokenizer(piece).input_ids for piece in prompt.split("<image>")]
    bos = getattr(tokenizer, "bos_token_id", None)
    has_bos = bool(chunks) and len(chunks[0]) > 0 and bos is not None and chunks[0][0] == bos
    ids = [chunks[0][0]] if has_bos else []
    skip = 1 if has_bos else 0
    for i, chunk in enumerate(chunks):
        if i > 0:
            ids.append(image_token_index)
        ids.extend(chunk[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def get_model_name_from_path(model_path: str) -> str:
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]
