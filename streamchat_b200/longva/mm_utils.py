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


def process_images(images, image_processor, model_cfg):
    """longva/mm_utils.py:314-338 for the aspect modes LongVA checkpoints use: no `image_aspect_ratio` (the
    processor is applied to the whole list) and "pad" (expand to a mean-coloured square first).  The high-res /
    any-res tilings belong to other model families and are not on the path."""
    mode = getattr(model_cfg, "image_aspect_ratio", None)
    if mode is None:
        return image_processor(images, return_tensors="pt")["pixel_values"]
    if mode == "pad":
        from PIL import Image
        fill = tuple(int(x * 255) for x in image_processor.image_mean)
        out = []
        for im in images:
            w, h = im.size
            if w != h:
                side = max(w, h)
                sq = Image.new(im.mode, (side, side), fill)
                sq.paste(im, ((side - w) // 2, (side - h) // 2))
                im = sq
            out.append(image_processor.preprocess(im, return_tensors="pt")["pixel_values"][0])
        return torch.stack(out, dim=0) if all(x.shape == out[0].shape for x in out) else out
    raise NotImplementedError(f"image_aspect_ratio={mode!r} is not used by LongVA (longva/mm_utils.py:317-328)")
