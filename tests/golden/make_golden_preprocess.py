"""Golden vectors for frame ingest (a-V1), produced by Pillow + the PIL-backend CLIPImageProcessor (the
arithmetic `process_images`, utiles.py:71-87, runs in the reference; `CLIPImageProcessorPil` is that slow
processor's name in transformers 5.x):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_preprocess.py

Writes preprocess_ref.npz: uint8 frames of several geometries and their `pixel_values` at crop size 56
(down-scale landscape / portrait, up-scale, identity size, one odd aspect ratio).
"""
import os

import numpy as np
import PIL
import transformers
from PIL import Image
from transformers import CLIPImageProcessorPil

HERE = os.path.dirname(os.path.abspath(__file__))
S = 56
proc = CLIPImageProcessorPil(size={"shortest_edge": S}, crop_size={"height": S, "width": S})
rng = np.random.default_rng(0)
out = {"size": np.array(S), "versions": np.array(f"Pillow {PIL.__version__}, transformers {transformers.__version__}")}
for i, (h, w) in enumerate([(90, 120), (130, 70), (40, 50), (56, 56), (57, 211), (200, 56)]):
    # smooth structure + noise so that both the anti-aliased and the interpolating regime matter
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(yy[..., None] / (3.0 + i) + np.arange(3)) * np.cos(xx[..., None] / (5.0 + i))
    frame = np.clip(base + rng.normal(0, 25, (h, w, 3)), 0, 255).astype(np.uint8)
    out[f"frame{i}"] = frame
    out[f"pixel_values{i}"] = proc.preprocess(Image.fromarray(frame), return_tensors="pt")["pixel_values"][0].numpy()
np.savez_compressed(os.path.join(HERE, "preprocess_ref.npz"), **out)
print("wrote preprocess_ref.npz", {k: v.shape for k, v in out.items()})
