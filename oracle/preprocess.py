"""oracle/preprocess.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the frame-ingest row a-V1: `process_images` (utiles.py:71-87) =
`image_processor.preprocess(PIL image, return_tensors="pt")["pixel_values"][0]` with the checkpoint's
CLIPImageProcessor (resize shortest edge -> S bicubic, centre crop S, x 1/255, normalise), after the
reader's BGR->RGB swap (inference_streaming_longva_v2.py:507-511).

Third-party arithmetic, not vendored under /root/reference:
  * Pillow (requirements.txt pins Pillow==10.3.0; 12.2.0 in this image) `Image.resize(size, BICUBIC)`:
    libImaging/Resample.c — precompute_coeffs (support 2.0 x max(scale,1), window normalised to sum 1),
    normalize_coeffs_8bpc (fixed point, PRECISION_BITS = 22), horizontal pass then vertical pass, each
    rounding to uint8 through clip8.  An image that already has the target size is returned unchanged.
  * transformers==4.37.2 CLIPImageProcessor (PIL backend; `CLIPImageProcessorPil` in 5.5.0 here):
    get_resize_output_image_size(shortest_edge, default_to_square=False) -> (S, int(S*long/short)),
    center_crop top=(h-S)//2 left=(w-S)//2, rescale = (uint8 * (1/255) in float64) -> float32,
    normalize = (x - mean) / std in float32.
Pinned by tests/golden/preprocess_ref.npz = outputs of Pillow + CLIPImageProcessorPil run here
(tests/golden/make_golden_preprocess.py).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = [0, in_size)).
    Returns (bounds int32 [out,2] = (first input index, count), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One resampling pass over `axis` of a uint8 [H, W, C] image (int accumulation, clip8)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        x0, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(n):
            acc += src[x0 + j] * int(kk[o, j])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_h, out_w):
    """PIL `Image.resize((out_w, out_h), BICUBIC)` on a uint8 [H, W, C] array."""
    h, w = img.shape[:2]
    if (h, w) == (out_h, out_w):
        return img.copy()
    out = img
    if w != out_w:
        out = _pass(out, *precompute_coeffs(w, out_w), axis=1)     # horizontal pass first
    if h != out_h:
        out = _pass(out, *precompute_coeffs(h, out_h), axis=0)
    return out


def resize_output_size(h, w, shortest_edge):
    """transformers get_resize_output_image_size(size=int, default_to_square=False)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)     # (out_h, out_w)


def clip_preprocess(frame_u8, size, bgr=False, mean=CLIP_MEAN, std=CLIP_STD):
    """uint8 [H, W, 3] frame -> float32 [3, size, size] pixel_values (a-V1)."""
    img = frame_u8[:, :, ::-1] if bgr else frame_u8                    # cv2.cvtColor(frame, COLOR_BGR2RGB)
    h, w = img.shape[:2]
    oh, ow = resize_output_size(h, w, size)
    img = resize_bicubic_u8(np.ascontiguousarray(img), oh, ow)
    top, left = (oh - size) // 2, (ow - size) // 2
    img = img[top:top + size, left:left + size]
    x = (img * (1 / 255)).astype(np.float32)                           # rescale: float64 product, cast to float32
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
