"""Frame ingest on the GPU (SURVEY §8 a-V1 / §8f-2): `process_images` (utiles.py:71-87) =
`CLIPImageProcessor.preprocess` — resize shortest edge to S (Pillow bicubic), centre crop S x S, x 1/255,
(x - mean) / std, HWC -> CHW — for uint8 frames already in HBM, after the reader's BGR -> RGB swap
(inference_streaming_longva_v2.py:507-511).  The reference does this per frame on the CPU through PIL and
copies 1.35 MB of fp32 per frame to the device; here a frame crosses PCIe once as uint8 (H x W x 3) and two
launches per batch produce the fp16 tensor `encode_images` consumes.

Bit-exact with Pillow + the PIL-backend CLIPImageProcessor: the resample is Pillow's own fixed-point
arithmetic (libImaging/Resample.c: window = 2 x max(scale, 1) taps either side, coefficients normalised to
sum 1 and rounded to 22 fractional bits, horizontal pass then vertical pass, each rounded to uint8), and
everything after it is a function of one uint8 value per channel — a 3 x 256 table evaluated here exactly as
numpy does for the reference (float64 product by 1/255, float32 cast, float32 subtract and divide).  The
tables are set-up data (like a TMA descriptor); the per-pixel work runs in `csrc/preprocess.cu`.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2


def _bicubic_window(t: np.ndarray) -> np.ndarray:
    """Pillow's bicubic kernel (a = -0.5), vectorised."""
    a = -0.5
    t = np.abs(t)
    near = ((a + 2.0) * t - (a + 3.0)) * t * t + 1
    far = (((t - 5) * t + 8) * t - 4) * a
    return np.where(t < 1.0, near, np.where(t < 2.0, far, 0.0))


def resample_tables(in_size: int, out_size: int, first: int, count: int) -> Tuple[np.ndarray, np.ndarray]:
    """Taps of output samples [first, first+count) of an axis resampled in_size -> out_size:
    (bounds int32 [count, 2] = (first input index, taps), coeffs int32 [count, ksize])."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    o = np.arange(first, first + count, dtype=np.float64)
    center = (o + 0.5) * scale
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast: operands are >= 0 here
    hi = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = hi - lo
    j = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic_window((j + lo[:, None] - center[:, None] + 0.5) / fscale)
    w = np.where(j < n[:, None], w, 0.0)
    # Pillow accumulates the window sum tap by tap in double; reproduce that order (np.sum is pairwise)
    tot = np.zeros(count, dtype=np.float64)
    for c in range(ksize):
        tot = tot + w[:, c]
    w = np.where(tot[:, None] != 0.0, w / np.where(tot == 0.0, 1.0, tot)[:, None], w)
    fixed = np.where(w < 0, -0.5 + w * (1 << _PRECISION_BITS), 0.5 + w * (1 << _PRECISION_BITS)).astype(np.int64)
    fixed = np.where(j < n[:, None], fixed, 0).astype(np.int32)             # (int) cast truncates toward zero
    return np.stack([lo, n], 1).astype(np.int32), np.ascontiguousarray(fixed)


def resize_output_size(h: int, w: int, shortest_edge: int) -> Tuple[int, int]:
    """HF get_resize_output_image_size(size=int, default_to_square=False) -> (out_h, out_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


class ClipFramePreprocessor:
    """`preprocess_frames(frames uint8 [N, H, W, 3] cuda, bgr=False) -> [N, 3, S, S]` fp16 (or fp32 =
    the reference's `pixel_values`)."""

    def __init__(self, size: int = 336, mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD,
                 device="cuda:0"):
        self.size = int(size)
        self.device = torch.device(device)
        v = (np.arange(256, dtype=np.uint8) * (1 / 255)).astype(np.float32)          # rescale()
        lut = (v[None, :] - np.asarray(mean, np.float32)[:, None]) / np.asarray(std, np.float32)[:, None]
        self.lut = torch.from_numpy(np.ascontiguousarray(lut.astype(np.float32))).to(self.device)
        self._tables: Dict[Tuple[int, int], dict] = {}

    def _plan(self, h: int, w: int) -> dict:
        plan = self._tables.get((h, w))
        if plan is not None:
            return plan
        S = self.size
        oh, ow = resize_output_size(h, w, S)
        if oh < S or ow < S:
            raise ValueError(f"frame {h}x{w} resizes to {oh}x{ow}, smaller than the {S}x{S} crop")
        top, left = (oh - S) // 2, (ow - S) // 2
        plan = dict(oh=oh, ow=ow, top=top, left=left, hb=None, vb=None, row0=0, rows=h)
        dev = self.device
        if h != oh:
            vb, vk = resample_tables(h, oh, top, S)
            row0 = int(vb[:, 0].min())
            rows = int((vb[:, 0] + vb[:, 1]).max()) - row0            # input rows the kept output rows read
            vb = vb.copy()
            vb[:, 0] -= row0
            plan.update(vb=torch.from_numpy(vb).to(dev), vk=torch.from_numpy(vk).to(dev), vks=vk.shape[1],
                        row0=row0, rows=rows)
        else:
            plan.update(row0=top, rows=S)
        if w != ow:
            hb, hk = resample_tables(w, ow, left, S)
            plan.update(hb=torch.from_numpy(hb).to(dev), hk=torch.from_numpy(hk).to(dev), hks=hk.shape[1])
        self._tables[(h, w)] = plan
        return plan

    @torch.no_grad()
    def preprocess_frames(self, frames: torch.Tensor, bgr: bool = False,
                          out_dtype: torch.dtype = torch.float16) -> torch.Tensor:
        if not frames.is_cuda:
            raise L.ScbError("preprocess_frames needs CUDA tensors: there is no CPU fallback")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be uint8 [N, H, W, 3]")
        if out_dtype not in (torch.float16, torch.float32):
            raise ValueError("out_dtype must be float16 or float32")
        frames = frames.contiguous()
        n, h, w, _ = frames.shape
        S = self.size
        p = self._plan(h, w)
        dev = frames.device
        out = torch.empty((n, 3, S, S), dtype=out_dtype, device=dev)
        code = L.SCB_F16 if out_dtype == torch.float16 else L.SCB_F32
        src = frames[:, p["row0"]:p["row0"] + p["rows"]]               # rows the crop needs
        img_stride, row_stride, x_off = frames.stride(0), frames.stride(1), p["left"]
        if p["hb"] is not None:
            tmp = torch.empty((n, p["rows"], S, 3), dtype=torch.uint8, device=dev)
            ops._call("scb_resample_h_u8", dev, ops._ptr(src), n, img_stride, row_stride, p["rows"],
                      ops._ptr(p["hb"]), ops._ptr(p["hk"]), p["hks"], S, ops._ptr(tmp))
            src, img_stride, row_stride, x_off = tmp, tmp.stride(0), tmp.stride(1), 0
        if p["vb"] is not None:
            ops._call("scb_resample_v_normalize", dev, ops._ptr(src), n, img_stride, row_stride, x_off,
                      ops._ptr(p["vb"]), ops._ptr(p["vk"]), p["vks"], 0, S, ops._ptr(self.lut), int(bgr),
                      ops._ptr(out), code)
        else:
            ops._call("scb_resample_v_normalize", dev, ops._ptr(src), n, img_stride, row_stride, x_off,
                      None, None, 0, 0, S, ops._ptr(self.lut), int(bgr), ops._ptr(out), code)
        return out

    # CLIPImageProcessor-like surface for one RGB frame (utiles.py:77), computed on the device
    def preprocess(self, image, return_tensors="pt"):
        arr = np.asarray(image)
        if arr.ndim == 2:
            arr = np.stack([arr] * 3, -1)
        t = torch.from_numpy(np.ascontiguousarray(arr[..., :3])).to(self.device)
        return {"pixel_values": self.preprocess_frames(t[None], out_dtype=torch.float32)}
