"""Frame ingest (a-V1): oracle/preprocess.py pinned against Pillow + CLIPImageProcessor outputs
(tests/golden/preprocess_ref.npz); the CUDA path against the same golden and, at 336, against the oracle.
uint8 / table work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as op

N_CASES = 6


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess_ref.npz"))


def test_oracle_matches_pillow_and_hf(golden_dir):
    g = _load(golden_dir)
    S = int(g["size"])
    for i in range(N_CASES):
        got = op.clip_preprocess(g[f"frame{i}"], S)
        np.testing.assert_array_equal(got, g[f"pixel_values{i}"])
    # the reader hands BGR frames and swaps them (inference_streaming_longva_v2.py:507)
    bgr = np.ascontiguousarray(g["frame0"][:, :, ::-1])
    np.testing.assert_array_equal(op.clip_preprocess(bgr, S, bgr=True), g["pixel_values0"])


def test_host_tap_tables_match_oracle():
    """The product's vectorised tap tables against the oracle's scalar restatement of Resample.c."""
    from streamchat_b200.preprocess import resample_tables, resize_output_size
    for i, o in [(640, 448), (720, 598), (200, 336), (1920, 597), (1080, 336), (337, 336), (50, 70)]:
        b, k = op.precompute_coeffs(i, o)
        b2, k2 = resample_tables(i, o, 0, o)
        np.testing.assert_array_equal(b, b2)
        np.testing.assert_array_equal(k, k2)
        b3, k3 = resample_tables(i, o, 5, 20)                   # a cropped range is a slice of the full table
        np.testing.assert_array_equal(b[5:25], b3)
        np.testing.assert_array_equal(k[5:25], k3)
    assert resize_output_size(480, 640, 336) == op.resize_output_size(480, 640, 336) == (336, 448)
    assert resize_output_size(720, 404, 336) == op.resize_output_size(720, 404, 336) == (598, 336)


@pytest.mark.gpu
def test_gpu_preprocess_matches_golden(cuda, golden_dir):
    from streamchat_b200.preprocess import ClipFramePreprocessor
    g = _load(golden_dir)
    S = int(g["size"])
    pre = ClipFramePreprocessor(S, device=cuda)
    for i in range(N_CASES):
        f = torch.from_numpy(g[f"frame{i}"]).to(cuda)
        ref = torch.from_numpy(g[f"pixel_values{i}"])
        got = pre.preprocess_frames(f[None], out_dtype=torch.float32)[0].cpu()
        assert torch.equal(got, ref), i
        half = pre.preprocess_frames(torch.stack([f, f]), out_dtype=torch.float16).cpu()
        assert torch.equal(half[0], ref.half()) and torch.equal(half[1], ref.half())
        bgr = pre.preprocess_frames(f.flip(-1)[None], bgr=True, out_dtype=torch.float32)[0].cpu()
        assert torch.equal(bgr, ref)
    pv = pre.preprocess(g["frame0"])["pixel_values"]          # CLIPImageProcessor-like surface (utiles.py:77)
    assert torch.equal(pv[0].cpu(), torch.from_numpy(g["pixel_values0"]))


@pytest.mark.gpu
def test_gpu_preprocess_full_size_vs_oracle(cuda):
    """Camera-sized frames at the model's 336 crop: down-scale with wide windows, batch of different content."""
    from streamchat_b200.preprocess import ClipFramePreprocessor
    rng = np.random.default_rng(5)
    pre = ClipFramePreprocessor(336, device=cuda)
    for h, w in [(480, 640), (720, 1280), (336, 336), (600, 338)]:
        frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        got = pre.preprocess_frames(torch.from_numpy(frames).to(cuda), out_dtype=torch.float32).cpu().numpy()
        for j in range(2):
            np.testing.assert_array_equal(got[j], op.clip_preprocess(frames[j], 336))
