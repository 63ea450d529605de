"""GPU: the native whole-stage entry points of SURVEY.md §8(b) — scb_vit_encode (encode_images as one C call) against
the per-launch Python loop over the same kernels (bit-identical), its tower output, its error behaviour, and CUDA-graph
capture (no allocation / synchronisation inside).  scb_qwen2_prefill's counterpart lives in test_gpu_fullwidth.py /
test_gpu_decode_mega.py."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _encoder(cuda, **kw):
    from streamchat_b200.longva.model.builder import random_weights
    from streamchat_b200.longva.model.llava_qwen import LlavaQwenConfig
    from streamchat_b200.qwen2 import Qwen2Config
    from streamchat_b200.vision import ClipVisionConfig, VisionEncoder
    vcfg = ClipVisionConfig(**kw)
    cfg = LlavaQwenConfig(llm=Qwen2Config(hidden_size=512, num_hidden_layers=0, vocab_size=16, num_attention_heads=4),
                          vision=vcfg)
    W = random_weights(cfg, cuda, seed=4, std=0.04)
    return VisionEncoder(W, vcfg, cuda, max_frames_per_pass=5), vcfg


@pytest.mark.parametrize("shape", ["hd64", "hd128"])
def test_vit_encode_native_equals_per_launch_loop(cuda, shape, monkeypatch):
    kw = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4, image_size=112) \
        if shape == "hd64" else dict(hidden_size=256, intermediate_size=768, num_hidden_layers=3, num_attention_heads=2,
                                     image_size=84)
    enc, vcfg = _encoder(cuda, **kw)
    torch.manual_seed(1)
    images = torch.randn(7, 3, vcfg.image_size, vcfg.image_size, device=cuda).half()      # 7 frames: passes of 5 + 2
    monkeypatch.setenv("SCB_NATIVE_VIT", "1")
    a = enc.encode_images(images)
    monkeypatch.setenv("SCB_NATIVE_VIT", "0")
    b = enc.encode_images(images)
    assert a.shape == b.shape == (7, (vcfg.grid // 2) ** 2, 512) and torch.equal(a, b)
    feats, tower = enc._native.encode(images[:3].contiguous(), want_tower=True)
    assert torch.equal(tower, enc.tower(images[:3])) and torch.equal(feats, b[:3])


def test_vit_encode_is_graph_capturable_and_validates(cuda):
    from streamchat_b200 import _lib as L
    enc, vcfg = _encoder(cuda, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                         image_size=56)
    images = torch.randn(4, 3, 56, 56, device=cuda).half()
    want = enc._native.encode(images)               # warm-up: sizes the workspace, sets kernel attributes
    lib = L.load()
    h = enc._native._h
    out = torch.zeros_like(want)
    ws = enc._native._ws
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rc = lib.scb_vit_encode(h, images.data_ptr(), 4, out.data_ptr(), None, ws.data_ptr(), ws.numel(),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    # errors: short workspace, bad frame count — status + message, nothing launched
    assert lib.scb_vit_encode(h, images.data_ptr(), 4, out.data_ptr(), None, ws.data_ptr(), 1024, None) == -3
    assert b"workspace" in lib.scb_last_error()
    assert lib.scb_vit_encode(h, images.data_ptr(), 0, out.data_ptr(), None, ws.data_ptr(), ws.numel(), None) == -1
    assert lib.scb_vit_encode_workspace_bytes(h, 8) > lib.scb_vit_encode_workspace_bytes(h, 4) > 0


def test_frame_prefetcher_overlaps_copies_without_changing_results(cuda):
    """§8f-3: segments staged on a copy stream (two buffers, ready / free events) are the frames a plain copy delivers,
    including ragged and empty segments, and a buffer is never overwritten before its consumer finished."""
    from streamchat_b200.pipeline import FramePrefetcher
    torch.manual_seed(0)
    host = torch.randn(23, 3, 8, 8).half().pin_memory()
    slices = [(0, 6), (6, 6), (6, 14), (14, 23)]
    outs = []
    pre = FramePrefetcher(host, slices, cuda)
    for i, seg in pre:
        if seg is None:
            outs.append(None)
            continue
        y = seg.float() * 2                           # consumer work enqueued on the current stream
        for _ in range(20):
            y = y + seg.float()                       # keeps reading the buffer while the next copy is queued
        outs.append(y)
    torch.cuda.synchronize()
    assert outs[1] is None and pre.bytes_copied == 23 * 3 * 8 * 8 * 2
    for (a, b), y in zip(slices, outs):
        if y is not None:
            assert torch.equal(y.cpu(), host[a:b].float() * 22)
    with pytest.raises(ValueError):
        FramePrefetcher(torch.zeros(4, 3, 8, 8), [(0, 4)], cuda)
