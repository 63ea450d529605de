"""GPU parity: tcgen05 linear kernel (through the C-ABI) vs a plain PyTorch fp32 reference of
the same op.  Tolerance: fp16 operands, fp32 accumulation -> |err| <= 2e-3 * (|ref| + 1)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None, residual=None, act=0, alpha=1.0):
    y = alpha * (a.double() @ w.double().transpose(-1, -2))
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        y = torch.nn.functional.silu(y)
    if residual is not None:
        y = y + residual.double()
    return y


def _check(out, ref, tol=2e-3):
    out = out.double().cpu()
    ref = ref.cpu()
    err = (out - ref).abs()
    bound = tol * (ref.abs() + 1.0)
    bad = (err > bound).sum().item()
    assert bad == 0, f"{bad} elements out of tolerance, max err {err.max().item():.4g}"


@pytest.mark.parametrize("m,n,k", [
    (128, 128, 64), (128, 256, 64), (256, 256, 128), (128, 128, 1024),
    (577, 1024, 1024),          # one ViT frame, out_proj
    (1154, 4096, 1024),         # ViT fc1 (two frames)
    (1154, 1024, 4096),         # ViT fc2
    (300, 3584, 592),           # ragged m, k tail (592 = 9.25 x 64)
    (130, 200, 72),             # ragged everything
    (4672, 4608, 3584),         # Qwen2 fused QKV at config #4 prefill length
])
def test_linear_plain(cuda, m, n, k):
    from streamchat_b200 import ops
    torch.manual_seed(m * 7 + n * 3 + k)
    a = (torch.randn(m, k) * 0.5).half().to(cuda)
    w = (torch.randn(n, k) * 0.05).half().to(cuda)
    out = ops.linear(a, w)
    torch.cuda.synchronize()
    _check(out, _ref(a.cpu(), w.cpu()))


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_linear_epilogues(cuda, act):
    from streamchat_b200 import ops
    torch.manual_seed(act)
    m, n, k = 600, 1000, 320
    a = torch.randn(m, k).half().to(cuda)
    w = (torch.randn(n, k) * 0.1).half().to(cuda)
    bias = torch.randn(n).half().to(cuda)
    res = torch.randn(m, n).half().to(cuda)
    out = ops.linear(a, w, bias=bias, residual=res, act=act, alpha=0.5)
    _check(out, _ref(a.cpu(), w.cpu(), bias.cpu(), res.cpu(), act, 0.5), tol=3e-3)
    out32 = ops.linear(a, w, bias=bias, act=act, out_dtype=torch.float32)
    assert out32.dtype == torch.float32
    _check(out32, _ref(a.cpu(), w.cpu(), bias.cpu(), None, act), tol=1e-3)


def test_linear_batched_and_strided(cuda):
    from streamchat_b200 import ops
    torch.manual_seed(9)
    b, m, n, k = 5, 577, 577, 64          # attention-score shape: Q K^T per (frame, head)
    q = torch.randn(b, m, k).half().to(cuda)
    kk = torch.randn(b, n, k).half().to(cuda)
    out = ops.linear(q, kk, alpha=0.125, out_dtype=torch.float32)
    _check(out, _ref(q.cpu(), kk.cpu(), alpha=0.125), tol=1e-3)
    # strided views (head slices of a fused QKV buffer): lda != k
    qkv = torch.randn(700, 3 * 256).half().to(cuda)
    a = qkv[:, 0:64]
    w = qkv[:, 256:320]
    out = ops.linear(a, w, out_dtype=torch.float32)
    _check(out, _ref(a.cpu(), w.cpu()), tol=1e-3)
    # 3-D activations with a shared 2-D weight
    x = torch.randn(3, 50, 128).half().to(cuda)
    w2 = torch.randn(96, 128).half().to(cuda)
    out = ops.linear(x, w2)
    assert out.shape == (3, 50, 96)
    _check(out, _ref(x.cpu(), w2.cpu()))


def test_linear_linearity_full_size(cuda):
    """Full-size property (Qwen2 MLP width, no CPU oracle needed): f(a1 + a2) == f(a1) + f(a2)
    up to fp16 rounding, and a one-hot activation row selects a weight column exactly."""
    from streamchat_b200 import ops
    torch.manual_seed(2)
    m, n, k = 1024, 18944, 3584
    w = (torch.randn(n, k, device=cuda) * 0.02).half()
    a = torch.zeros(m, k, device=cuda, dtype=torch.float16)
    cols = torch.randint(0, k, (m,), device=cuda)
    a[torch.arange(m, device=cuda), cols] = 1.0
    out = ops.linear(a, w, out_dtype=torch.float32)
    assert torch.equal(out, w[:, cols].t().float())
    a1 = (torch.randn(m, k, device=cuda) * 0.5).half()
    a2 = (torch.randn(m, k, device=cuda) * 0.5).half()
    y1 = ops.linear(a1, w, out_dtype=torch.float32)
    y2 = ops.linear(a2, w, out_dtype=torch.float32)
    y12 = ops.linear((a1.float() + a2.float()).half(), w, out_dtype=torch.float32)
    assert (y12 - (y1 + y2)).abs().max().item() < 2e-2


def test_linear_fp16_rows_through_staged_box_stores(cuda):
    """fp16 outputs leave through shared memory + TMA box stores (32 rows x 64 columns) when their rows are 16-byte
    aligned: ragged m / n (clipped boxes), a row pitch wider than n (column slice of a bigger buffer: the neighbours
    must stay untouched), batched outputs, and the per-thread store path for an unaligned pitch."""
    from streamchat_b200 import ops
    torch.manual_seed(21)
    for m, n, k in [(33, 40, 64), (257, 72, 128), (600, 1000, 320), (95, 264, 64)]:
        a = torch.randn(m, k).half().to(cuda)
        w = (torch.randn(n, k) * 0.1).half().to(cuda)
        bias = torch.randn(n).half().to(cuda)
        res = torch.randn(m, n).half().to(cuda)
        big = torch.full((m + 3, n + 24), 7.0, dtype=torch.float16, device=cuda)
        view = big[1:m + 1, 8:8 + n]                       # pitch n + 24 (multiple of 8), 16-byte aligned start
        ops.linear(a, w, bias=bias, residual=res, act=1, out=view)
        _check(view, _ref(a.cpu(), w.cpu(), bias.cpu(), res.cpu(), 1), tol=3e-3)
        untouched = big.clone()
        untouched[1:m + 1, 8:8 + n] = 7.0
        assert bool((untouched == 7.0).all()), "a box store wrote outside the output view"
    # batched fp16 output
    b, m, n, k = 3, 100, 160, 64
    q = torch.randn(b, m, k).half().to(cuda)
    kk = torch.randn(b, n, k).half().to(cuda)
    out = ops.linear(q, kk, alpha=0.125)
    assert out.dtype == torch.float16
    _check(out, _ref(q.cpu(), kk.cpu(), alpha=0.125))
    # pitch not a multiple of 8 elements: rows are not 16-byte aligned, every thread stores its own row pieces
    m, n, k = 70, 100, 64
    a = torch.randn(m, k).half().to(cuda)
    w = (torch.randn(n, k) * 0.1).half().to(cuda)
    out = ops.linear(a, w)
    assert out.stride(0) == 100
    _check(out, _ref(a.cpu(), w.cpu()))


def test_linear_prefill_shapes_against_fp32_matmul(cuda):
    """o_proj / down_proj of a 5 816-token caption prefill (CTA-pair tiles, staged box stores, ragged m = 45.4 x 128,
    a partial last wave) against an fp32 matmul of the same fp16 operands, bias and residual included."""
    from streamchat_b200 import ops
    torch.manual_seed(31)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for m, n, k in [(5816, 3584, 3584), (5816, 3584, 18944)]:
            a = (torch.randn(m, k, device=cuda) * 0.5).half()
            w = (torch.randn(n, k, device=cuda) * 0.02).half()
            bias = torch.randn(n, device=cuda).half()
            res = torch.randn(m, n, device=cuda).half()
            out = ops.linear(a, w, bias=bias, residual=res)
            ref = a.float() @ w.float().t() + bias.float() + res.float()
            err = (out.float() - ref).abs()
            assert bool((err <= 2e-3 * (ref.abs() + 1.0)).all()), float(err.max())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
