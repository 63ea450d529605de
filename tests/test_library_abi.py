"""CPU: the C-ABI library loads and exports every symbol include/streamchat_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from streamchat_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_symbols_exported():
    _lib = _ensure_built()
    header = open(os.path.join(ROOT, "include", "streamchat_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_version_and_error_string():
    _lib = _ensure_built()
    lib = _lib.load()
    assert lib.scb_version() >= 100
    assert isinstance(lib.scb_last_error(), bytes)
    # argument validation happens before any CUDA call: NULL args -> INVALID_ARG + message
    assert lib.scb_cosine_topk(None, None) == -1
    assert b"NULL" in lib.scb_last_error()
    assert lib.scb_cosine_topk_workspace_bytes(10000, 1, 8) > 0
    assert lib.scb_cosine_topk_workspace_bytes(10000, 1, 64) == 0
    assert lib.scb_kmeans_step_workspace_bytes(400, 5, 1, 516096) > 0
    # handle-based entry points reject a NULL / foreign handle before touching CUDA
    assert lib.scb_qwen2_prefill_workspace_bytes(None, 16) == 0
    assert lib.scb_qwen2_prefill(None, None, 16, 0, 0, None, None, 0, None) == -1
    assert b"handle" in lib.scb_last_error()
    assert lib.scb_qwen2_decoder_destroy(None) == -1
    assert lib.scb_qwen2_prefill_batch(None, None, 2, None, None, None, None, 0, None, 0, None) == -1
    assert lib.scb_qwen2_decoder_max_rows() >= 1
    assert lib.scb_vit_encode_workspace_bytes(None, 4) == 0
    assert lib.scb_vit_encode(None, None, 4, None, None, None, 0, None) == -1
    assert b"handle" in lib.scb_last_error()
    assert lib.scb_vit_encoder_destroy(None) == -1
    assert lib.scb_kmeans_assign_update(None, None) == -1          # SURVEY §8(b)'s name for scb_kmeans_step


def test_product_path_has_no_cpu_fallback():
    import torch
    from streamchat_b200 import ops, _lib
    with pytest.raises(_lib.ScbError):
        ops.cosine_topk(torch.randn(4, 8), torch.randn(8), k=1)
    with pytest.raises(_lib.ScbError):
        ops.kmeans_step(torch.randn(4, 8), torch.randn(1, 2, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "streamchat_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product modules import the oracle: {bad}"


def test_ctypes_structs_match_header_layout(tmp_path):
    """Every argument struct of the header, compiled by gcc, has the size and the field offsets of its ctypes
    mirror in streamchat_b200/_lib.py (a renamed or re-ordered field fails to compile or to compare)."""
    import subprocess

    from streamchat_b200 import _lib
    pairs = {
        "scb_cosine_topk_args": _lib.CosineTopkArgs, "scb_kmeans_step_args": _lib.KmeansStepArgs,
        "scb_linear_args": _lib.LinearArgs, "scb_attention_args": _lib.AttentionArgs, "scb_gemv_args": _lib.GemvArgs,
        "scb_decode_attention_args": _lib.DecodeAttentionArgs,
        "scb_decode_attention_fused_args": _lib.DecodeAttentionFusedArgs,
        "scb_qwen2_decoder_desc": _lib.Qwen2DecoderDesc, "scb_qwen2_decode_step_args": _lib.Qwen2DecodeStepArgs,
        "scb_vit_encoder_desc": _lib.VitEncoderDesc, "scb_linear_splitk_args": _lib.LinearSplitKArgs,
        "scb_allgather_embed_args": _lib.AllgatherEmbedArgs,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "streamchat_b200.h")}"',
             'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True, capture_output=True)
    got = {}
    for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
