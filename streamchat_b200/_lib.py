"""ctypes binding of libstreamchat_b200.so (the C-ABI declared in include/streamchat_b200.h).

The library is built in-tree by `make -C streamchat_b200/csrc` (see __graft_entry__.build).
There is no CPU fallback: every op raises if the library is missing or a call fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstreamchat_b200.so")

SCB_F32, SCB_F16 = 0, 1
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_SILU, ACT_SILU_PAIR = 0, 1, 2, 3, 4
EPI_NONE, EPI_SILU_PAIR = 0, 1


class ScbError(RuntimeError):
    pass


class CosineTopkArgs(C.Structure):
    _fields_ = [
        ("index", C.c_void_p), ("queries", C.c_void_p),
        ("n_rows", C.c_int64), ("dim", C.c_int32), ("n_queries", C.c_int32),
        ("ld_index", C.c_int64), ("ld_query", C.c_int64),
        ("index_dtype", C.c_int32), ("query_dtype", C.c_int32),
        ("k", C.c_int32), ("_pad", C.c_int32),
        ("out_indices", C.c_void_p), ("out_scores", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class KmeansStepArgs(C.Structure):
    _fields_ = [
        ("points", C.c_void_p), ("weights", C.c_void_p), ("centroids", C.c_void_p),
        ("n_points", C.c_int32), ("n_clusters", C.c_int32), ("n_batch", C.c_int32),
        ("points_dtype", C.c_int32), ("dim", C.c_int64),
        ("dist_eps", C.c_float), ("do_update", C.c_int32),
        ("out_labels", C.c_void_p), ("out_dist", C.c_void_p), ("out_centroids", C.c_void_p),
        ("out_wsum", C.c_void_p), ("out_shift", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class LinearArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldr", C.c_int64), ("ldo", C.c_int64),
        ("act", C.c_int32), ("out_dtype", C.c_int32), ("batch", C.c_int32), ("_pad", C.c_int32),
        ("stride_a", C.c_int64), ("stride_w", C.c_int64), ("stride_r", C.c_int64),
        ("stride_o", C.c_int64),
        ("alpha", C.c_float), ("_pad2", C.c_int32),
    ]


class LinearSplitKArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("partial", C.c_void_p), ("out", C.c_void_p), ("post_norm_gamma", C.c_void_p), ("post_norm_out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldr", C.c_int64), ("ldo", C.c_int64), ("post_norm_ld", C.c_int64),
        ("n_splits", C.c_int32), ("post_norm_eps", C.c_float),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("batch", C.c_int32), ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("q_len", C.c_int64), ("kv_len", C.c_int64),
        ("q_stride_b", C.c_int64), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_b", C.c_int64), ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_b", C.c_int64), ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_b", C.c_int64), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("scale", C.c_float), ("causal", C.c_int32),
    ]


class GemvArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("gamma", C.c_void_p), ("out", C.c_void_p),
        ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldr", C.c_int64), ("ldo", C.c_int64),
        ("rows", C.c_int32), ("prologue", C.c_int32), ("out_dtype", C.c_int32), ("eps", C.c_float),
        ("argmax_out", C.c_void_p), ("counters", C.c_void_p), ("n_counters", C.c_int32), ("epilogue", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("post_norm_gamma", C.c_void_p), ("post_norm_out", C.c_void_p), ("post_norm_ld", C.c_int64),
        ("post_norm_eps", C.c_float), ("_pad3", C.c_int32),
        ("prefetch", C.c_void_p), ("prefetch_bytes", C.c_int64),
    ]


class DecodeAttentionArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("out", C.c_void_p),
        ("kv_len_dev", C.c_void_p), ("max_pos", C.c_int64), ("kv_len", C.c_int32),
        ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("n_splits", C.c_int32), ("scale", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class DecodeAttentionFusedArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("inv_freq", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("out", C.c_void_p), ("kv_len_dev", C.c_void_p), ("max_pos", C.c_int64),
        ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32), ("n_splits", C.c_int32),
        ("scale", C.c_float), ("theta", C.c_float), ("n_seqs", C.c_int32), ("_pad", C.c_int32),
        ("ld_qkv", C.c_int64), ("cache_seq_stride", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("prefetch", C.c_void_p), ("prefetch_bytes", C.c_int64), ("ctx_cap", C.c_int64),
    ]


class Qwen2DecoderDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("hidden", C.c_int32), ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("inter", C.c_int32), ("vocab", C.c_int32), ("max_seqs", C.c_int32),
        ("max_ctx", C.c_int64), ("rms_eps", C.c_float), ("rope_theta", C.c_float),
        ("inv_freq", C.c_void_p), ("embed", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p),
        ("qkv_w", C.POINTER(C.c_void_p)), ("qkv_b", C.POINTER(C.c_void_p)), ("o_w", C.POINTER(C.c_void_p)),
        ("gu_w", C.POINTER(C.c_void_p)), ("down_w", C.POINTER(C.c_void_p)), ("ln1", C.POINTER(C.c_void_p)),
        ("ln2", C.POINTER(C.c_void_p)), ("k_cache", C.POINTER(C.c_void_p)), ("v_cache", C.POINTER(C.c_void_p)),
    ]


class Qwen2DecodeStepArgs(C.Structure):
    _fields_ = [
        ("decoder", C.c_void_p), ("rows", C.c_int32), ("n_counters", C.c_int32),
        ("next_id", C.c_void_p), ("kv_len", C.c_void_p), ("counters", C.c_void_p),
        ("logits", C.c_void_p), ("ld_logits", C.c_int64), ("ctx_cap", C.c_int64), ("ctx_sum", C.c_int64),
    ]


class VitEncoderDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("hidden", C.c_int32), ("n_heads", C.c_int32), ("inter", C.c_int32),
        ("image_size", C.c_int32), ("patch", C.c_int32), ("channels", C.c_int32), ("kpad", C.c_int32),
        ("out_dim", C.c_int32), ("ln_eps", C.c_float),
        ("patch_w", C.c_void_p), ("cls", C.c_void_p), ("pos", C.c_void_p), ("pre_g", C.c_void_p), ("pre_b", C.c_void_p),
        ("proj1_w", C.c_void_p), ("proj1_b", C.c_void_p), ("proj2_w", C.c_void_p), ("proj2_b", C.c_void_p),
    ] + [(n, C.POINTER(C.c_void_p)) for n in ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_g", "ln2_b",
                                              "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class AllgatherEmbedArgs(C.Structure):
    _fields_ = [("peer_bufs", C.POINTER(C.c_void_p)), ("rank", C.c_int32), ("world", C.c_int32),
                ("slice_bytes", C.c_int64), ("epoch", C.c_uint32), ("_pad", C.c_int32)]


PRO_NONE, PRO_RMSNORM, PRO_SILU_MUL = 0, 1, 2

_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/streamchat_b200.h declares
SYMBOLS = {
    "scb_version": (C.c_int, []),
    "scb_last_error": (C.c_char_p, []),
    "scb_launch_count": (C.c_uint64, []),
    "scb_cosine_topk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "scb_cosine_topk": (C.c_int, [C.POINTER(CosineTopkArgs), C.c_void_p]),
    "scb_kmeans_step_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "scb_kmeans_step": (C.c_int, [C.POINTER(KmeansStepArgs), C.c_void_p]),
    "scb_kmeans_assign_update": (C.c_int, [C.POINTER(KmeansStepArgs), C.c_void_p]),
    "scb_vit_encoder_create": (C.c_int, [C.POINTER(VitEncoderDesc), C.POINTER(C.c_void_p)]),
    "scb_vit_encoder_destroy": (C.c_int, [C.c_void_p]),
    "scb_vit_encode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "scb_vit_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    "scb_allgather_embed_buffer_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "scb_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "scb_peer_free": (C.c_int, [C.c_void_p]),
    "scb_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "scb_peer_close": (C.c_int, [C.c_void_p]),
    "scb_allgather_embed": (C.c_int, [C.POINTER(AllgatherEmbedArgs), C.c_void_p]),
    "scb_linear": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "scb_linear_splitk": (C.c_int, [C.POINTER(LinearSplitKArgs), C.c_void_p]),
    "scb_linear_splitk_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "scb_attention": (C.c_int, [C.POINTER(AttentionArgs), C.c_void_p]),
    "scb_gemv": (C.c_int, [C.POINTER(GemvArgs), C.c_void_p]),
    "scb_gemv_workspace_bytes": (C.c_size_t, [_i32]),
    "scb_decode_attention_workspace_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "scb_decode_attention": (C.c_int, [C.POINTER(DecodeAttentionArgs), C.c_void_p]),
    "scb_decode_attention_fused_workspace_bytes": (C.c_size_t, [_i32, _i32, _i32, _i32, _i32]),
    "scb_decode_attention_fused": (C.c_int, [C.POINTER(DecodeAttentionFusedArgs), C.c_void_p]),
    "scb_qwen2_decoder_create": (C.c_int, [C.POINTER(Qwen2DecoderDesc), C.POINTER(C.c_void_p)]),
    "scb_qwen2_decoder_destroy": (C.c_int, [C.c_void_p]),
    "scb_qwen2_decoder_max_rows": (C.c_int, []),
    "scb_qwen2_decode_step": (C.c_int, [C.POINTER(Qwen2DecodeStepArgs), C.c_void_p]),
    "scb_qwen2_prefill_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "scb_qwen2_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "scb_qwen2_prefill_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "scb_qwen2_decoder_profile": (C.c_longlong, [C.c_void_p, C.POINTER(C.c_longlong), C.c_longlong]),
    "scb_im2col_patch": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp]),
    "scb_layernorm": (C.c_int, [_vp, _i64, _vp, _vp, _f, _i64, _i32, _vp, _i64, _vp]),
    "scb_rmsnorm": (C.c_int, [_vp, _i64, _vp, _f, _i64, _i32, _vp, _i64, _vp]),
    "scb_vit_embed_layernorm": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _f, _i32, _i32, _i32, _vp, _i64, _vp]),
    "scb_resample_h_u8": (C.c_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "scb_resample_v_normalize": (C.c_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp,
                                           _i32, _vp]),
    "scb_text_embed_layernorm": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _f, _i32, _vp, _i64, _vp]),
    "scb_mean_pool_normalize": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i64, _i32, _vp]),
    "scb_avgpool2x2_tokens": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "scb_silu_mul": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _i64, _vp]),
    "scb_rope_append": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i32, _f, _vp, _vp, _vp, _i64, _vp]),
    "scb_gather_rows": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "scb_gather_rows_rmsnorm": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _vp, _f, _vp, _i64, _vp]),
    "scb_splitk_reduce": (C.c_int, [_vp, _i64, _i64, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _f, _vp]),
    "scb_argmax_rows": (C.c_int, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _i32, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and type every exported symbol. Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScbError(
            f"{LIB_PATH} not found: build it with `make -C streamchat_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().scb_last_error().decode("utf-8", "replace")
        raise ScbError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().scb_launch_count())
