/*
 * streamchat_b200.h — C-ABI of libstreamchat_b200.so (sm_100a only).
 *
 * The reference (hmxiong/StreamChat) is 100 % Python and has no FFI layer; its
 * "operator interface" for the hot path is the Python call surface listed in
 * SURVEY.md §8(b).  Each entry point below replaces the arithmetic that sits
 * under one of those Python calls; the citation next to every function names the
 * reference file:line whose math it takes over.  The Python mirrors of the
 * reference surface (streamchat_b200/*.py and dropin/) bind these symbols with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every buffer is caller-owned
 *     DEVICE memory unless the name ends in _host;
 *   - no allocation, no host synchronisation and no default-stream use inside:
 *     work is enqueued on the cudaStream_t passed in (as void*);
 *   - scratch memory comes from a caller-provided workspace whose size is
 *     returned by the matching *_workspace_bytes() query;
 *   - return 0 on success, a negative scb_status on failure; the message is in
 *     scb_last_error() (thread-local).  Nothing throws, nothing calls exit().
 */
#ifndef STREAMCHAT_B200_H
#define STREAMCHAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum scb_status {
  SCB_OK = 0,
  SCB_ERR_INVALID_ARG = -1,
  SCB_ERR_UNSUPPORTED = -2,
  SCB_ERR_WORKSPACE = -3,
  SCB_ERR_CUDA = -4,
  SCB_ERR_NO_DEVICE = -5
} scb_status;

typedef enum scb_dtype {
  SCB_F32 = 0,
  SCB_F16 = 1
} scb_dtype;

/* Library version (major*10000 + minor*100 + patch). */
int scb_version(void);
/* Last error message of the calling thread ("" if none). */
const char* scb_last_error(void);
/* Number of kernels this library has launched since load (for bench.py's
 * gpu_launches claim). */
uint64_t scb_launch_count(void);

/* ------------------------------------------------------------------------- *
 * cosine top-k retrieval
 * Replaces: sentence_transformers.util.cos_sim + the running arg-max in
 *   utiles.py:732-742,764-772 (memory-tree search, a-R1) and FAISS
 *   IndexFlatL2.search under memory_bank/memory_retrieval/local_doc_qa.py:140,270
 *   (dialogue retrieval, a-R2; on unit vectors L2 ranking == cosine ranking).
 * Computes, for every query q and index row x (rows are NOT pre-normalised):
 *     sim = <x,q> / (max(|x|,1e-12) * max(|q|,1e-12))          (fp32 math)
 * and returns the k largest per query ordered by (sim desc, row asc) — i.e.
 * lowest row index wins ties (hazard H4).
 * ------------------------------------------------------------------------- */
typedef struct scb_cosine_topk_args {
  const void* index;     /* [n_rows, dim] row-major, leading dim ld_index elements */
  const void* queries;   /* [n_queries, dim] row-major, leading dim ld_query       */
  int64_t n_rows;
  int32_t dim;
  int32_t n_queries;
  int64_t ld_index;
  int64_t ld_query;
  int32_t index_dtype;   /* scb_dtype */
  int32_t query_dtype;   /* scb_dtype */
  int32_t k;             /* 1..32 */
  int32_t _pad;
  int64_t* out_indices;  /* [n_queries, k]; -1 where fewer than k rows exist */
  float* out_scores;     /* [n_queries, k]; -inf where fewer than k rows     */
  void* workspace;       /* the first 4*n_queries bytes (block tickets) must be ZERO before the first call; */
  size_t workspace_bytes;/* every call leaves them zero, so one zero-initialised allocation serves all calls */
} scb_cosine_topk_args;

size_t scb_cosine_topk_workspace_bytes(int64_t n_rows, int32_t n_queries, int32_t k);
int scb_cosine_topk(const scb_cosine_topk_args* a, void* stream);

/* ------------------------------------------------------------------------- *
 * k-means (Lloyd) assign + update
 * Replaces: weighted_kmeans_feature utiles.py:291-330 (a-S3) and the
 *   _assign / group_by_label_mean pair of torch_kmeans/clustering/kmeans.py:572-597
 *   + torch_kmeans/utils/utils.py:33-66 (a-S4).
 * One call = one Lloyd iteration over a batch of independent problems that
 * share the same points:
 *   assign:  label[b,n] = argmin_k dist(x[n], c[b,k])   (first minimum wins)
 *            dist = sqrt(sum (x-c)^2)                      if eps == 0   (utiles.py:299)
 *                 = sqrt(sum (x-c+eps)^2)                  torch pairwise_distance (distances.py:134)
 *   update:  new_c[b,k] = sum_{label==k} w[n] x[n] / sum_{label==k} w[n]
 *            clusters with zero weight keep an all-zero row and are reported in
 *            out_wsum so the host can apply the reference's repair rule.
 *   shift:   out_shift[b,k] = |c[b,k] - new_c[b,k]|_2
 * Centroids are fp32; points are fp32 or fp16 (math always fp32 — deviation H3).
 * ------------------------------------------------------------------------- */
typedef struct scb_kmeans_step_args {
  const void* points;        /* [n_points, dim] row-major (contiguous)            */
  const float* weights;      /* [n_points] or NULL (== 1)                         */
  const float* centroids;    /* [n_batch, n_clusters, dim] fp32                   */
  int32_t n_points;
  int32_t n_clusters;        /* 1..64 */
  int32_t n_batch;
  int32_t points_dtype;      /* scb_dtype */
  int64_t dim;
  float dist_eps;            /* 0 or 1e-6 (see above) */
  int32_t do_update;         /* 0: assign only */
  int32_t* out_labels;       /* [n_batch, n_points]                               */
  float* out_dist;           /* [n_batch, n_points] distance to chosen centroid, or NULL */
  float* out_centroids;      /* [n_batch, n_clusters, dim] fp32 (do_update)       */
  float* out_wsum;           /* [n_batch, n_clusters]        (do_update)          */
  float* out_shift;          /* [n_batch, n_clusters]        (do_update)          */
  void* workspace;
  size_t workspace_bytes;
} scb_kmeans_step_args;

size_t scb_kmeans_step_workspace_bytes(int32_t n_points, int32_t n_clusters, int32_t n_batch, int64_t dim);
int scb_kmeans_step(const scb_kmeans_step_args* a, void* stream);

/* ------------------------------------------------------------------------- *
 * Dense linear layer on the 5th-gen tensor cores (tcgen05 + TMEM, TMA-fed)
 * Replaces: every nn.Linear / Conv2d-as-GEMM the reference reaches through HF
 *   transformers — CLIP (longva/model/multimodal_encoder/clip_encoder.py:76),
 *   mm_projector (longva/model/multimodal_projector/builder.py:41-48) and
 *   Qwen2 (longva/model/language_model/llava_qwen.py:155).
 *     out[m,n] = act( sum_k a[m,k] * w[n,k] + bias[n] ) + residual[m,n]
 * a and w are fp16 K-major (PyTorch Linear layout), accumulation fp32 in TMEM.
 * ------------------------------------------------------------------------- */
typedef enum scb_act {
  SCB_ACT_NONE = 0,
  SCB_ACT_QUICK_GELU = 1,  /* x * sigmoid(1.702 x)   — CLIP MLP           */
  SCB_ACT_GELU_ERF = 2,    /* exact GELU             — mm_projector        */
  SCB_ACT_SILU = 3,
  SCB_ACT_SILU_PAIR = 4    /* columns (2i, 2i+1) are (gate_i, up_i): out[m, i] = fp16(silu(fp16 gate)) * fp16 up,
                            * out is [m, n/2]; no bias / residual */
} scb_act;

typedef struct scb_linear_args {
  const void* a;         /* [m, k] fp16, leading dim lda (elements, multiple of 8) */
  const void* w;         /* [n, k] fp16, leading dim ldw (elements, multiple of 8) */
  const void* bias;      /* [n] fp16 or NULL                                       */
  const void* residual;  /* [m, n] fp16 (leading dim ldr) or NULL                  */
  void* out;             /* [m, n], leading dim ldo                                */
  int64_t m, n, k;
  int64_t lda, ldw, ldr, ldo;
  int32_t act;           /* scb_act */
  int32_t out_dtype;     /* scb_dtype: SCB_F16 or SCB_F32 */
  int32_t batch;         /* >=1: independent problems (strided) */
  int32_t _pad;
  int64_t stride_a, stride_w, stride_r, stride_o; /* batch strides in elements */
  float alpha;           /* scale applied to the accumulator before bias (1.0 default) */
  int32_t _pad2;
} scb_linear_args;

int scb_linear(const scb_linear_args* a, void* stream);


/* ------------------------------------------------------------------------- *
 * Row-wise glue kernels (fp16 I/O, fp32 statistics).  Operator order follows the
 * third-party math the reference executes (HF transformers).
 * ------------------------------------------------------------------------- */

/* CLIP patch-embedding im2col (Conv2d k=s=patch, no bias; modeling_clip.py:148-154 reached
 * from longva/model/multimodal_encoder/clip_encoder.py:76).  pixels [n,c,h,w] fp16 ->
 * out [n*(h/p)*(w/p), ld_out] fp16, columns c*p*p.. zero-padded. */
int scb_im2col_patch(const void* pixels, void* out, int32_t n_img, int32_t channels, int32_t height,
                     int32_t width, int32_t patch, int64_t ld_out, void* stream);

/* nn.LayerNorm over the last dim (CLIP layer_norm1/2, modeling_clip.py:354-386). */
int scb_layernorm(const void* x, int64_t ld_x, const void* gamma, const void* beta, float eps,
                  int64_t rows, int32_t dim, void* out, int64_t ld_out, void* stream);

/* Qwen2RMSNorm (modeling_qwen2.py:249-263): fp32 normalise, round to fp16, then * weight. */
int scb_rmsnorm(const void* x, int64_t ld_x, const void* weight, float eps, int64_t rows,
                int32_t dim, void* out, int64_t ld_out, void* stream);

/* CLIP token assembly + pre_layrnorm (modeling_clip.py:209-217,677): row (n,t) =
 * LN((t==0 ? cls : patch[n*(T-1)+t-1]) + pos[t]). */
int scb_vit_embed_layernorm(const void* patch_embeds, int64_t ld_patch, const void* cls_embedding,
                            const void* pos_embedding, const void* gamma, const void* beta, float eps,
                            int32_t n_img, int32_t tokens_per_img, int32_t dim, void* out,
                            int64_t ld_out, void* stream);

/* ---- frame ingest (a-V1): `process_images` utiles.py:71-87 = CLIPImageProcessor.preprocess on the
 * BGR->RGB-swapped cv2 frame (inference_streaming_longva_v2.py:507-511).  Pillow's 8-bit bicubic resample is
 * two passes of fixed-point convolutions (libImaging/Resample.c: 22 fractional bits, every pass rounded to
 * uint8); bounds[o] = (first input index, tap count), coeffs[o][0..ksize) int32, both for the output range
 * that survives the centre crop.
 * Horizontal pass: src uint8 [n, rows, W, 3] (strides in bytes) -> dst uint8 [n, rows, out_w, 3]. */
int scb_resample_h_u8(const uint8_t* src, int64_t n_img, int64_t img_stride, int64_t row_stride, int32_t rows,
                      const int32_t* bounds, const int32_t* coeffs, int32_t ksize, int32_t out_w, uint8_t* dst,
                      void* stream);

/* Vertical pass (bounds/coeffs NULL: no resample, rows start at row_offset) + crop + value table + CHW store:
 * out[n, c, y, x] = lut[c][ pixel(n, y, x + x_offset)[swap_rb ? 2-c : c] ], lut float [3,256] holding
 * ((v * 1/255) - mean_c) / std_c; out fp16 or fp32 [n, 3, out_size, out_size]. */
int scb_resample_v_normalize(const uint8_t* src, int64_t n_img, int64_t img_stride, int64_t row_stride,
                             int32_t x_offset, const int32_t* bounds, const int32_t* coeffs, int32_t ksize,
                             int32_t row_offset, int32_t out_size, const float* lut, int32_t swap_rb, void* out,
                             int32_t out_dtype, void* stream);

/* BertEmbeddings.forward of the retrieval text encoder (the reference embeds the query and every node text
 * with `model(**ids).last_hidden_state[:, 0]`, utiles.py:704-708,723-726,757-760; HF modeling_bert.py
 * BertEmbeddings): row r = LayerNorm(word[ids[r]] + token_type[0] + position[position_ids[r]]).
 * ids / position_ids are device int64 arrays; tables fp16; out fp16 [n_tokens, dim]. */
int scb_text_embed_layernorm(const int64_t* ids, const int64_t* position_ids, int64_t n_tokens,
                             const void* word_embeddings, const void* position_embeddings,
                             const void* token_type_row, const void* gamma, const void* beta, float eps,
                             int32_t dim, void* out, int64_t ld_out, void* stream);

/* Sentence-embedding tail of the dialogue-memory encoder (memory_bank/memory_retrieval/local_doc_qa.py:185-194:
 * langchain HuggingFaceEmbeddings = sentence-transformers all-MiniLM-L6-v2 = BertModel -> Pooling(mean over the
 * attention mask) -> Normalize(p=2)).  hidden: packed last_hidden_state rows fp16; text t owns rows
 * [offsets[t], offsets[t+1]) (device int32 [n_texts + 1]); out[t] = mean of its rows (fp32 sums, divided by
 * max(count, 1e-9)), then divided by max(|.|_2, 1e-12) when normalize != 0.  out fp32 [n_texts, ld_out]. */
int scb_mean_pool_normalize(const void* hidden, int64_t ld_hidden, const int32_t* offsets, int32_t n_texts,
                            int32_t dim, float* out, int64_t ld_out, int32_t normalize, void* stream);

/* SpatialPool average, stride 2 (longva/model/multimodal_resampler/spatial_pool.py:23-32):
 * in rows [n*img_stride_rows + tok_offset + y*grid + x] -> out rows [n*(grid/2)^2 + ...]. */
int scb_avgpool2x2_tokens(const void* in, int64_t ld_in, int64_t img_stride_rows, int32_t tok_offset,
                          int32_t grid, int32_t n_img, int32_t dim, void* out, int64_t ld_out,
                          void* stream);

/* Qwen2 MLP gate (modeling_qwen2.py:41-48): out[m,i] = silu(gu[m,i]) * gu[m,inter+i]. */
int scb_silu_mul(const void* gate_up, int64_t ld_gu, int64_t rows, int32_t inter, void* out,
                 int64_t ld_out, void* stream);

/* Qwen2 rotary embedding, NeoX half-split (modeling_qwen2.py:116-146), applied in place to the
 * q and k slices of fused qkv rows [n_q*hd | n_kv*hd | n_kv*hd]; k (rotated) and v are appended
 * to the layer's KV cache [n_kv][max_pos][hd] when the cache pointers are non-NULL.
 * positions == NULL -> pos_base + row.  inv_freq == NULL -> theta^(-2i/hd). */
int scb_rope_append(void* qkv, int64_t ld_qkv, int64_t rows, int32_t n_q_heads, int32_t n_kv_heads,
                    int32_t head_dim, const int32_t* positions, int32_t pos_base, float theta,
                    const float* inv_freq, void* k_cache, void* v_cache, int64_t max_pos,
                    void* stream);

/* embed_tokens lookup (inference_streaming_longva_v2.py:177, llava_arch.py:264). ids<0 -> zeros. */
int scb_gather_rows(const void* table, const int64_t* ids, int64_t n_ids, int32_t dim, void* out,
                    int64_t ld_out, void* stream);

/* Decode-step entry: embed_tokens lookup fused with the first layer's input RMSNorm
 * (modeling_qwen2.py:249-263): out[r] = table[ids[r]], norm_out[r] = RMSNorm(out[r]) * gamma. */
int scb_gather_rows_rmsnorm(const void* table, const int64_t* ids, int64_t n_ids, int32_t dim, void* out,
                            int64_t ld_out, const void* gamma, float eps, void* norm_out,
                            int64_t ld_norm, void* stream);

/* Split-K linear with the reduction fused into the GEMM launch (decode steps of wide batches: o_proj / down_proj / qkv of
 * nn.Linear, modeling_qwen2.py:74-88,229-243, run for a handful of rows, where N alone gives too few tiles to pull the
 * weights at HBM speed): out[r,n] = fp16(sum_s a[r, s*ks:(s+1)*ks] . w[n, s*ks:(s+1)*ks] + bias[n] + residual[r,n]),
 * slices summed in index order, then optionally post_norm_out[r,:] = RMSNorm(out[r,:]) * gamma (modeling_qwen2.py:249-263).
 * One CTA per (n tile of 128, K slice): ceil(m/128) * ceil(n/128) * n_splits must not exceed the SM count
 * (SCB_ERR_UNSUPPORTED otherwise: run scb_linear with fp32 partials + scb_splitk_reduce).  `partial` holds
 * n_splits*m*n floats of scratch; `workspace` (scb_linear_splitk_workspace_bytes) must be zero before the FIRST call and
 * is left zero by every call, so one buffer serves a whole stream-ordered launch chain. */
typedef struct scb_linear_splitk_args {
  const void* a;         /* [m, k] fp16, leading dim lda (multiple of 8) */
  const void* w;         /* [n, k] fp16, leading dim ldw (multiple of 8) */
  const void* bias;      /* [n] fp16 or NULL */
  const void* residual;  /* [m, n] fp16, leading dim ldr, or NULL */
  void* partial;         /* fp32 scratch, n_splits*m*n elements */
  void* out;             /* [m, n] fp16, leading dim ldo */
  const void* post_norm_gamma; /* [n] fp16 or NULL */
  void* post_norm_out;   /* [m, n] fp16, leading dim post_norm_ld, or NULL */
  void* workspace;
  int64_t workspace_bytes;
  int64_t m, n, k;
  int64_t lda, ldw, ldr, ldo, post_norm_ld;
  int32_t n_splits;
  float post_norm_eps;
} scb_linear_splitk_args;

int64_t scb_linear_splitk_workspace_bytes(int64_t m, int64_t n);
int scb_linear_splitk(const scb_linear_splitk_args* a, void* stream);

/* Tail of a split-K decode layer (wide batches run o_proj / down_proj / qkv as S independent K-slices
 * through scb_linear with fp32 partial outputs): out[r,n] = fp16(sum_s partial[s,r,n] + bias[n] +
 * residual[r,n]) summed in split order, then optionally post_norm_out[r,:] = RMSNorm(out[r,:]) * gamma. */
int scb_splitk_reduce(const float* partial, int64_t split_stride, int64_t ld_partial, int32_t n_splits,
                      int64_t rows, int32_t n, const void* bias, const void* residual, int64_t ld_res,
                      void* out, int64_t ld_out, const void* post_norm_gamma, void* post_norm_out,
                      int64_t post_norm_ld, float post_norm_eps, void* stream);

/* Greedy token choice: arg-max per row of logits (fp16 or fp32), lowest index wins ties —
 * torch.argmax over the fp16 lm_head output as HF generate does.  If counters != NULL the
 * n_counters device int32 values it points to are each incremented by one (position / length
 * bookkeeping of a replayed CUDA graph). */
int scb_argmax_rows(const void* logits, int32_t dtype, int64_t ld, int64_t rows, int32_t n,
                    int64_t* out, int32_t* counters, int32_t n_counters, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused attention softmax(scale * Q K^T [+ causal mask]) V on tcgen05 tensor cores.
 * Replaces the SDPA reached through HF transformers: CLIP self-attention
 * (modeling_clip.py:282-337 via clip_encoder.py:76) and Qwen2 prefill attention with
 * GQA (modeling_qwen2.py:161-184 via llava_qwen.py:155).
 * Element (b, head, token, :) of q/k/v/out lives at base + b*stride_b + token*stride_t
 * + head*stride_h with head_dim contiguous fp16 values; strides in elements, multiples
 * of 8.  causal: query t sees keys s <= t + (kv_len - q_len).
 * ------------------------------------------------------------------------- */
typedef struct scb_attention_args {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int32_t batch, n_q_heads, n_kv_heads, head_dim;   /* head_dim: 64 or 128 */
  int64_t q_len, kv_len;
  int64_t q_stride_b, q_stride_t, q_stride_h;
  int64_t k_stride_b, k_stride_t, k_stride_h;
  int64_t v_stride_b, v_stride_t, v_stride_h;
  int64_t o_stride_b, o_stride_t, o_stride_h;
  float scale;
  int32_t causal;
} scb_attention_args;

int scb_attention(const scb_attention_args* a, void* stream);

/* ------------------------------------------------------------------------- *
 * Decode-step kernels (cached generation, HBM-bound).  Replace the batch-1 Linear and
 * SDPA-over-DynamicCache calls of Qwen2ForCausalLM.generate reached from
 * longva/model/language_model/llava_qwen.py:155 (use_cache forced on:
 * generation/utils.py:2425-2428).
 * ------------------------------------------------------------------------- */
typedef enum scb_gemv_prologue {
  SCB_PRO_NONE = 0,
  SCB_PRO_RMSNORM = 1,   /* a' = fp16(a * rsqrt(mean(a^2)+eps)) * gamma   (modeling_qwen2.py:249-263) */
  SCB_PRO_SILU_MUL = 2   /* a' = fp16(silu(a[:k])) * a[k:2k]              (modeling_qwen2.py:41-48)   */
} scb_gemv_prologue;

typedef struct scb_gemv_args {
  const void* a;          /* [rows, lda] fp16 activations (2k wide for SCB_PRO_SILU_MUL) */
  const void* w;          /* [n, ldw] fp16 weights, k contiguous */
  const void* bias;       /* [n] fp16 or NULL */
  const void* residual;   /* [rows, ldr] fp16 or NULL */
  const void* gamma;      /* [k] fp16 RMSNorm weight (SCB_PRO_RMSNORM) */
  void* out;              /* [rows, ldo] */
  int64_t n, k;
  int64_t lda, ldw, ldr, ldo;
  int32_t rows;           /* 1..16 (k % 32 == 0, no prologue), else 1, 2 or 4 */
  int32_t prologue;       /* scb_gemv_prologue */
  int32_t out_dtype;      /* scb_dtype */
  float eps;
  /* optional fused greedy choice (lm_head): argmax_out[r] = arg-max over n of the STORED output
   * row r (lowest index wins ties), then counters[0..n_counters) += 1.  Needs the workspace, whose
   * first 256 bytes (the ticket) must be zero before the first call and are left zero. */
  int64_t* argmax_out;
  int32_t* counters;
  int32_t n_counters;
  int32_t epilogue;       /* 0: none.  1: SiLU-pair — rows (2i, 2i+1) of w are (gate_i, up_i) and
                           * out[r, i] = fp16(silu(fp16 gate)) * fp16 up, out is [rows, n/2] fp16
                           * (Qwen2 MLP gate, modeling_qwen2.py:41-48, fused into the gate|up layer) */
  void* workspace;
  size_t workspace_bytes;
  /* optional (k % 32 == 0, no prologue, fp16 out): the last block to finish writes
   * post_norm_out[r, :] = RMSNorm(out[r, :]) * post_norm_gamma (the next layer's input norm) */
  const void* post_norm_gamma;
  void* post_norm_out;
  int64_t post_norm_ld;
  float post_norm_eps;
  int32_t _pad3;
  /* optional: blocks that run out of work prefetch [prefetch, prefetch + prefetch_bytes) into L2
   * (the next kernel's weights) */
  const void* prefetch;
  int64_t prefetch_bytes;
} scb_gemv_args;

size_t scb_gemv_workspace_bytes(int32_t rows);
int scb_gemv(const scb_gemv_args* a, void* stream);

typedef struct scb_decode_attention_args {
  const void* q;            /* [n_q_heads, head_dim] fp16 (already rotated) */
  const void* k_cache;      /* [n_kv_heads, max_pos, head_dim] fp16 */
  const void* v_cache;      /* [n_kv_heads, max_pos, head_dim] fp16 */
  void* out;                /* [n_q_heads, head_dim] fp16 */
  const int32_t* kv_len_dev;/* device scalar with the context length, or NULL -> kv_len */
  int64_t max_pos;
  int32_t kv_len;
  int32_t n_q_heads, n_kv_heads, head_dim;
  int32_t n_splits;
  float scale;
  void* workspace;
  size_t workspace_bytes;
} scb_decode_attention_args;

size_t scb_decode_attention_workspace_bytes(int32_t n_q_heads, int32_t head_dim, int32_t n_splits);
int scb_decode_attention(const scb_decode_attention_args* a, void* stream);

/* One launch for the attention block of a decode step: RoPE of the new q/k (same rounding as
 * scb_rope_append), append of k/v at position kv_len-1, split-KV attention for every q head and
 * the cross-split combine — for n_seqs independent sequences (batched chunk captions): sequence
 * s uses fused row qkv + s*ld_qkv ([n_q*hd | n_kv*hd | n_kv*hd], un-rotated), caches
 * k_cache + s*cache_seq_stride, length kv_len_dev[s] (INCLUDING the token being decoded) and
 * writes out[s].  The first 256 bytes of the workspace hold the tickets and must be zero before
 * the first call (they are left zero). */
typedef struct scb_decode_attention_fused_args {
  const void* qkv;
  const float* inv_freq;     /* [head_dim/2] fp32 or NULL -> theta^(-2i/hd) */
  void* k_cache;             /* [n_kv_heads, max_pos, head_dim] fp16 */
  void* v_cache;
  void* out;                 /* [n_seqs, n_q_heads, head_dim] fp16 */
  const int32_t* kv_len_dev; /* [n_seqs] */
  int64_t max_pos;
  int32_t n_q_heads, n_kv_heads, head_dim, n_splits;
  float scale;
  float theta;
  int32_t n_seqs;            /* >= 1, n_seqs * n_kv_heads <= 256 */
  int32_t _pad;
  int64_t ld_qkv;            /* elements between the sequences' qkv rows */
  int64_t cache_seq_stride;  /* elements between the sequences' caches */
  void* workspace;
  size_t workspace_bytes;
  const void* prefetch;      /* optional: pulled into L2 while the attention runs (next layer weights) */
  int64_t prefetch_bytes;
  int64_t ctx_cap;           /* optional upper bound of every kv_len (0 = max_pos): sizes the per-split
                              * score buffer in shared memory */
} scb_decode_attention_fused_args;

size_t scb_decode_attention_fused_workspace_bytes(int32_t n_seqs, int32_t n_q_heads,
                                                  int32_t n_kv_heads, int32_t head_dim,
                                                  int32_t n_splits);
int scb_decode_attention_fused(const scb_decode_attention_fused_args* a, void* stream);

/* ------------------------------------------------------------------------- *
 * scb_qwen2_decoder_* — ONE persistent kernel per cached decode step (1..4 sequences in
 * lock-step).  Replaces the whole per-token body of Qwen2ForCausalLM.generate reached from
 * longva/model/language_model/llava_qwen.py:155 (transformers modeling_qwen2.py:269-311 per
 * layer, :474-475 lm_head on the last position, greedy arg-max of generation/utils.py): token
 * embedding gather, per layer RMSNorm -> qkv (+bias) -> RoPE + KV append + attention over the
 * cache -> o_proj (+residual) -> RMSNorm -> gate|up (SiLU gate) -> down (+residual), final
 * RMSNorm, lm_head, arg-max (lowest index wins), position/length bookkeeping.
 *
 * One CTA per SM stays resident for the step.  A producer warp streams every weight matrix
 * through a shared-memory ring with TMA (16 rows x 512 k tiles, 128-byte swizzle) and runs ahead
 * of the layer dependencies — weights depend on nothing — so HBM keeps streaming across the
 * grid-wide barriers that separate the stages; eight consumer warps feed the tiles to
 * mma.sync.m16n8k16 (weights = A operand, activations = B operand from shared memory) and own the
 * attention stage.  HBM-bound: algorithmic bytes per step = all weights + the live KV cache.
 *
 * The handle owns device memory allocated at creation (tensor maps, per-layer pointer table,
 * a few MB of workspace); weights and caches stay caller-owned and must outlive the handle.
 * Weight layouts: PyTorch Linear [n, k] fp16, k contiguous, k % 64 == 0, n % 16 == 0;
 * gu_w is the row-interleaved gate|up matrix [2*inter, hidden] (row 2i = gate_i, 2i+1 = up_i).
 * ------------------------------------------------------------------------- */
typedef struct scb_qwen2_decoder_desc {
  int32_t n_layers, hidden, n_q_heads, n_kv_heads, head_dim, inter, vocab;
  int32_t max_seqs;          /* cache slots: k/v caches are [max_seqs, n_kv_heads, max_ctx, head_dim] */
  int64_t max_ctx;
  float rms_eps, rope_theta;
  const float* inv_freq;     /* device [head_dim/2] fp32 or NULL -> theta^(-2i/hd) */
  const void* embed;         /* device [vocab, hidden] fp16 */
  const void* final_norm;    /* device [hidden] fp16 */
  const void* lm_head;       /* device [vocab, hidden] fp16 */
  /* HOST arrays of n_layers device pointers */
  const void* const* qkv_w;  /* [ (n_q+2*n_kv)*hd, hidden ] */
  const void* const* qkv_b;  /* [ (n_q+2*n_kv)*hd ] or NULL entries */
  const void* const* o_w;    /* [ hidden, n_q*hd ] */
  const void* const* gu_w;   /* [ 2*inter, hidden ] interleaved */
  const void* const* down_w; /* [ hidden, inter ] */
  const void* const* ln1;    /* [hidden] input RMSNorm weight */
  const void* const* ln2;    /* [hidden] post-attention RMSNorm weight */
  void* const* k_cache;
  void* const* v_cache;
} scb_qwen2_decoder_desc;

typedef struct scb_qwen2_decode_step_args {
  void* decoder;             /* handle from scb_qwen2_decoder_create */
  int32_t rows;              /* 1..scb_qwen2_decoder_max_rows(): sequences decoding in lock-step (cache slots 0..rows-1) */
  int32_t n_counters;
  int64_t* next_id;          /* device [rows] in: token to feed, out: greedy next token */
  const int32_t* kv_len;     /* device [rows]: context length INCLUDING the token being decoded */
  int32_t* counters;         /* optional device ints incremented by one at the end of the step */
  void* logits;              /* device [rows, ld_logits] fp16, written */
  int64_t ld_logits;
  int64_t ctx_cap;           /* upper bound of every kv_len (0 = max_ctx): sizes the score buffer */
  int64_t ctx_sum;           /* optional upper bound of the SUM of the rows' kv_len (0 = rows * ctx_cap): a tighter
                              * score buffer when a long context decodes next to short ones */
} scb_qwen2_decode_step_args;

int scb_qwen2_decoder_create(const scb_qwen2_decoder_desc* d, void** out_handle);
int scb_qwen2_decoder_destroy(void* handle);
int scb_qwen2_decoder_max_rows(void);
int scb_qwen2_decode_step(const scb_qwen2_decode_step_args* a, void* stream);
/* Long-context prefill of ONE sequence as a single native call over the decoder handle: appends n_tokens
 * rows of inputs_embeds ([n_tokens, hidden] fp16) to cache slot `slot` at position pos0 (per layer RMSNorm ->
 * qkv -> RoPE + KV append -> fused causal attention over the cache -> o_proj -> RMSNorm -> gate|up -> down) and
 * writes the fp16 lm_head row of the LAST position to logits_row [vocab]  (modeling_qwen2.py:269-311, :474-475
 * reached from llava_qwen.py:155).  Same kernels, order and rounding points as the per-launch path
 * (scb_rmsnorm / scb_linear / scb_rope_append / scb_attention / scb_gemv).  The first
 * scb_gemv_workspace_bytes(1) bytes of the workspace must be zero before the first call (they are left zero). */
size_t scb_qwen2_prefill_workspace_bytes(void* decoder, int64_t n_tokens);
int scb_qwen2_prefill(void* decoder, const void* inputs_embeds, int64_t n_tokens, int32_t slot, int64_t pos0,
                      void* logits_row, void* workspace, size_t workspace_bytes, void* stream);
/* The same for several sequences in ONE pass: inputs_embeds stacks the prompts' rows (sequence after sequence,
 * [sum(n_tokens), hidden]); n_tokens / slots / pos0 are HOST arrays of n_seqs entries (distinct cache slots); sequence
 * q's last-position logits go to logits_rows + q * ld_logits.  Linear layers and norms run once over the stack (a lone
 * 5.8 k-token prompt leaves the last tile wave of the o_proj / down_proj GEMMs two-thirds empty), RoPE + KV append +
 * causal attention per sequence.  Bit-identical to n_seqs one-sequence calls; workspace =
 * scb_qwen2_prefill_workspace_bytes(decoder, sum(n_tokens)). */
int scb_qwen2_prefill_batch(void* decoder, const void* inputs_embeds, int32_t n_seqs, const int64_t* n_tokens_host,
                            const int32_t* slots_host, const int64_t* pos0_host, void* logits_rows, int64_t ld_logits,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Debug aid: per-CTA clock64 stamps of the last step, 16 per layer (layer start, then after every
 * prologue / stage / grid barrier).  The first call arms the recording and returns 0; later calls
 * copy [n_sm][n_layers][16] int64 into host_out (capacity in elements) and return the count; -1 on error. */
long long scb_qwen2_decoder_profile(void* decoder, long long* host_out, long long capacity);

/* ------------------------------------------------------------------------- *
 * scb_vit_encode — `model.encode_images(images)` as ONE native call.
 * Replaces: LlavaMetaForCausalLM.encode_images longva/model/llava_arch.py:179-184 =
 *   CLIPVisionTower.forward / feature_select  longva/model/multimodal_encoder/clip_encoder.py:46-79
 *   (HF CLIPVisionModel, hidden_states[select_layer] without the CLS token), mm_projector mlp2x_gelu
 *   longva/model/multimodal_projector/builder.py:41-48 and SpatialPool (2x2 average)
 *   longva/model/multimodal_resampler/spatial_pool.py:23-32.
 * The handle holds host copies of the weight pointer tables (weights stay caller-owned device
 * memory, fp16, PyTorch Linear layout; qkv_w / qkv_b are the q|k|v projections concatenated along
 * the output dim; patch_w is the Conv2d kernel flattened to [hidden, kpad], columns beyond
 * channels*patch^2 zero).  One call enqueues the whole pass for n_frames frames into the
 * caller's workspace: no allocation, no synchronisation, CUDA-graph capturable.
 * n_layers = the layers that feed hidden_states[select_layer] (23 of 24 for ViT-L/14, select -2).
 * ------------------------------------------------------------------------- */
typedef struct scb_vit_encoder_desc {
  int32_t n_layers, hidden, n_heads, inter;
  int32_t image_size, patch, channels, kpad;
  int32_t out_dim;           /* projector / LLM width */
  float ln_eps;
  const void* patch_w;       /* [hidden, kpad] */
  const void* cls;           /* [hidden] */
  const void* pos;           /* [tokens, hidden] */
  const void* pre_g;         /* pre_layrnorm weight / bias [hidden] */
  const void* pre_b;
  const void* proj1_w;       /* [out_dim, hidden] + bias */
  const void* proj1_b;
  const void* proj2_w;       /* [out_dim, out_dim] + bias */
  const void* proj2_b;
  /* HOST arrays of n_layers device pointers */
  const void* const* ln1_g;
  const void* const* ln1_b;
  const void* const* qkv_w;  /* [3*hidden, hidden] */
  const void* const* qkv_b;  /* [3*hidden] */
  const void* const* out_w;  /* [hidden, hidden] */
  const void* const* out_b;
  const void* const* ln2_g;
  const void* const* ln2_b;
  const void* const* fc1_w;  /* [inter, hidden] */
  const void* const* fc1_b;
  const void* const* fc2_w;  /* [hidden, inter] */
  const void* const* fc2_b;
} scb_vit_encoder_desc;

int scb_vit_encoder_create(const scb_vit_encoder_desc* d, void** out_handle);
int scb_vit_encoder_destroy(void* handle);
size_t scb_vit_encode_workspace_bytes(void* handle, int32_t n_frames);
/* pixels [n_frames, channels, S, S] fp16 (CLIP-normalised) -> out [n_frames, (S/patch/2)^2, out_dim] fp16.
 * tower_out (optional, may be NULL): hidden_states[select_layer] [n_frames * tokens, hidden] fp16, CLS rows included. */
int scb_vit_encode(void* handle, const void* pixels, int32_t n_frames, void* out, void* tower_out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * scb_allgather_embed — the ONE exchange step of the frame-sharded stream: every rank's
 * memory embeddings (pooled frame features = depth-0 node centroids, [chunks_per_rank * 40*144, H] fp16) gathered
 * IN PLACE into the buffer the consumers read (k-means merges, the answer's visual context, cosine top-k) by a
 * push over NVLink / NVSwitch peer memory.  No reference counterpart: the reference runs two replicas serially
 * (inference_streaming_longva_v2.py:697,700,879).
 *
 * Buffers: every rank allocates its gather buffer with scb_peer_alloc(scb_allgather_embed_buffer_bytes(...)) —
 * cudaMalloc memory laid out as [world][slice_bytes] followed by a 256-byte flag block, zero-initialised — sends
 * the 64-byte IPC handle to its peers over any control plane (torch.distributed in this repo) and maps theirs with
 * scb_peer_open.  Rank r writes its own contribution at offset r*slice_bytes of ITS buffer, then every rank calls
 * scb_allgather_embed once with the same, strictly increasing epoch: a copy kernel stores the rank's slice into
 * every peer's buffer, publishes the epoch in the peers' flag blocks (system-scope release) and waits until every
 * peer's epoch arrived in its own; when the kernel has finished on the caller's stream the whole buffer is valid.
 * The caller separates two gathers by any inter-rank synchronisation that proves the consumers of the first are
 * done (in this repo: the collective that ends the answer stage).
 * ------------------------------------------------------------------------- */
typedef struct scb_allgather_embed_args {
  void* const* peer_bufs;    /* HOST array [world]: every rank's buffer as mapped in THIS process (own at [rank]) */
  int32_t rank, world;       /* world <= 16 */
  int64_t slice_bytes;       /* bytes per rank, multiple of 16 */
  uint32_t epoch;            /* >= 1, strictly increasing over the calls on one set of buffers; 0: the call number is kept
                              * on the device (in the flag block), so the launch can be replayed from a CUDA graph — do
                              * not mix both modes on one set of buffers */
  int32_t _pad;
} scb_allgather_embed_args;

size_t scb_allgather_embed_buffer_bytes(int64_t slice_bytes, int32_t world);
int scb_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out /* 64 bytes */);
int scb_peer_free(void* dev_ptr);
int scb_peer_open(const void* ipc_handle /* 64 bytes */, void** mapped_ptr);
int scb_peer_close(void* mapped_ptr);
int scb_allgather_embed(const scb_allgather_embed_args* a, void* stream);

/* scb_kmeans_assign_update — the name SURVEY.md §8(b) lists for the Lloyd assign + update step: identical to
 * scb_kmeans_step (same argument struct, same workspace query). */
int scb_kmeans_assign_update(const scb_kmeans_step_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STREAMCHAT_B200_H */
