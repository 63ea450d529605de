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
  void* workspace;
  size_t workspace_bytes;
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
  SCB_ACT_SILU = 3
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

#ifdef __cplusplus
}
#endif
#endif /* STREAMCHAT_B200_H */
