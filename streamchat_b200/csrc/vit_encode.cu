// vit_encode.cu — scb_vit_encode: the whole `encode_images` of the reference (longva/model/llava_arch.py:179-184:
// CLIP vision tower -> mm_projector -> 2x2 spatial pool) as ONE native call over a handle that holds the weight
// pointer tables.  It issues, from C++, the launch sequence of the per-op path (im2col -> patch GEMM -> token
// assembly + pre-LN -> L x { LN, fused QKV GEMM, tcgen05 attention, out-proj GEMM (+residual), LN, fc1 GEMM
// (+quick-GELU), fc2 GEMM (+residual) } -> projector GEMM-1 (+GELU-erf) -> 2x2 mean -> projector GEMM-2) into a
// caller-provided workspace: no allocation, no host synchronisation, capturable in a CUDA graph.  Same kernels,
// order and rounding points as streamchat_b200/vision.py (bit-identical; tests/test_gpu_native_entry_points.py).
#include "scb_common.cuh"
#include <new>
#include <string.h>
#include <math.h>

namespace {

constexpr uint32_t kVitMagic = 0x56495445u;  // "VITE"

struct VitLayer {
  const void *ln1_g, *ln1_b, *qkv_w, *qkv_b, *out_w, *out_b, *ln2_g, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};

struct VitHandle {
  uint32_t magic;
  scb_vit_encoder_desc d;   // pointer tables inside are NOT valid after create(): use `layers`
  VitLayer* layers;
  int grid, tokens, pooled;
};

struct VitBuffers {
  char *cols, *patch, *xa, *xb, *y, *wide, *fc1, *ctx, *pool;
  size_t total;
};

// rows = n_frames * tokens.  `wide` holds qkv [rows, 3D] inside the tower and projector-1's output [rows, H] after it.
VitBuffers carve(const VitHandle* h, int64_t n, char* base) {
  const scb_vit_encoder_desc& d = h->d;
  const size_t rows = (size_t)n * h->tokens, prow = (size_t)n * (h->tokens - 1);
  const size_t D = d.hidden, I = d.inter, H = d.out_dim;
  const size_t wide_cols = 3 * D > H ? 3 * D : H;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += scb_align_up(bytes, 256);
    return p;
  };
  VitBuffers b;
  b.cols = take(prow * d.kpad * 2);
  b.patch = take(prow * D * 2);
  b.xa = take(rows * D * 2);
  b.xb = take(rows * D * 2);
  b.y = take(rows * D * 2);
  b.ctx = take(rows * D * 2);
  b.wide = take(rows * wide_cols * 2);
  b.fc1 = take(rows * I * 2);
  b.pool = take((size_t)n * h->pooled * H * 2);
  b.total = off;
  return b;
}

}  // namespace

extern "C" {

int scb_vit_encoder_create(const scb_vit_encoder_desc* d, void** out_handle) {
  SCB_REQUIRE(d != nullptr && out_handle != nullptr, SCB_ERR_INVALID_ARG, "scb_vit_encoder_create: NULL argument");
  *out_handle = nullptr;
  SCB_REQUIRE(d->n_layers >= 0 && d->hidden >= 64 && d->n_heads >= 1 && d->inter >= 8 && d->out_dim >= 8 &&
                  d->patch >= 1 && d->image_size >= 2 * d->patch && d->channels >= 1,
              SCB_ERR_INVALID_ARG, "scb_vit_encoder_create: bad dimensions");
  SCB_REQUIRE(d->hidden % d->n_heads == 0 && (d->hidden / d->n_heads == 64 || d->hidden / d->n_heads == 128),
              SCB_ERR_UNSUPPORTED, "scb_vit_encoder_create: head_dim %d unsupported (64 or 128)",
              d->n_heads ? d->hidden / d->n_heads : 0);
  const int grid = d->image_size / d->patch;
  SCB_REQUIRE(d->image_size % d->patch == 0 && grid % 2 == 0, SCB_ERR_UNSUPPORTED,
              "scb_vit_encoder_create: image_size must be a multiple of patch with an even token grid");
  SCB_REQUIRE(d->kpad >= d->channels * d->patch * d->patch && d->kpad % 8 == 0, SCB_ERR_INVALID_ARG,
              "scb_vit_encoder_create: kpad must be >= channels*patch^2 and a multiple of 8");
  SCB_REQUIRE(d->hidden % 8 == 0 && d->inter % 8 == 0 && d->out_dim % 8 == 0, SCB_ERR_UNSUPPORTED,
              "scb_vit_encoder_create: widths must be multiples of 8");
  SCB_REQUIRE(d->patch_w && d->cls && d->pos && d->pre_g && d->pre_b && d->proj1_w && d->proj1_b && d->proj2_w &&
                  d->proj2_b,
              SCB_ERR_INVALID_ARG, "scb_vit_encoder_create: NULL weight pointer");
  SCB_REQUIRE(d->n_layers == 0 || (d->ln1_g && d->ln1_b && d->qkv_w && d->qkv_b && d->out_w && d->out_b && d->ln2_g &&
                                   d->ln2_b && d->fc1_w && d->fc1_b && d->fc2_w && d->fc2_b),
              SCB_ERR_INVALID_ARG, "scb_vit_encoder_create: NULL pointer table");
  VitHandle* h = new (std::nothrow) VitHandle();
  VitLayer* layers = new (std::nothrow) VitLayer[d->n_layers > 0 ? d->n_layers : 1];
  if (!h || !layers) {
    delete h;
    delete[] layers;
    SCB_REQUIRE(false, SCB_ERR_CUDA, "scb_vit_encoder_create: out of host memory");
  }
  for (int l = 0; l < d->n_layers; ++l) {
    VitLayer& y = layers[l];
    y.ln1_g = d->ln1_g[l]; y.ln1_b = d->ln1_b[l]; y.qkv_w = d->qkv_w[l]; y.qkv_b = d->qkv_b[l];
    y.out_w = d->out_w[l]; y.out_b = d->out_b[l]; y.ln2_g = d->ln2_g[l]; y.ln2_b = d->ln2_b[l];
    y.fc1_w = d->fc1_w[l]; y.fc1_b = d->fc1_b[l]; y.fc2_w = d->fc2_w[l]; y.fc2_b = d->fc2_b[l];
    if (!y.ln1_g || !y.ln1_b || !y.qkv_w || !y.qkv_b || !y.out_w || !y.out_b || !y.ln2_g || !y.ln2_b || !y.fc1_w ||
        !y.fc1_b || !y.fc2_w || !y.fc2_b) {
      delete h;
      delete[] layers;
      SCB_REQUIRE(false, SCB_ERR_INVALID_ARG, "scb_vit_encoder_create: NULL pointer for layer %d", l);
    }
  }
  h->magic = kVitMagic;
  h->d = *d;
  h->layers = layers;
  h->grid = grid;
  h->tokens = grid * grid + 1;
  h->pooled = (grid / 2) * (grid / 2);
  *out_handle = h;
  return SCB_OK;
}

int scb_vit_encoder_destroy(void* handle) {
  VitHandle* h = reinterpret_cast<VitHandle*>(handle);
  SCB_REQUIRE(h != nullptr && h->magic == kVitMagic, SCB_ERR_INVALID_ARG, "scb_vit_encoder_destroy: bad handle");
  h->magic = 0;
  delete[] h->layers;
  delete h;
  return SCB_OK;
}

size_t scb_vit_encode_workspace_bytes(void* handle, int32_t n_frames) {
  VitHandle* h = reinterpret_cast<VitHandle*>(handle);
  if (h == nullptr || h->magic != kVitMagic || n_frames < 1) return 0;
  return carve(h, n_frames, nullptr).total;
}

int scb_vit_encode(void* handle, const void* pixels, int32_t n_frames, void* out, void* tower_out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  VitHandle* h = reinterpret_cast<VitHandle*>(handle);
  SCB_REQUIRE(h != nullptr && h->magic == kVitMagic, SCB_ERR_INVALID_ARG, "scb_vit_encode: bad encoder handle");
  SCB_REQUIRE(pixels && out && workspace, SCB_ERR_INVALID_ARG, "scb_vit_encode: NULL buffer");
  SCB_REQUIRE(n_frames >= 1, SCB_ERR_INVALID_ARG, "scb_vit_encode: n_frames=%d", n_frames);
  const size_t need = carve(h, n_frames, nullptr).total;
  SCB_REQUIRE(workspace_bytes >= need, SCB_ERR_WORKSPACE, "scb_vit_encode: workspace %zu B < required %zu B",
              workspace_bytes, need);
  const scb_vit_encoder_desc& d = h->d;
  const VitBuffers b = carve(h, n_frames, reinterpret_cast<char*>(workspace));
  const int64_t n = n_frames, T = h->tokens, rows = n * T, prow = n * (T - 1);
  const int64_t D = d.hidden, I = d.inter, H = d.out_dim;
  const int hd = d.hidden / d.n_heads;

  auto linear = [&](const void* a, int64_t m, int64_t k, int64_t lda, const void* w, int64_t nn, const void* bias,
                    const void* res, void* o, int act) {
    scb_linear_args la;
    memset(&la, 0, sizeof(la));
    la.a = a; la.w = w; la.bias = bias; la.residual = res; la.out = o;
    la.m = m; la.n = nn; la.k = k;
    la.lda = lda; la.ldw = k; la.ldr = nn; la.ldo = nn;
    la.act = act; la.out_dtype = SCB_F16; la.batch = 1; la.alpha = 1.f;
    return scb_linear(&la, stream);
  };

  int rc = scb_im2col_patch(pixels, b.cols, n_frames, d.channels, d.image_size, d.image_size, d.patch, d.kpad, stream);
  if (rc == SCB_OK) rc = linear(b.cols, prow, d.kpad, d.kpad, d.patch_w, D, nullptr, nullptr, b.patch, SCB_ACT_NONE);
  if (rc == SCB_OK)
    rc = scb_vit_embed_layernorm(b.patch, D, d.cls, d.pos, d.pre_g, d.pre_b, d.ln_eps, n_frames, (int32_t)T, (int32_t)D,
                                 b.xa, D, stream);
  char* x = b.xa;
  char* x2 = b.xb;
  const float scale = 1.f / sqrtf((float)hd);
  for (int l = 0; l < d.n_layers && rc == SCB_OK; ++l) {
    const VitLayer& ly = h->layers[l];
    rc = scb_layernorm(x, D, ly.ln1_g, ly.ln1_b, d.ln_eps, rows, (int32_t)D, b.y, D, stream);
    if (rc == SCB_OK) rc = linear(b.y, rows, D, D, ly.qkv_w, 3 * D, ly.qkv_b, nullptr, b.wide, SCB_ACT_NONE);
    if (rc == SCB_OK) {
      scb_attention_args aa;
      memset(&aa, 0, sizeof(aa));
      __half* qkv = reinterpret_cast<__half*>(b.wide);
      aa.q = qkv; aa.k = qkv + D; aa.v = qkv + 2 * D; aa.out = b.ctx;
      aa.batch = n_frames; aa.n_q_heads = d.n_heads; aa.n_kv_heads = d.n_heads; aa.head_dim = hd;
      aa.q_len = T; aa.kv_len = T;
      aa.q_stride_b = aa.k_stride_b = aa.v_stride_b = T * 3 * D;
      aa.q_stride_t = aa.k_stride_t = aa.v_stride_t = 3 * D;
      aa.q_stride_h = aa.k_stride_h = aa.v_stride_h = hd;
      aa.o_stride_b = T * D; aa.o_stride_t = D; aa.o_stride_h = hd;
      aa.scale = scale; aa.causal = 0;
      rc = scb_attention(&aa, stream);
    }
    if (rc == SCB_OK) rc = linear(b.ctx, rows, D, D, ly.out_w, D, ly.out_b, x, x2, SCB_ACT_NONE);
    if (rc == SCB_OK) rc = scb_layernorm(x2, D, ly.ln2_g, ly.ln2_b, d.ln_eps, rows, (int32_t)D, b.y, D, stream);
    if (rc == SCB_OK) rc = linear(b.y, rows, D, D, ly.fc1_w, I, ly.fc1_b, nullptr, b.fc1, SCB_ACT_QUICK_GELU);
    if (rc == SCB_OK) rc = linear(b.fc1, rows, I, I, ly.fc2_w, D, ly.fc2_b, x2, x, SCB_ACT_NONE);
  }
  if (rc != SCB_OK) return rc;
  if (tower_out != nullptr)   // hidden_states[select_layer] incl. the CLS rows (feature_select drops them later)
    SCB_CUDA_OK(cudaMemcpyAsync(tower_out, x, (size_t)rows * D * 2, cudaMemcpyDeviceToDevice,
                                reinterpret_cast<cudaStream_t>(stream)));
  // projector Linear-1 + exact GELU over every row (CLS rows ride along), 2x2 mean over the patch grid (drops CLS),
  // then Linear-2 on the pooled rows (the mean commutes with Linear-2 exactly in real arithmetic)
  rc = linear(x, rows, D, D, d.proj1_w, H, d.proj1_b, nullptr, b.wide, SCB_ACT_GELU_ERF);
  if (rc == SCB_OK)
    rc = scb_avgpool2x2_tokens(b.wide, H, T, 1, h->grid, n_frames, (int32_t)H, b.pool, H, stream);
  if (rc == SCB_OK)
    rc = linear(b.pool, n * h->pooled, H, H, d.proj2_w, H, d.proj2_b, nullptr, out, SCB_ACT_NONE);
  return rc;
}

}  // extern "C"
