// decode_ops.cu — batch-1..4 token generation kernels (HBM-bound weight / KV streaming).
//
// Replaces, for the cached decode steps of HF generate (llava_qwen.py:155 ->
// Qwen2ForCausalLM.generate, forced use_cache=True: generation/utils.py:2425-2428), the cuBLAS
// GEMV-shaped Linear calls and the batch-1 SDPA over the DynamicCache.
//
//   gemv_kernel        out[r,n] = sum_k a'[r,k] w[n,k] (+bias[n]) (+residual[r,n]),  r < R <= 4.
//                      a' = a | RMSNorm(a)*gamma | silu(a[:K])*a[K:]  (prologue fused while the
//                      activation is staged in shared memory, so a decode layer is 4 GEMV
//                      launches and no stand-alone norm / gate kernels).  Every warp streams whole
//                      weight rows with 128-bit L1-bypassing loads, 4 in flight per lane.
//                      Algorithmic bytes: n*k*2 (+ activations).
//   decode_attn_kernel one query token, GQA: block = (kv head, split of the cached positions);
//                      a warp owns a position at a time, K/V rows are read once for all q heads
//                      of the group; online softmax per warp, merged per block, partial
//                      (m, l, o) written for the combine kernel.  Context length is read from
//                      device memory so the launch is CUDA-graph replayable.
//                      Algorithmic bytes: 2 * n_kv * ctx * head_dim * 2.
#include "scb_common.cuh"
#include <math.h>

namespace {

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;

SCB_DEVICE void unpack8(const uint4& v, float* f) {
  float2 a = h2_to_f2(v.x), b = h2_to_f2(v.y), c = h2_to_f2(v.z), d = h2_to_f2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

enum { PRO_NONE = 0, PRO_RMSNORM = 1, PRO_SILU_MUL = 2 };

struct GemvParams {
  const __half* a;      // [R, lda]
  const __half* w;      // [N, ldw]
  const __half* bias;   // [N] or null
  const __half* residual;  // [R, ldr] or null
  const __half* gamma;  // RMSNorm weight [K] (PRO_RMSNORM)
  void* out;            // [R, ldo] fp16 / fp32
  long long lda, ldw, ldr, ldo;
  int N, K;
  int out_f32;
  float eps;
};

template <int R, int PRO>
__global__ void __launch_bounds__(kGemvThreads)
gemv_kernel(const GemvParams p) {
  extern __shared__ __align__(16) __half a_s[];  // [R][K] transformed activations (fp16 values)
  __shared__ float red[kGemvWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K;

  // ---- prologue: stage (and transform) the activation rows
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const __half* ar = p.a + (long long)r * p.lda;
    if constexpr (PRO == PRO_RMSNORM) {
      float ss = 0.f;
      for (int i = tid; i < (K >> 3); i += kGemvThreads) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(ar + i * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
      }
      ss = warp_sum(ss);
      __syncthreads();
      if (lane == 0) red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kGemvWarps; ++w) tot += red[w];
      const float rstd = rsqrtf(tot / (float)K + p.eps);
      for (int i = tid; i < (K >> 3); i += kGemvThreads) {
        float f[8], g[8];
        unpack8(*reinterpret_cast<const uint4*>(ar + i * 8), f);
        unpack8(*reinterpret_cast<const uint4*>(p.gamma + i * 8), g);
        uint4 o;
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = __half2float(__float2half_rn(f[e] * rstd)) * g[e];
        o.x = f2_to_h2(t[0], t[1]); o.y = f2_to_h2(t[2], t[3]);
        o.z = f2_to_h2(t[4], t[5]); o.w = f2_to_h2(t[6], t[7]);
        *reinterpret_cast<uint4*>(a_s + r * K + i * 8) = o;
      }
    } else if constexpr (PRO == PRO_SILU_MUL) {
      for (int i = tid; i < (K >> 3); i += kGemvThreads) {
        float g[8], u[8], t[8];
        unpack8(*reinterpret_cast<const uint4*>(ar + i * 8), g);
        unpack8(*reinterpret_cast<const uint4*>(ar + K + i * 8), u);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          t[e] = __half2float(__float2half_rn(g[e] / (1.f + __expf(-g[e])))) * u[e];
        uint4 o;
        o.x = f2_to_h2(t[0], t[1]); o.y = f2_to_h2(t[2], t[3]);
        o.z = f2_to_h2(t[4], t[5]); o.w = f2_to_h2(t[6], t[7]);
        *reinterpret_cast<uint4*>(a_s + r * K + i * 8) = o;
      }
    } else {
      for (int i = tid; i < (K >> 3); i += kGemvThreads)
        *reinterpret_cast<uint4*>(a_s + r * K + i * 8) =
            *reinterpret_cast<const uint4*>(ar + i * 8);
    }
  }
  __syncthreads();

  // ---- main loop: one weight row per warp at a time
  const int nvec = K >> 3;
  for (int n = blockIdx.x * kGemvWarps + warp; n < p.N; n += gridDim.x * kGemvWarps) {
    const __half* wr = p.w + (long long)n * p.ldw;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    int i = lane;
    for (; i + 96 < nvec; i += 128) {  // 4 x 16 B per lane in flight
      uint4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = ld_stream_u4(wr + (i + u * 32) * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float wf[8];
        unpack8(wv[u], wf);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float af[8];
          unpack8(*reinterpret_cast<const uint4*>(a_s + r * K + (i + u * 32) * 8), af);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r] = fmaf(wf[e], af[e], acc[r]);
        }
      }
    }
    for (; i < nvec; i += 32) {
      float wf[8];
      unpack8(ld_stream_u4(wr + i * 8), wf);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float af[8];
        unpack8(*reinterpret_cast<const uint4*>(a_s + r * K + i * 8), af);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r] = fmaf(wf[e], af[e], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float v = acc[r];
        if (p.bias) v += __half2float(p.bias[n]);
        if (p.residual) v += __half2float(p.residual[(long long)r * p.ldr + n]);
        if (p.out_f32)
          reinterpret_cast<float*>(p.out)[(long long)r * p.ldo + n] = v;
        else
          reinterpret_cast<__half*>(p.out)[(long long)r * p.ldo + n] = __float2half_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// decode attention
// ------------------------------------------------------------------------------------------
constexpr int kDAThreads = 256;
constexpr int kDAWarps = kDAThreads / 32;
constexpr int kMaxGroup = 8;  // q heads per kv head

template <int HD, int G>
__global__ void __launch_bounds__(kDAThreads)
decode_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ k_cache,
                   const __half* __restrict__ v_cache, long long max_pos,
                   const int* __restrict__ kv_len_dev, int kv_len_host, float scale_log2,
                   float* __restrict__ part_m, float* __restrict__ part_l,
                   float* __restrict__ part_o) {
  // lane owns HD/32 consecutive dims of every row
  constexpr int E = HD / 32;
  const int hk = blockIdx.x, split = blockIdx.y, n_splits = gridDim.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kv_len = kv_len_dev ? *kv_len_dev : kv_len_host;
  const int per = (kv_len + n_splits - 1) / n_splits;
  const int begin = split * per, end = min(kv_len, begin + per);

  float qv[G][E];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int e = 0; e < E; ++e)
      qv[g][e] = __half2float(q[(long long)(hk * G + g) * HD + lane * E + e]);

  float m[G], l[G], o[G][E];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) o[g][e] = 0.f;
  }
  const __half* kb = k_cache + (long long)hk * max_pos * HD;
  const __half* vb = v_cache + (long long)hk * max_pos * HD;
  for (int pos = begin + warp; pos < end; pos += kDAWarps) {
    float kf[E], vf[E];
    if constexpr (E == 4) {
      const uint2 kr = *reinterpret_cast<const uint2*>(kb + (long long)pos * HD + lane * 4);
      const uint2 vr = *reinterpret_cast<const uint2*>(vb + (long long)pos * HD + lane * 4);
      float2 a = h2_to_f2(kr.x), b = h2_to_f2(kr.y), c = h2_to_f2(vr.x), d = h2_to_f2(vr.y);
      kf[0] = a.x; kf[1] = a.y; kf[2] = b.x; kf[3] = b.y;
      vf[0] = c.x; vf[1] = c.y; vf[2] = d.x; vf[3] = d.y;
    } else {
      const uint32_t kr = *reinterpret_cast<const uint32_t*>(kb + (long long)pos * HD + lane * 2);
      const uint32_t vr = *reinterpret_cast<const uint32_t*>(vb + (long long)pos * HD + lane * 2);
      float2 a = h2_to_f2(kr), c = h2_to_f2(vr);
      kf[0] = a.x; kf[1] = a.y; vf[0] = c.x; vf[1] = c.y;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) s = fmaf(qv[g][e], kf[e], s);
      s = warp_sum(s) * scale_log2;
      const float mn = fmaxf(m[g], s);
      const float alpha = exp2f(m[g] - mn);
      const float pexp = exp2f(s - mn);
      l[g] = l[g] * alpha + pexp;
#pragma unroll
      for (int e = 0; e < E; ++e) o[g][e] = fmaf(o[g][e], alpha, pexp * vf[e]);
      m[g] = mn;
    }
  }
  // ---- merge the warps of the block (fixed order)
  __shared__ float sm_m[kDAWarps][G], sm_l[kDAWarps][G];
  __shared__ float sm_o[kDAWarps][G][HD];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (lane == 0) {
      sm_m[warp][g] = m[g];
      sm_l[warp][g] = l[g];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) sm_o[warp][g][lane * E + e] = o[g][e];
  }
  __syncthreads();
  for (int idx = tid; idx < G * HD; idx += kDAThreads) {
    const int g = idx / HD, d = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDAWarps; ++w) mm = fmaxf(mm, sm_m[w][g]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < kDAWarps; ++w) {
      const float sc = (sm_m[w][g] == -INFINITY) ? 0.f : exp2f(sm_m[w][g] - mm);
      ll = fmaf(sm_l[w][g], sc, ll);
      oo = fmaf(sm_o[w][g][d], sc, oo);
    }
    const long long pi = (long long)(hk * G + g) * n_splits + split;
    part_o[pi * HD + d] = oo;
    if (d == 0) {
      part_m[pi] = mm;
      part_l[pi] = ll;
    }
  }
}

template <int HD>
__global__ void decode_attn_combine_kernel(const float* __restrict__ part_m,
                                           const float* __restrict__ part_l,
                                           const float* __restrict__ part_o, int n_splits,
                                           __half* __restrict__ out) {
  const int h = blockIdx.x, d = threadIdx.x;
  float mm = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mm = fmaxf(mm, part_m[(long long)h * n_splits + s]);
  float ll = 0.f, oo = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const long long pi = (long long)h * n_splits + s;
    const float sc = (part_m[pi] == -INFINITY) ? 0.f : exp2f(part_m[pi] - mm);
    ll = fmaf(part_l[pi], sc, ll);
    oo = fmaf(part_o[pi * HD + d], sc, oo);
  }
  out[(long long)h * HD + d] = __float2half_rn(ll > 0.f ? oo / ll : 0.f);
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int R>
int launch_gemv(const GemvParams& p, int prologue, int grid, size_t smem, cudaStream_t st) {
#define SCB_GEMV_CASE(PRO)                                                                       \
  do {                                                                                           \
    auto kern = gemv_kernel<R, PRO>;                                                             \
    if (smem > 48 * 1024)                                                                        \
      SCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                       (int)smem));                                              \
    kern<<<grid, kGemvThreads, smem, st>>>(p);                                                   \
  } while (0)
  if (prologue == PRO_RMSNORM) SCB_GEMV_CASE(PRO_RMSNORM);
  else if (prologue == PRO_SILU_MUL) SCB_GEMV_CASE(PRO_SILU_MUL);
  else SCB_GEMV_CASE(PRO_NONE);
#undef SCB_GEMV_CASE
  SCB_LAUNCH_OK("gemv_kernel");
  return SCB_OK;
}

}  // namespace

extern "C" {

int scb_gemv(const scb_gemv_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_gemv: args is NULL");
  SCB_REQUIRE(a->a && a->w && a->out, SCB_ERR_INVALID_ARG, "scb_gemv: NULL buffer");
  SCB_REQUIRE(a->rows == 1 || a->rows == 2 || a->rows == 4, SCB_ERR_UNSUPPORTED,
              "scb_gemv: rows=%d (1, 2 or 4; larger batches use scb_linear)", a->rows);
  SCB_REQUIRE(a->n >= 1 && a->k >= 8 && a->k % 8 == 0, SCB_ERR_INVALID_ARG,
              "scb_gemv: n=%lld k=%lld (k must be a multiple of 8)", (long long)a->n, (long long)a->k);
  SCB_REQUIRE(a->prologue >= PRO_NONE && a->prologue <= PRO_SILU_MUL, SCB_ERR_UNSUPPORTED,
              "scb_gemv: prologue %d unsupported", a->prologue);
  SCB_REQUIRE(a->prologue != PRO_RMSNORM || a->gamma, SCB_ERR_INVALID_ARG,
              "scb_gemv: RMSNorm prologue needs gamma");
  const long long min_lda = a->prologue == PRO_SILU_MUL ? 2 * a->k : a->k;
  SCB_REQUIRE(a->lda >= min_lda && a->ldw >= a->k && a->ldo >= a->n && a->lda % 8 == 0 &&
                  a->ldw % 8 == 0,
              SCB_ERR_INVALID_ARG, "scb_gemv: bad leading dimensions");
  SCB_REQUIRE(al16(a->a) && al16(a->w) && (!a->gamma || al16(a->gamma)), SCB_ERR_INVALID_ARG,
              "scb_gemv: a / w / gamma must be 16-byte aligned");
  SCB_REQUIRE(!a->residual || a->ldr >= a->n, SCB_ERR_INVALID_ARG, "scb_gemv: ldr too small");
  const size_t smem = (size_t)a->rows * a->k * sizeof(__half);
  SCB_REQUIRE(smem <= 200 * 1024, SCB_ERR_UNSUPPORTED,
              "scb_gemv: rows*k=%lld does not fit in shared memory", (long long)a->rows * a->k);
  GemvParams p;
  p.a = reinterpret_cast<const __half*>(a->a);
  p.w = reinterpret_cast<const __half*>(a->w);
  p.bias = reinterpret_cast<const __half*>(a->bias);
  p.residual = reinterpret_cast<const __half*>(a->residual);
  p.gamma = reinterpret_cast<const __half*>(a->gamma);
  p.out = a->out;
  p.lda = a->lda; p.ldw = a->ldw; p.ldr = a->ldr; p.ldo = a->ldo;
  p.N = (int)a->n; p.K = (int)a->k;
  p.out_f32 = a->out_dtype == SCB_F32;
  p.eps = a->eps;
  // resident blocks per SM are bounded by the staged activations; aim for >= 32 warps per SM
  const int sms = scb_num_sms();
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)a->n + kGemvWarps - 1) / kGemvWarps;
  long long cap = (long long)sms * per_sm;
  int grid = (int)(want < cap ? want : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->rows == 1) return launch_gemv<1>(p, a->prologue, grid, smem, st);
  if (a->rows == 2) return launch_gemv<2>(p, a->prologue, grid, smem, st);
  return launch_gemv<4>(p, a->prologue, grid, smem, st);
}

size_t scb_decode_attention_workspace_bytes(int32_t n_q_heads, int32_t head_dim, int32_t n_splits) {
  if (n_q_heads < 1 || head_dim < 1 || n_splits < 1) return 0;
  return sizeof(float) * (size_t)n_q_heads * n_splits * (head_dim + 2);
}

int scb_decode_attention(const scb_decode_attention_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_decode_attention: args is NULL");
  SCB_REQUIRE(a->q && a->k_cache && a->v_cache && a->out && a->workspace, SCB_ERR_INVALID_ARG,
              "scb_decode_attention: NULL buffer");
  SCB_REQUIRE(a->head_dim == 128 || a->head_dim == 64, SCB_ERR_UNSUPPORTED,
              "scb_decode_attention: head_dim=%d (64 or 128)", a->head_dim);
  SCB_REQUIRE(a->n_kv_heads >= 1 && a->n_q_heads % a->n_kv_heads == 0, SCB_ERR_INVALID_ARG,
              "scb_decode_attention: bad head counts");
  const int G = a->n_q_heads / a->n_kv_heads;
  SCB_REQUIRE(G == 1 || G == 2 || G == 4 || G == 7 || G == 8, SCB_ERR_UNSUPPORTED,
              "scb_decode_attention: group size %d unsupported (1,2,4,7,8)", G);
  SCB_REQUIRE(a->n_splits >= 1 && a->n_splits <= 1024, SCB_ERR_INVALID_ARG,
              "scb_decode_attention: n_splits out of range");
  SCB_REQUIRE(a->kv_len_dev || (a->kv_len >= 1 && a->kv_len <= a->max_pos), SCB_ERR_INVALID_ARG,
              "scb_decode_attention: kv_len out of range");
  const size_t need = scb_decode_attention_workspace_bytes(a->n_q_heads, a->head_dim, a->n_splits);
  SCB_REQUIRE(a->workspace_bytes >= need, SCB_ERR_WORKSPACE,
              "scb_decode_attention: workspace %zu B < %zu B", a->workspace_bytes, need);
  float* part_m = reinterpret_cast<float*>(a->workspace);
  float* part_l = part_m + (size_t)a->n_q_heads * a->n_splits;
  float* part_o = part_l + (size_t)a->n_q_heads * a->n_splits;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(a->n_kv_heads, a->n_splits);
  const float sl2 = a->scale * 1.4426950408889634f;
  const __half* q = reinterpret_cast<const __half*>(a->q);
  const __half* kc = reinterpret_cast<const __half*>(a->k_cache);
  const __half* vc = reinterpret_cast<const __half*>(a->v_cache);
#define SCB_DA(HD, GG)                                                                         \
  decode_attn_kernel<HD, GG><<<grid, kDAThreads, 0, st>>>(q, kc, vc, a->max_pos, a->kv_len_dev, \
                                                          a->kv_len, sl2, part_m, part_l, part_o)
  if (a->head_dim == 128) {
    switch (G) {
      case 1: SCB_DA(128, 1); break;
      case 2: SCB_DA(128, 2); break;
      case 4: SCB_DA(128, 4); break;
      case 7: SCB_DA(128, 7); break;
      default: SCB_DA(128, 8); break;
    }
  } else {
    switch (G) {
      case 1: SCB_DA(64, 1); break;
      case 2: SCB_DA(64, 2); break;
      case 4: SCB_DA(64, 4); break;
      case 7: SCB_DA(64, 7); break;
      default: SCB_DA(64, 8); break;
    }
  }
#undef SCB_DA
  SCB_LAUNCH_OK("decode_attn_kernel");
  if (a->head_dim == 128)
    decode_attn_combine_kernel<128><<<a->n_q_heads, 128, 0, st>>>(
        part_m, part_l, part_o, a->n_splits, reinterpret_cast<__half*>(a->out));
  else
    decode_attn_combine_kernel<64><<<a->n_q_heads, 64, 0, st>>>(
        part_m, part_l, part_o, a->n_splits, reinterpret_cast<__half*>(a->out));
  SCB_LAUNCH_OK("decode_attn_combine_kernel");
  return SCB_OK;
}

}  // extern "C"
