// rowwise_ops.cu — the HBM-bound glue between the tensor-core layers: patch im2col, token
// assembly + LayerNorm, LayerNorm, RMSNorm, 2x2 token average pool, SiLU-gate, RoPE + KV-cache
// append, embedding gather, arg-max.  All fp16 I/O with fp32 statistics, 128-bit accesses.
//
// Operator order follows HF transformers 5.5.0 as executed by the reference:
//   CLIP embeddings/pre-LN  models/clip/modeling_clip.py:148-217,659-677  (via clip_encoder.py:76)
//   SpatialPool             longva/model/multimodal_resampler/spatial_pool.py:23-32
//   Qwen2 RMSNorm / RoPE    models/qwen2/modeling_qwen2.py:249-263,51-146 (via llava_qwen.py:155)
#include "scb_common.cuh"
#include <math.h>

namespace {

SCB_DEVICE float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

SCB_DEVICE void unpack8(const uint4& v, float* f) {
  float2 a = h2_to_f2(v.x), b = h2_to_f2(v.y), c = h2_to_f2(v.z), d = h2_to_f2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
SCB_DEVICE uint4 pack8(const float* f) {
  uint4 o;
  o.x = f2_to_h2(f[0], f[1]); o.y = f2_to_h2(f[2], f[3]);
  o.z = f2_to_h2(f[4], f[5]); o.w = f2_to_h2(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------
// im2col for a stride==kernel patch convolution (CLIP patch embed, Conv2d k=14 s=14 no bias):
// out[(n*gh + py)*gw + px][c*P*P + ky*P + kx] = pix[n][c][py*P+ky][px*P+kx]; columns >= C*P*P
// (padding up to ld_out so that the row stride is TMA-legal) are zero.
// ------------------------------------------------------------------------------------------
__global__ void im2col_patch_kernel(const __half* __restrict__ pix, __half* __restrict__ out,
                                    int n_img, int C, int H, int W, int P, int ld_out) {
  const int gh = H / P, gw = W / P;
  const long long row = blockIdx.x;  // one block per output row (patch)
  const int n = (int)(row / (gh * gw));
  const int pr = (int)(row % (gh * gw));
  const int py = pr / gw, px = pr % gw;
  const int kdim = C * P * P;
  for (int col = threadIdx.x; col < ld_out; col += blockDim.x) {
    __half v = __float2half(0.f);
    if (col < kdim) {
      const int c = col / (P * P), r = col % (P * P), ky = r / P, kx = r % P;
      v = pix[(((long long)n * C + c) * H + py * P + ky) * W + px * P + kx];
    }
    out[row * ld_out + col] = v;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm family.  One block of 128 threads per row; the row lives in registers (D <= 8192).
//   kind 0: LayerNorm(x)                       (gamma, beta, eps)   two-pass mean / variance
//   kind 1: RMSNorm(x)                         HF order: fp32 normalise -> round to fp16 -> * weight
//   kind 2: LayerNorm(token assembly)          x = (t==0 ? cls : patch[n*(T-1)+t-1]) + pos[t]
// ------------------------------------------------------------------------------------------
constexpr int kNormThreads = 128;
constexpr int kNormMaxChunks = 8;  // 128 threads * 8 halfs * 8 chunks = 8192

template <int KIND>
__global__ void __launch_bounds__(kNormThreads)
norm_kernel(const __half* __restrict__ x, long long ld_x, const __half* __restrict__ gamma,
            const __half* __restrict__ beta, float eps, int D, __half* __restrict__ out,
            long long ld_out, const __half* __restrict__ cls, const __half* __restrict__ pos,
            int tokens_per_img) {
  __shared__ float red[kNormThreads / 32];
  pdl_launch_dependents();
  pdl_wait();   // decode chain: x comes from the preceding GEMV
  const long long row = blockIdx.x;
  const int nvec = D >> 3;
  float v[kNormMaxChunks][8];
  const __half* src;
  const __half* posrow = nullptr;
  if constexpr (KIND == 2) {
    const int n = (int)(row / tokens_per_img), t = (int)(row % tokens_per_img);
    src = (t == 0) ? cls : x + ((long long)n * (tokens_per_img - 1) + (t - 1)) * ld_x;
    posrow = pos + (long long)t * D;
  } else {
    src = x + row * ld_x;
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      unpack8(*reinterpret_cast<const uint4*>(src + i * 8), v[c]);
      if constexpr (KIND == 2) {
        float p[8];
        unpack8(*reinterpret_cast<const uint4*>(posrow + i * 8), p);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = __half2float(__float2half_rn(v[c][e] + p[e]));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += (KIND == 1) ? v[c][e] * v[c][e] : v[c][e];
    }
  }
  s = block_sum(s, red);
  float mean = 0.f, rstd;
  if constexpr (KIND == 1) {
    rstd = rsqrtf(s / (float)D + eps);
  } else {
    mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int i = c * kNormThreads + threadIdx.x;
      if (i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[c][e] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    q = block_sum(q, red);
    rstd = rsqrtf(q / (float)D + eps);
  }
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      float g[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + i * 8), g);
      if constexpr (KIND == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float nrm = __half2float(__float2half_rn(v[c][e] * rstd));
          o[e] = nrm * g[e];
        }
      } else {
        float b[8];
        unpack8(*reinterpret_cast<const uint4*>(beta + i * 8), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * g[e] + b[e];
      }
      *reinterpret_cast<uint4*>(out + row * ld_out + i * 8) = pack8(o);
    }
  }
}

// Warp-per-row variant for D <= 2048 (CLIP width): 8 rows per block, the row lives in the lanes'
// registers (<= 8 x 16 B each), statistics by shuffles only — no block barriers, 4x fewer blocks.
constexpr int kWarpNormMaxChunks = 8;

// CHUNKS = 16-byte pieces of the row per lane: 4 for D <= 1024 (the CLIP width: half the registers, twice the resident
// warps — the kernel is bound by bytes in flight per SM), 8 up to 2048.
template <int KIND, int CHUNKS = kWarpNormMaxChunks>
__global__ void __launch_bounds__(256, CHUNKS <= 4 ? 4 : 2)
norm_warp_kernel(const __half* __restrict__ x, long long ld_x, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, float eps, int D, long long rows,
                 __half* __restrict__ out, long long ld_out, const __half* __restrict__ cls,
                 const __half* __restrict__ pos, int tokens_per_img) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = D >> 3;
  float v[CHUNKS][8];
  const __half* src;
  const __half* posrow = nullptr;
  if constexpr (KIND == 2) {
    const int n = (int)(row / tokens_per_img), t = (int)(row % tokens_per_img);
    src = (t == 0) ? cls : x + ((long long)n * (tokens_per_img - 1) + (t - 1)) * ld_x;
    posrow = pos + (long long)t * D;
  } else {
    src = x + row * ld_x;
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = c * 32 + lane;
    if (i < nvec) {
      unpack8(*reinterpret_cast<const uint4*>(src + i * 8), v[c]);
      if constexpr (KIND == 2) {
        float p[8];
        unpack8(*reinterpret_cast<const uint4*>(posrow + i * 8), p);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = __half2float(__float2half_rn(v[c][e] + p[e]));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += (KIND == 1) ? v[c][e] * v[c][e] : v[c][e];
    }
  }
  s = warp_sum(s);
  float mean = 0.f, rstd;
  if constexpr (KIND == 1) {
    rstd = rsqrtf(s / (float)D + eps);
  } else {
    mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = c * 32 + lane;
      if (i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[c][e] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    q = warp_sum(q);
    rstd = rsqrtf(q / (float)D + eps);
  }
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = c * 32 + lane;
    if (i < nvec) {
      float g[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + i * 8), g);
      if constexpr (KIND == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __half2float(__float2half_rn(v[c][e] * rstd)) * g[e];
      } else {
        float b[8];
        unpack8(*reinterpret_cast<const uint4*>(beta + i * 8), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * g[e] + b[e];
      }
      *reinterpret_cast<uint4*>(out + row * ld_out + i * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Sentence embedding tail (sentence-transformers Pooling(mean) + Normalize): text t owns token rows
// [offsets[t], offsets[t+1]) of the packed hidden states; out[t] = mean over its rows, divided by max(|mean|_2, 1e-12).
// One CTA per text, thread = 8 consecutive dims (D <= 8 * blockDim), fp32 sums in token order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mean_pool_normalize_kernel(const __half* __restrict__ x, long long ld_x,
                                                                  const int* __restrict__ offsets, int D,
                                                                  float* __restrict__ out, long long ld_out,
                                                                  int normalize) {
  const int t = blockIdx.x;
  const int r0 = offsets[t], r1 = offsets[t + 1];
  const int i = threadIdx.x;
  const bool live = i < (D >> 3);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live)
    for (int r = r0; r < r1; ++r) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + (long long)r * ld_x + i * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  const float inv = 1.f / fmaxf((float)(r1 - r0), 1e-9f);          // Pooling: sum / clamp(sum(mask), 1e-9)
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] *= inv;
    ss += acc[e] * acc[e];
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const float scale = normalize ? 1.f / fmaxf(sqrtf(tot), 1e-12f) : 1.f;   // F.normalize(p=2, eps=1e-12)
  if (live) {
#pragma unroll
    for (int e = 0; e < 8; ++e) out[(long long)t * ld_out + i * 8 + e] = acc[e] * scale;
  }
}

// ------------------------------------------------------------------------------------------
// 2x2 average pool over a g x g token grid (SpatialPool avg, stride 2): in rows are
// [n][tok_offset + y*g + x], out rows [n][(y/2)*(g/2) + x/2].
// ------------------------------------------------------------------------------------------
__global__ void pool2x2_kernel(const __half* __restrict__ in, long long ld_in,
                               long long img_stride_rows, int tok_offset, int g, int D,
                               __half* __restrict__ out, long long ld_out) {
  const int og = g / 2;
  const long long orow = blockIdx.x;
  const int n = (int)(orow / (og * og)), r = (int)(orow % (og * og));
  const int oy = r / og, ox = r % og;
  const __half* base = in + ((long long)n * img_stride_rows + tok_offset) * ld_in;
  for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(
                    base + (long long)((oy * 2 + dy) * g + ox * 2 + dx) * ld_in + i * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;
    *reinterpret_cast<uint4*>(out + orow * ld_out + i * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------
// SiLU gate: out[m][i] = silu(gu[m][i]) * gu[m][I + i]   (Qwen2 MLP, modeling_qwen2.py:41-48;
// fp16 rounding after silu and after the product, as the eager fp16 graph does)
// ------------------------------------------------------------------------------------------
__global__ void silu_mul_kernel(const __half* __restrict__ gu, long long ld_gu, int I,
                                __half* __restrict__ out, long long ld_out, long long rows) {
  const long long nvec = (long long)(I >> 3);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * nvec;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / nvec;
    const int i = (int)(idx % nvec) * 8;
    float g[8], u[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + m * ld_gu + i), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + m * ld_gu + I + i), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = __half2float(__float2half_rn(__fdividef(g[e], 1.f + __expf(-g[e]))));
      o[e] = s * u[e];
    }
    *reinterpret_cast<uint4*>(out + m * ld_out + i) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------
// RoPE (NeoX half-split) on the q and k slices of a fused qkv row + append of k, v to the
// per-layer KV cache.  qkv row layout: [n_q*hd | n_kv*hd | n_kv*hd].  Angles in fp32, cos/sin
// rounded to fp16, products and sum rounded to fp16 like the eager fp16 graph
// (modeling_qwen2.py:116-146).  Cache layout: [n_kv][max_pos][hd] for K and for V.
// ------------------------------------------------------------------------------------------
__global__ void rope_append_kernel(__half* __restrict__ qkv, long long ld_qkv, int n_q, int n_kv,
                                   int hd, const int* __restrict__ positions, int pos_base,
                                   float theta, const float* __restrict__ inv_freq_tab,
                                   __half* __restrict__ k_cache,
                                   __half* __restrict__ v_cache, long long max_pos,
                                   long long rows) {
  const long long row = blockIdx.x;
  if (row >= rows) return;
  const int pos = positions ? positions[row] : pos_base + (int)row;
  __half* r = qkv + row * ld_qkv;
  const int half = hd >> 1;
  const int n_rot = n_q + n_kv;
  for (int idx = threadIdx.x; idx < n_rot * half; idx += blockDim.x) {
    const int h = idx / half, i = idx % half;
    const float inv_freq = inv_freq_tab ? inv_freq_tab[i] : powf(theta, -(float)(2 * i) / (float)hd);
    const float ang = (float)pos * inv_freq;
    const float c = __half2float(__float2half_rn(cosf(ang)));
    const float s = __half2float(__float2half_rn(sinf(ang)));
    __half* p = r + (long long)h * hd;
    const float x1 = __half2float(p[i]), x2 = __half2float(p[i + half]);
    // out1 = x1*cos + (-x2)*sin ; out2 = x2*cos + x1*sin, each product rounded to fp16
    const float a1 = __half2float(__float2half_rn(x1 * c)), b1 = __half2float(__float2half_rn(-x2 * s));
    const float a2 = __half2float(__float2half_rn(x2 * c)), b2 = __half2float(__float2half_rn(x1 * s));
    const __half o1 = __float2half_rn(a1 + b1), o2 = __float2half_rn(a2 + b2);
    p[i] = o1;
    p[i + half] = o2;
    if (h >= n_q && k_cache) {
      __half* kc = k_cache + ((long long)(h - n_q) * max_pos + pos) * hd;
      kc[i] = o1;
      kc[i + half] = o2;
    }
  }
  if (v_cache) {
    const __half* v = r + (long long)(n_q + n_kv) * hd;
    for (int idx = threadIdx.x; idx < n_kv * hd; idx += blockDim.x) {
      const int h = idx / hd, i = idx % hd;
      v_cache[((long long)h * max_pos + pos) * hd + i] = v[idx];
    }
  }
}

// Vector form for head_dim % 16 == 0 (the shapes on the path): the hd/2 angles of the row's position are
// evaluated once per block, every thread then rotates 8 + 8 elements with 16-byte accesses.
__global__ void __launch_bounds__(256)
rope_append_vec_kernel(__half* __restrict__ qkv, long long ld_qkv, int n_q, int n_kv, int hd,
                       const int* __restrict__ positions, int pos_base, float theta,
                       const float* __restrict__ inv_freq_tab, __half* __restrict__ k_cache,
                       __half* __restrict__ v_cache, long long max_pos) {
  __shared__ float cs[128], sn[128];
  const long long row = blockIdx.x;
  const int pos = positions ? positions[row] : pos_base + (int)row;
  __half* r = qkv + row * ld_qkv;
  const int half = hd >> 1;
  for (int i = threadIdx.x; i < half; i += 256) {
    const float inv_freq = inv_freq_tab ? inv_freq_tab[i] : powf(theta, -(float)(2 * i) / (float)hd);
    const float ang = (float)pos * inv_freq;
    cs[i] = __half2float(__float2half_rn(cosf(ang)));
    sn[i] = __half2float(__float2half_rn(sinf(ang)));
  }
  __syncthreads();
  const int cpr = half >> 3;   // 16-byte chunks per half head
  for (int idx = threadIdx.x; idx < (n_q + n_kv) * cpr; idx += 256) {
    const int h = idx / cpr, c8 = (idx % cpr) * 8;
    __half* p = r + (long long)h * hd + c8;
    const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + half);
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x1 = h2_to_f2(lw[e]), x2 = h2_to_f2(hw[e]);
      const float c0 = cs[c8 + 2 * e], c1 = cs[c8 + 2 * e + 1], s0 = sn[c8 + 2 * e], s1 = sn[c8 + 2 * e + 1];
      // out1 = x1*cos + (-x2)*sin ; out2 = x2*cos + x1*sin, each product rounded to fp16
      const float2 a1 = h2_to_f2(f2_to_h2(x1.x * c0, x1.y * c1)), b1 = h2_to_f2(f2_to_h2(-x2.x * s0, -x2.y * s1));
      const float2 a2 = h2_to_f2(f2_to_h2(x2.x * c0, x2.y * c1)), b2 = h2_to_f2(f2_to_h2(x1.x * s0, x1.y * s1));
      o1[e] = f2_to_h2(a1.x + b1.x, a1.y + b1.y);
      o2[e] = f2_to_h2(a2.x + b2.x, a2.y + b2.y);
    }
    const uint4 v1 = make_uint4(o1[0], o1[1], o1[2], o1[3]), v2 = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    *reinterpret_cast<uint4*>(p) = v1;
    *reinterpret_cast<uint4*>(p + half) = v2;
    if (h >= n_q && k_cache) {
      __half* kc = k_cache + ((long long)(h - n_q) * max_pos + pos) * hd + c8;
      *reinterpret_cast<uint4*>(kc) = v1;
      *reinterpret_cast<uint4*>(kc + half) = v2;
    }
  }
  if (v_cache) {
    const __half* v = r + (long long)(n_q + n_kv) * hd;
    const int vpr = hd >> 3;
    for (int idx = threadIdx.x; idx < n_kv * vpr; idx += 256) {
      const int h = idx / vpr, c8 = (idx % vpr) * 8;
      *reinterpret_cast<uint4*>(v_cache + ((long long)h * max_pos + pos) * hd + c8) =
          *reinterpret_cast<const uint4*>(v + (long long)h * hd + c8);
    }
  }
}

// ------------------------------------------------------------------------------------------
// BERT embeddings (modeling_bert.py BertEmbeddings.forward): row r = LayerNorm(word[ids[r]] +
// position[pos[r]] + token_type[0]), sums and statistics in fp32.  One block per token.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNormThreads)
text_embed_layernorm_kernel(const long long* __restrict__ ids, const long long* __restrict__ pos,
                            const __half* __restrict__ word, const __half* __restrict__ pos_tab,
                            const __half* __restrict__ type_row, const __half* __restrict__ gamma,
                            const __half* __restrict__ beta, float eps, int D, __half* __restrict__ out,
                            long long ld_out) {
  __shared__ float red[kNormThreads / 32];
  const long long row = blockIdx.x;
  const __half* w = word + ids[row] * D;
  const __half* pp = pos_tab + pos[row] * D;
  const int nvec = D >> 3;
  float v[kNormMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      float a[8], b[8], t[8];
      unpack8(*reinterpret_cast<const uint4*>(w + i * 8), a);
      unpack8(*reinterpret_cast<const uint4*>(pp + i * 8), b);
      unpack8(*reinterpret_cast<const uint4*>(type_row + i * 8), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[c][e] = a[e] + t[e] + b[e];   // BertEmbeddings order: (word + token_type) + position
        s += v[c][e];
      }
    }
  }
  s = block_sum(s, red);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[c][e] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  q = block_sum(q, red);
  const float rstd = rsqrtf(q / (float)D + eps);
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      float g[8], b[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + i * 8), g);
      unpack8(*reinterpret_cast<const uint4*>(beta + i * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * g[e] + b[e];
      *reinterpret_cast<uint4*>(out + row * ld_out + i * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Embedding gather: out[r] = table[ids[r]]  (ids < 0 -> zeros)
// ------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const __half* __restrict__ table, const long long* __restrict__ ids,
                                   int D, __half* __restrict__ out, long long ld_out) {
  pdl_launch_dependents();
  pdl_wait();   // ids may come from the previous decode step's fused arg-max
  const long long row = blockIdx.x;
  const long long id = ids[row];
  for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (id >= 0) v = *reinterpret_cast<const uint4*>(table + id * D + i * 8);
    *reinterpret_cast<uint4*>(out + row * ld_out + i * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------
// Embedding gather fused with the first RMSNorm of the decode step: out[r] = table[ids[r]],
// norm_out[r] = RMSNorm(out[r]) * gamma (HF order).  One block per row, row in registers.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNormThreads)
gather_rmsnorm_kernel(const __half* __restrict__ table, const long long* __restrict__ ids, int D,
                      __half* __restrict__ out, long long ld_out, const __half* __restrict__ gamma,
                      float eps, __half* __restrict__ norm_out, long long ld_norm) {
  __shared__ float red[kNormThreads / 32];
  pdl_launch_dependents();
  pdl_wait();   // ids come from the previous decode step's fused arg-max
  const long long row = blockIdx.x;
  const long long id = ids[row];
  const int nvec = D >> 3;
  float v[kNormMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      uint4 q = make_uint4(0, 0, 0, 0);
      if (id >= 0) q = *reinterpret_cast<const uint4*>(table + id * D + i * 8);
      *reinterpret_cast<uint4*>(out + row * ld_out + i * 8) = q;
      unpack8(q, v[c]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(v[c][e], v[c][e], s);
    }
  }
  s = block_sum(s, red);
  const float rstd = rsqrtf(s / (float)D + eps);
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c) {
    const int i = c * kNormThreads + threadIdx.x;
    if (i < nvec) {
      float g[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + i * 8), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = __half2float(__float2half_rn(v[c][e] * rstd)) * g[e];
      *reinterpret_cast<uint4*>(norm_out + row * ld_norm + i * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Split-K tail for the wide-batch decode layers: out[r,n] = fp16(sum_s partial[s,r,n] + bias[n] +
// residual[r,n]) in split order (deterministic), optionally followed by the RMSNorm of the finished
// row (the next layer's input norm).  One block per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, long long split_stride, long long ld_partial,
                     int n_splits, int N, const __half* __restrict__ bias,
                     const __half* __restrict__ residual, long long ld_res, __half* __restrict__ out,
                     long long ld_out, const __half* __restrict__ pn_gamma, __half* __restrict__ pn_out,
                     long long pn_ld, float pn_eps) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  float ss = 0.f;
  constexpr int kC = 6;                    // rows up to 6144 wide stay in registers between the two passes
  const bool in_regs = N <= 256 * 4 * kC;
  uint2 keep[kC];
  if (in_regs) {
    float4 acc[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < n_splits; ++sp) {      // splits are summed in index order (deterministic)
      float4 v[kC];
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const int i = threadIdx.x + c * 256;
        v[c] = i < (N >> 2) ? __ldcg(reinterpret_cast<const float4*>(partial + sp * split_stride + (long long)r * ld_partial) + i)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        acc[c].x += v[c].x; acc[c].y += v[c].y; acc[c].z += v[c].z; acc[c].w += v[c].w;
      }
    }
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const int i = threadIdx.x + c * 256;
      keep[c] = make_uint2(0, 0);
      if (i < (N >> 2)) {
        float f[4] = {acc[c].x, acc[c].y, acc[c].z, acc[c].w};
        if (bias) {
          const uint2 b = *reinterpret_cast<const uint2*>(bias + i * 4);
          const float2 b0 = h2_to_f2(b.x), b1 = h2_to_f2(b.y);
          f[0] += b0.x; f[1] += b0.y; f[2] += b1.x; f[3] += b1.y;
        }
        if (residual) {
          const uint2 q = *reinterpret_cast<const uint2*>(residual + (long long)r * ld_res + i * 4);
          const float2 r0 = h2_to_f2(q.x), r1 = h2_to_f2(q.y);
          f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
        }
        uint2 o;
        o.x = f2_to_h2(f[0], f[1]);
        o.y = f2_to_h2(f[2], f[3]);
        *reinterpret_cast<uint2*>(out + (long long)r * ld_out + i * 4) = o;
        keep[c] = o;
        const float2 h0 = h2_to_f2(o.x), h1 = h2_to_f2(o.y);   // the norm sees the stored fp16 values
        ss += h0.x * h0.x + h0.y * h0.y + h1.x * h1.x + h1.y * h1.y;
      }
    }
  } else {
    for (int i = threadIdx.x; i < (N >> 2); i += 256) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sp = 0; sp < n_splits; ++sp) {
        const float4 v = *reinterpret_cast<const float4*>(partial + sp * split_stride + (long long)r * ld_partial + i * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      float f[4] = {acc.x, acc.y, acc.z, acc.w};
      if (bias) {
        const uint2 b = *reinterpret_cast<const uint2*>(bias + i * 4);
        const float2 b0 = h2_to_f2(b.x), b1 = h2_to_f2(b.y);
        f[0] += b0.x; f[1] += b0.y; f[2] += b1.x; f[3] += b1.y;
      }
      if (residual) {
        const uint2 q = *reinterpret_cast<const uint2*>(residual + (long long)r * ld_res + i * 4);
        const float2 r0 = h2_to_f2(q.x), r1 = h2_to_f2(q.y);
        f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
      }
      uint2 o;
      o.x = f2_to_h2(f[0], f[1]);
      o.y = f2_to_h2(f[2], f[3]);
      *reinterpret_cast<uint2*>(out + (long long)r * ld_out + i * 4) = o;
      const float2 h0 = h2_to_f2(o.x), h1 = h2_to_f2(o.y);
      ss += h0.x * h0.x + h0.y * h0.y + h1.x * h1.x + h1.y * h1.y;
    }
  }
  if (pn_out == nullptr) return;
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)N + pn_eps);
  if (in_regs) {
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const int i = threadIdx.x + c * 256;
      if (i < (N >> 2)) {
        const uint2 g = *reinterpret_cast<const uint2*>(pn_gamma + i * 4);
        const float2 x0 = h2_to_f2(keep[c].x), x1 = h2_to_f2(keep[c].y), g0 = h2_to_f2(g.x), g1 = h2_to_f2(g.y);
        uint2 o;
        o.x = f2_to_h2(__half2float(__float2half_rn(x0.x * rstd)) * g0.x, __half2float(__float2half_rn(x0.y * rstd)) * g0.y);
        o.y = f2_to_h2(__half2float(__float2half_rn(x1.x * rstd)) * g1.x, __half2float(__float2half_rn(x1.y * rstd)) * g1.y);
        *reinterpret_cast<uint2*>(pn_out + (long long)r * pn_ld + i * 4) = o;
      }
    }
    return;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (N >> 2); i += 256) {
    const uint2 q = *reinterpret_cast<const uint2*>(out + (long long)r * ld_out + i * 4);
    const uint2 g = *reinterpret_cast<const uint2*>(pn_gamma + i * 4);
    const float2 x0 = h2_to_f2(q.x), x1 = h2_to_f2(q.y), g0 = h2_to_f2(g.x), g1 = h2_to_f2(g.y);
    uint2 o;
    o.x = f2_to_h2(__half2float(__float2half_rn(x0.x * rstd)) * g0.x, __half2float(__float2half_rn(x0.y * rstd)) * g0.y);
    o.y = f2_to_h2(__half2float(__float2half_rn(x1.x * rstd)) * g1.x, __half2float(__float2half_rn(x1.y * rstd)) * g1.y);
    *reinterpret_cast<uint2*>(pn_out + (long long)r * pn_ld + i * 4) = o;
  }
}

// ------------------------------------------------------------------------------------------
// Arg-max over rows of fp32 logits, lowest index wins ties (greedy decoding,
// generation/utils.py greedy branch == torch.argmax).  One block per row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
argmax_kernel(const T* __restrict__ logits, long long ld, int n, long long* __restrict__ out,
              int* __restrict__ counter, int n_counters) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const T* p = logits + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = (float)p[i];
    if (v > best || bi == 0x7fffffff) {  // indices ascend within a thread: first maximum stays
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      out[blockIdx.x] = bi;
      if (counter && blockIdx.x == 0)
        for (int c = 0; c < n_counters; ++c) counter[c] += 1;
    }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int scb_im2col_patch(const void* pixels, void* out, int32_t n_img, int32_t channels, int32_t height,
                     int32_t width, int32_t patch, int64_t ld_out, void* stream) {
  SCB_REQUIRE(pixels && out, SCB_ERR_INVALID_ARG, "scb_im2col_patch: NULL buffer");
  SCB_REQUIRE(n_img >= 1 && channels >= 1 && patch >= 1 && height % patch == 0 && width % patch == 0,
              SCB_ERR_INVALID_ARG, "scb_im2col_patch: bad geometry");
  SCB_REQUIRE(ld_out >= (int64_t)channels * patch * patch, SCB_ERR_INVALID_ARG,
              "scb_im2col_patch: ld_out too small");
  const long long rows = (long long)n_img * (height / patch) * (width / patch);
  im2col_patch_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(pixels), reinterpret_cast<__half*>(out), n_img, channels,
      height, width, patch, (int)ld_out);
  SCB_LAUNCH_OK("im2col_patch_kernel");
  return SCB_OK;
}

static int norm_common(int kind, const void* x, int64_t ld_x, const void* gamma, const void* beta,
                       float eps, int64_t rows, int32_t dim, void* out, int64_t ld_out,
                       const void* cls, const void* pos, int32_t tokens_per_img, void* stream) {
  SCB_REQUIRE(x && gamma && out && rows >= 1, SCB_ERR_INVALID_ARG, "scb norm: NULL buffer / rows < 1");
  SCB_REQUIRE(dim % 8 == 0 && dim >= 8 && dim <= kNormThreads * 8 * kNormMaxChunks,
              SCB_ERR_INVALID_ARG, "scb norm: dim=%d must be a multiple of 8 in 8..%d", dim,
              kNormThreads * 8 * kNormMaxChunks);
  SCB_REQUIRE(ld_x % 8 == 0 && ld_out % 8 == 0 && aligned16(x) && aligned16(out) && aligned16(gamma) &&
                  (!beta || aligned16(beta)),
              SCB_ERR_INVALID_ARG, "scb norm: rows must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half* xh = reinterpret_cast<const __half*>(x);
  const __half* g = reinterpret_cast<const __half*>(gamma);
  const __half* b = reinterpret_cast<const __half*>(beta);
  __half* o = reinterpret_cast<__half*>(out);
  if (dim <= 32 * 8 * kWarpNormMaxChunks && rows >= 64) {
    // many narrow rows (CLIP tokens): one warp per row
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    const __half* c16 = reinterpret_cast<const __half*>(cls);
    const __half* p16 = reinterpret_cast<const __half*>(pos);
    const bool narrow = dim <= 32 * 8 * 4;
    if (kind == 0) {
      if (narrow) norm_warp_kernel<0, 4><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, nullptr, nullptr, 0);
      else norm_warp_kernel<0><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, nullptr, nullptr, 0);
    } else if (kind == 1) {
      if (narrow) norm_warp_kernel<1, 4><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, nullptr, nullptr, 0);
      else norm_warp_kernel<1><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, nullptr, nullptr, 0);
    } else {
      if (narrow) norm_warp_kernel<2, 4><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, c16, p16, tokens_per_img);
      else norm_warp_kernel<2><<<blocks, 256, 0, st>>>(xh, ld_x, g, b, eps, dim, rows, o, ld_out, c16, p16, tokens_per_img);
    }
    SCB_LAUNCH_OK("norm_warp_kernel");
    return SCB_OK;
  }
  if (kind == 0)
    norm_kernel<0><<<(unsigned)rows, kNormThreads, 0, st>>>(xh, ld_x, g, b, eps, dim, o, ld_out,
                                                            nullptr, nullptr, 0);
  else if (kind == 1)
    SCB_CUDA_OK(scb_launch_pdl(norm_kernel<1>, dim3((unsigned)rows), dim3(kNormThreads), 0, st, xh,
                               (long long)ld_x, g, b, eps, (int)dim, o, (long long)ld_out,
                               (const __half*)nullptr, (const __half*)nullptr, 0));
  else
    norm_kernel<2><<<(unsigned)rows, kNormThreads, 0, st>>>(
        xh, ld_x, g, b, eps, dim, o, ld_out, reinterpret_cast<const __half*>(cls),
        reinterpret_cast<const __half*>(pos), tokens_per_img);
  SCB_LAUNCH_OK("norm_kernel");
  return SCB_OK;
}

int scb_layernorm(const void* x, int64_t ld_x, const void* gamma, const void* beta, float eps,
                  int64_t rows, int32_t dim, void* out, int64_t ld_out, void* stream) {
  SCB_REQUIRE(beta != nullptr, SCB_ERR_INVALID_ARG, "scb_layernorm: beta is NULL");
  return norm_common(0, x, ld_x, gamma, beta, eps, rows, dim, out, ld_out, nullptr, nullptr, 0, stream);
}

int scb_rmsnorm(const void* x, int64_t ld_x, const void* weight, float eps, int64_t rows,
                int32_t dim, void* out, int64_t ld_out, void* stream) {
  return norm_common(1, x, ld_x, weight, nullptr, eps, rows, dim, out, ld_out, nullptr, nullptr, 0,
                     stream);
}

int scb_vit_embed_layernorm(const void* patch_embeds, int64_t ld_patch, const void* cls_embedding,
                            const void* pos_embedding, const void* gamma, const void* beta, float eps,
                            int32_t n_img, int32_t tokens_per_img, int32_t dim, void* out,
                            int64_t ld_out, void* stream) {
  SCB_REQUIRE(cls_embedding && pos_embedding && beta && aligned16(cls_embedding) &&
                  aligned16(pos_embedding),
              SCB_ERR_INVALID_ARG, "scb_vit_embed_layernorm: NULL / unaligned embedding tables");
  SCB_REQUIRE(n_img >= 1 && tokens_per_img >= 2, SCB_ERR_INVALID_ARG,
              "scb_vit_embed_layernorm: bad token geometry");
  return norm_common(2, patch_embeds, ld_patch, gamma, beta, eps, (int64_t)n_img * tokens_per_img, dim,
                     out, ld_out, cls_embedding, pos_embedding, tokens_per_img, stream);
}

int scb_text_embed_layernorm(const int64_t* ids, const int64_t* position_ids, int64_t n_tokens,
                             const void* word_embeddings, const void* position_embeddings,
                             const void* token_type_row, const void* gamma, const void* beta, float eps,
                             int32_t dim, void* out, int64_t ld_out, void* stream) {
  SCB_REQUIRE(ids && position_ids && word_embeddings && position_embeddings && token_type_row && gamma &&
                  beta && out && n_tokens >= 1,
              SCB_ERR_INVALID_ARG, "scb_text_embed_layernorm: NULL buffer / no tokens");
  SCB_REQUIRE(dim % 8 == 0 && dim <= kNormThreads * 8 * kNormMaxChunks && ld_out % 8 == 0 &&
                  aligned16(word_embeddings) && aligned16(position_embeddings) && aligned16(token_type_row) &&
                  aligned16(gamma) && aligned16(beta) && aligned16(out),
              SCB_ERR_INVALID_ARG, "scb_text_embed_layernorm: dim %d / alignment unsupported", dim);
  text_embed_layernorm_kernel<<<(unsigned)n_tokens, kNormThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), reinterpret_cast<const long long*>(position_ids),
      reinterpret_cast<const __half*>(word_embeddings), reinterpret_cast<const __half*>(position_embeddings),
      reinterpret_cast<const __half*>(token_type_row), reinterpret_cast<const __half*>(gamma),
      reinterpret_cast<const __half*>(beta), eps, (int)dim, reinterpret_cast<__half*>(out), (long long)ld_out);
  SCB_LAUNCH_OK("text_embed_layernorm_kernel");
  return SCB_OK;
}

int scb_avgpool2x2_tokens(const void* in, int64_t ld_in, int64_t img_stride_rows, int32_t tok_offset,
                          int32_t grid, int32_t n_img, int32_t dim, void* out, int64_t ld_out,
                          void* stream) {
  SCB_REQUIRE(in && out && n_img >= 1 && grid >= 2 && grid % 2 == 0 && dim % 8 == 0,
              SCB_ERR_INVALID_ARG, "scb_avgpool2x2_tokens: bad arguments");
  SCB_REQUIRE(ld_in % 8 == 0 && ld_out % 8 == 0 && aligned16(in) && aligned16(out), SCB_ERR_INVALID_ARG,
              "scb_avgpool2x2_tokens: rows must be 16-byte aligned");
  const long long rows = (long long)n_img * (grid / 2) * (grid / 2);
  pool2x2_kernel<<<(unsigned)rows, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(in), ld_in, img_stride_rows, tok_offset, grid, dim,
      reinterpret_cast<__half*>(out), ld_out);
  SCB_LAUNCH_OK("pool2x2_kernel");
  return SCB_OK;
}

int scb_mean_pool_normalize(const void* hidden, int64_t ld_hidden, const int32_t* offsets, int32_t n_texts,
                            int32_t dim, float* out, int64_t ld_out, int32_t normalize, void* stream) {
  SCB_REQUIRE(hidden && offsets && out && n_texts >= 1, SCB_ERR_INVALID_ARG, "scb_mean_pool_normalize: bad arguments");
  SCB_REQUIRE(dim >= 8 && dim % 8 == 0 && dim <= 2048 && ld_hidden % 8 == 0 && aligned16(hidden), SCB_ERR_INVALID_ARG,
              "scb_mean_pool_normalize: dim=%d must be a multiple of 8 up to 2048 with 16-byte aligned rows", dim);
  mean_pool_normalize_kernel<<<(unsigned)n_texts, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(hidden), ld_hidden, offsets, dim, out, ld_out, normalize);
  SCB_LAUNCH_OK("mean_pool_normalize_kernel");
  return SCB_OK;
}

int scb_silu_mul(const void* gate_up, int64_t ld_gu, int64_t rows, int32_t inter, void* out,
                 int64_t ld_out, void* stream) {
  SCB_REQUIRE(gate_up && out && rows >= 1 && inter % 8 == 0 && ld_gu % 8 == 0 && ld_out % 8 == 0 &&
                  aligned16(gate_up) && aligned16(out),
              SCB_ERR_INVALID_ARG, "scb_silu_mul: bad arguments");
  const long long work = rows * (inter / 8);
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)scb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  silu_mul_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(gate_up), ld_gu, inter, reinterpret_cast<__half*>(out), ld_out,
      rows);
  SCB_LAUNCH_OK("silu_mul_kernel");
  return SCB_OK;
}

int scb_rope_append(void* qkv, int64_t ld_qkv, int64_t rows, int32_t n_q_heads, int32_t n_kv_heads,
                    int32_t head_dim, const int32_t* positions, int32_t pos_base, float theta,
                    const float* inv_freq, void* k_cache, void* v_cache, int64_t max_pos,
                    void* stream) {
  SCB_REQUIRE(qkv && rows >= 1 && n_q_heads >= 1 && n_kv_heads >= 0 && head_dim % 2 == 0,
              SCB_ERR_INVALID_ARG, "scb_rope_append: bad arguments");
  SCB_REQUIRE((k_cache == nullptr) == (v_cache == nullptr), SCB_ERR_INVALID_ARG,
              "scb_rope_append: k_cache and v_cache must both be given or both be NULL");
  SCB_REQUIRE(!k_cache || positions || pos_base + rows <= max_pos, SCB_ERR_INVALID_ARG,
              "scb_rope_append: positions exceed the cache (%lld + %lld > %lld)", (long long)pos_base,
              (long long)rows, (long long)max_pos);
  if (head_dim % 16 == 0 && head_dim <= 256 && ld_qkv % 8 == 0 && aligned16(qkv) &&
      (!k_cache || (aligned16(k_cache) && aligned16(v_cache))))
    rope_append_vec_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__half*>(qkv), ld_qkv, n_q_heads, n_kv_heads, head_dim, positions, pos_base,
        theta, inv_freq, reinterpret_cast<__half*>(k_cache), reinterpret_cast<__half*>(v_cache), max_pos);
  else
    rope_append_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__half*>(qkv), ld_qkv, n_q_heads, n_kv_heads, head_dim, positions, pos_base,
        theta, inv_freq, reinterpret_cast<__half*>(k_cache), reinterpret_cast<__half*>(v_cache), max_pos,
        rows);
  SCB_LAUNCH_OK("rope_append_kernel");
  return SCB_OK;
}

int scb_gather_rows(const void* table, const int64_t* ids, int64_t n_ids, int32_t dim, void* out,
                    int64_t ld_out, void* stream) {
  SCB_REQUIRE(table && ids && out && n_ids >= 1 && dim % 8 == 0 && ld_out % 8 == 0 &&
                  aligned16(table) && aligned16(out),
              SCB_ERR_INVALID_ARG, "scb_gather_rows: bad arguments");
  SCB_CUDA_OK(scb_launch_pdl(gather_rows_kernel, dim3((unsigned)n_ids), dim3(128), 0,
                             reinterpret_cast<cudaStream_t>(stream),
                             reinterpret_cast<const __half*>(table),
                             reinterpret_cast<const long long*>(ids), (int)dim,
                             reinterpret_cast<__half*>(out), (long long)ld_out));
  SCB_LAUNCH_OK("gather_rows_kernel");
  return SCB_OK;
}

int scb_gather_rows_rmsnorm(const void* table, const int64_t* ids, int64_t n_ids, int32_t dim, void* out,
                            int64_t ld_out, const void* gamma, float eps, void* norm_out,
                            int64_t ld_norm, void* stream) {
  SCB_REQUIRE(table && ids && out && gamma && norm_out && n_ids >= 1 && dim % 8 == 0 && ld_out % 8 == 0 &&
                  ld_norm % 8 == 0 && dim <= kNormThreads * 8 * kNormMaxChunks && aligned16(table) &&
                  aligned16(out) && aligned16(gamma) && aligned16(norm_out),
              SCB_ERR_INVALID_ARG, "scb_gather_rows_rmsnorm: bad arguments");
  SCB_CUDA_OK(scb_launch_pdl(gather_rmsnorm_kernel, dim3((unsigned)n_ids), dim3(kNormThreads), 0,
                             reinterpret_cast<cudaStream_t>(stream),
                             reinterpret_cast<const __half*>(table),
                             reinterpret_cast<const long long*>(ids), (int)dim,
                             reinterpret_cast<__half*>(out), (long long)ld_out,
                             reinterpret_cast<const __half*>(gamma), eps,
                             reinterpret_cast<__half*>(norm_out), (long long)ld_norm));
  SCB_LAUNCH_OK("gather_rmsnorm_kernel");
  return SCB_OK;
}

int scb_splitk_reduce(const float* partial, int64_t split_stride, int64_t ld_partial, int32_t n_splits,
                      int64_t rows, int32_t n, const void* bias, const void* residual, int64_t ld_res,
                      void* out, int64_t ld_out, const void* post_norm_gamma, void* post_norm_out,
                      int64_t post_norm_ld, float post_norm_eps, void* stream) {
  SCB_REQUIRE(partial && out && rows >= 1 && n >= 4 && n % 4 == 0 && n_splits >= 1, SCB_ERR_INVALID_ARG,
              "scb_splitk_reduce: bad arguments");
  SCB_REQUIRE(ld_partial % 4 == 0 && split_stride % 4 == 0 && ld_out % 4 == 0 && (!residual || ld_res % 4 == 0) &&
                  (!post_norm_out || (post_norm_gamma && post_norm_ld % 4 == 0)),
              SCB_ERR_INVALID_ARG, "scb_splitk_reduce: rows must be 8-byte aligned");
  SCB_CUDA_OK(scb_launch_pdl(splitk_reduce_kernel, dim3((unsigned)rows), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), partial, (long long)split_stride,
                             (long long)ld_partial, (int)n_splits, (int)n,
                             reinterpret_cast<const __half*>(bias),
                             reinterpret_cast<const __half*>(residual), (long long)ld_res,
                             reinterpret_cast<__half*>(out), (long long)ld_out,
                             reinterpret_cast<const __half*>(post_norm_gamma),
                             reinterpret_cast<__half*>(post_norm_out), (long long)post_norm_ld,
                             post_norm_eps));
  SCB_LAUNCH_OK("splitk_reduce_kernel");
  return SCB_OK;
}

int scb_argmax_rows(const void* logits, int32_t dtype, int64_t ld, int64_t rows, int32_t n,
                    int64_t* out, int32_t* counters, int32_t n_counters, void* stream) {
  SCB_REQUIRE(logits && out && rows >= 1 && n >= 1, SCB_ERR_INVALID_ARG, "scb_argmax_rows: bad arguments");
  SCB_REQUIRE(dtype == SCB_F32 || dtype == SCB_F16, SCB_ERR_UNSUPPORTED, "scb_argmax_rows: dtype %d", dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SCB_F32)
    argmax_kernel<float><<<(unsigned)rows, 1024, 0, st>>>(reinterpret_cast<const float*>(logits), ld, n,
                                                          reinterpret_cast<long long*>(out), counters,
                                                          n_counters);
  else
    argmax_kernel<__half><<<(unsigned)rows, 1024, 0, st>>>(reinterpret_cast<const __half*>(logits), ld,
                                                           n, reinterpret_cast<long long*>(out), counters,
                                                          n_counters);
  SCB_LAUNCH_OK("argmax_kernel");
  return SCB_OK;
}

}  // extern "C"
