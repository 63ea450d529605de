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
#include <stdlib.h>

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
  int silu_pair;  // rows (2i, 2i+1) of w are (gate_i, up_i): out[r][i] = fp16(silu(gate)) * up
  float eps;
  // fused RMSNorm of the finished output rows by the last block to leave the kernel (decode chain:
  // the next layer's input norm), and an L2 prefetch of the next kernel's weights issued by blocks
  // that ran out of work (fills the launch gap / ramp-up of the next kernel with useful traffic)
  const __half* pn_gamma;
  __half* pn_out;
  long long pn_ld;
  float pn_eps;
  unsigned* done_ticket;
  const char* next_w;
  long long next_w_bytes;
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
// gemv_mma_kernel — the decode linear layer for 1..8 activation rows as a streaming kernel that
// uses the warp-level tensor-core MMA (m16n8k16, fp16 in / fp32 accumulate) purely as a
// dot-product engine: the kernel is HBM-bound, the point of the MMA is that the k-reduction
// happens inside the instruction (no shuffle tree) and that 1..8 activation rows cost the same.
//
//   block  = W warps working on one group of 8 weight rows at a time (grid-stride over groups);
//            the K range is split evenly over the W warps, partial sums meet in shared memory in
//            a fixed order (deterministic).
//   lane   (g = lane/4, t = lane%4) loads 16 B of weight row n0+g at k = tile*32 + t*8: a warp
//            load covers 8 rows x 64 contiguous bytes (full 32-byte sectors).  The 8 halfs feed two
//            MMAs as the B fragment; the matching A fragment is ONE 16-byte shared-memory read of
//            activation row g at the same k offset (rows g >= R read as zero; MMA rows 8..15 are
//            zero registers), i.e. the logical k order inside an MMA is a fixed permutation of the
//            physical one, identical for A and B.
//   loads  U = 8 tiles (4 KB per warp) are issued before the previous group is consumed; the
//            first group is requested BEFORE the activation prologue, so the RMSNorm / SiLU-gate
//            staging overlaps the weight stream.
//   lm_head: optional fused arg-max (block-local best, last block reduces; lowest index wins).
// ------------------------------------------------------------------------------------------
constexpr int kMU = 8;             // k-tiles (32 halfs each) in flight per warp: 4 KB
constexpr int kMaxMmaWarps = 16;

SCB_DEVICE void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                         uint32_t b1) {
  // a0/a2: activation row g (k low / high half), a1/a3: activation row g+8 (zero when rows <= 8)
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct ArgmaxFuse {
  float* blk_val;          // [gridDim.x * R]
  int* blk_idx;            // [gridDim.x * R]
  unsigned* ticket;        // [1], zero on entry, reset on exit
  long long* out_ids;      // [R]
  int* counters;           // optional device ints to increment (graph bookkeeping)
  int n_counters;
};

// Work unit = (group of 8 weight rows, K chunk).  Skinny layers (few row groups, long K: o_proj,
// down_proj, qkv) are cut into several K chunks so that one wave of 8 blocks x 148 SMs is full;
// the chunks of a group meet through a partial buffer and an atomic ticket — the last block to
// arrive sums them in chunk order (deterministic) and runs the epilogue.
// kHi: activation rows 8..15 are live (batches of 9..16 sequences); costs a few registers, so the
// <= 8 row variant keeps the 64-register / 32-warps-per-SM budget and the wide one runs 24 warps.
template <int kGW, bool kHi>
__global__ void __launch_bounds__(kGW * 32, kHi ? (24 / kGW > 0 ? 24 / kGW : 1) : 32 / kGW)
gemv_mma_kernel(const GemvParams p, const int R, const int C, float* __restrict__ partial,
                unsigned* __restrict__ gtickets, const ArgmaxFuse am) {
  __shared__ float red[kGW][16][8];
  __shared__ unsigned flag_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_groups = (p.N + 7) >> 3;
  const int n_units = n_groups * C;
  const int ktiles = p.K >> 5;

  // tile range of (chunk c, this warp): chunk c covers [c*T/C, (c+1)*T/C), split over the warps
  auto tile_range = [&](int c, int& b, int& e) {
    const int cb = (int)((long long)ktiles * c / C), ce = (int)((long long)ktiles * (c + 1) / C);
    const int per = (ce - cb + kGW - 1) / kGW;
    b = cb + warp * per;
    e = min(ce, b + per);
  };
  auto issue = [&](uint4 (&w)[kMU], int unit, int kt0, int kt_end) {
    const int row = (unit / C) * 8 + g;
    const bool row_ok = unit < n_units && row < p.N;
    const __half* wr = p.w + (long long)row * p.ldw + t * 8;
#pragma unroll
    for (int u = 0; u < kMU; ++u) {
      const int kt = kt0 + u;
      w[u] = (row_ok && kt < kt_end) ? ld_stream_u4(wr + kt * 32) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 wv[kMU];
  int unit = blockIdx.x;
  int kb = 0, ke = 0;
  if (unit < n_units) tile_range(unit % C, kb, ke);
  issue(wv, unit, kb, ke);    // weights do not depend on the previous kernel: request them first
  pdl_launch_dependents();
  pdl_wait();                 // activations (and everything we write) belong to the dependency chain

  const bool a_ok = g < R, a_ok2 = kHi && (g + 8 < R);   // rows 8..15 of the MMA carry sequences 8..15
  const __half* a_row = p.a + (long long)(a_ok ? g : 0) * p.lda + t * 8;   // read through L1 (tiny, shared)
  const __half* a_row2 = p.a + (long long)(a_ok2 ? g + 8 : 0) * p.lda + t * 8;
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;
  while (unit < n_units) {
    const int grp = unit / C, c = unit - grp * C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kt0 = kb; kt0 < ke; kt0 += kMU) {
      if (kt0 != kb) issue(wv, unit, kt0, ke);
#pragma unroll
      for (int u = 0; u < kMU; ++u) {
        const int kt = kt0 + u;
        uint4 av = make_uint4(0, 0, 0, 0), av2 = make_uint4(0, 0, 0, 0);
        if (a_ok && kt < ke) av = __ldg(reinterpret_cast<const uint4*>(a_row + kt * 32));
        if constexpr (kHi)
          if (a_ok2 && kt < ke) av2 = __ldg(reinterpret_cast<const uint4*>(a_row2 + kt * 32));
        mma16816(acc, av.x, av2.x, av.y, av2.y, wv[u].x, wv[u].y);
        mma16816(acc, av.z, av2.z, av.w, av2.w, wv[u].z, wv[u].w);
      }
    }
    const int nunit = unit + gridDim.x;
    int nkb = 0, nke = 0;
    if (nunit < n_units) tile_range(nunit % C, nkb, nke);
    issue(wv, nunit, nkb, nke);   // next unit's first batch flies during the combine
    // ---- combine the warps in order: acc[0], acc[1] = out[row g][n = grp*8 + t*2 + {0,1}]
    red[warp][g][t * 2] = acc[0];
    red[warp][g][t * 2 + 1] = acc[1];
    red[warp][g + 8][t * 2] = acc[2];
    red[warp][g + 8][t * 2 + 1] = acc[3];
    __syncthreads();
    const int r = tid >> 3, j = tid & 7;  // threads 0..127: (row r of 16, column j of the 8-row group)
    const int n = grp * 8 + j;
    float v = 0.f;
    bool have = false;                    // this thread holds a finished output value
    if (tid < 128 && r < R) {
#pragma unroll
      for (int w = 0; w < kGW; ++w) v += red[w][r][j];
      have = (C == 1);
    }
    if (C > 1) {
      if (tid < 128 && r < R) partial[((long long)grp * C + c) * 128 + tid] = v;
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const unsigned tk = atomicAdd(&gtickets[grp], 1u);
        flag_s = (tk == (unsigned)C - 1u) ? 1u : 0u;
      }
      __syncthreads();
      if (flag_s) {
        __threadfence();
        if (tid < 128 && r < R) {
          v = 0.f;
          for (int cc = 0; cc < C; ++cc) v += __ldcg(partial + ((long long)grp * C + cc) * 128 + tid);
          have = true;
        }
        if (tid == 0) gtickets[grp] = 0;
      }
    }
    // ---- epilogue (warp-uniform branch on silu_pair; `have` only where a value is final)
    if (p.silu_pair) {
      if (tid < 128) {
        const float gt = __half2float(__float2half_rn(v));       // Linear output is fp16 in HF
        const float up = __shfl_xor_sync(0xffffffffu, gt, 1);
        if (have && n < p.N && (j & 1) == 0) {
          const float sg = __half2float(__float2half_rn(__fdividef(gt, 1.f + __expf(-gt))));
          reinterpret_cast<__half*>(p.out)[(long long)r * p.ldo + (n >> 1)] = __float2half_rn(sg * up);
        }
      }
    } else if (have && n < p.N) {
      if (p.bias) v += __half2float(p.bias[n]);
      if (p.residual) v += __half2float(p.residual[(long long)r * p.ldr + n]);
      if (p.out_f32) {
        reinterpret_cast<float*>(p.out)[(long long)r * p.ldo + n] = v;
      } else {
        const __half hv = __float2half_rn(v);
        reinterpret_cast<__half*>(p.out)[(long long)r * p.ldo + n] = hv;
        v = __half2float(hv);          // arg-max sees the stored (rounded) logits
      }
      if (am.out_ids && (v > best_v || (v == best_v && n < best_i))) {
        best_v = v;
        best_i = n;
      }
    }
    __syncthreads();   // red / flag_s reuse
    unit = nunit;
    kb = nkb;
    ke = nke;
  }

  // ---- tail 1: blocks that ran out of work pull the next kernel's weights into L2
  if (p.next_w) {
    const long long per_blk = (p.next_w_bytes / gridDim.x) & ~127ll;
    const char* base = p.next_w + (long long)blockIdx.x * per_blk;
    for (long long off = (long long)tid * 128; off < per_blk; off += (long long)kGW * 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }
  // ---- tail 2: the last block to finish normalises the complete output rows (RMSNorm, HF order)
  if (p.pn_out) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned tk = atomicAdd(p.done_ticket, 1u);
      flag_s = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (flag_s) {
      __threadfence();
      float* sred = &red[0][0][0];
      const __half* o16 = reinterpret_cast<const __half*>(p.out);
      for (int r = 0; r < R; ++r) {
        float ss = 0.f;
        for (int i = tid; i < (p.N >> 3); i += kGW * 32) {
          float f[8];
          const uint4 q = __ldcg(reinterpret_cast<const uint4*>(o16 + (long long)r * p.ldo + i * 8));
          unpack8(q, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
        }
        ss = warp_sum(ss);
        __syncthreads();
        if (lane == 0) sred[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kGW; ++w) tot += sred[w];
        const float rstd = rsqrtf(tot / (float)p.N + p.pn_eps);
        for (int i = tid; i < (p.N >> 3); i += kGW * 32) {
          float f[8], gm[8], o[8];
          unpack8(__ldcg(reinterpret_cast<const uint4*>(o16 + (long long)r * p.ldo + i * 8)), f);
          unpack8(*reinterpret_cast<const uint4*>(p.pn_gamma + i * 8), gm);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __half2float(__float2half_rn(f[e] * rstd)) * gm[e];
          uint4 q;
          q.x = f2_to_h2(o[0], o[1]); q.y = f2_to_h2(o[2], o[3]);
          q.z = f2_to_h2(o[4], o[5]); q.w = f2_to_h2(o[6], o[7]);
          *reinterpret_cast<uint4*>(p.pn_out + (long long)r * p.pn_ld + i * 8) = q;
        }
      }
      if (tid == 0) *p.done_ticket = 0;
    }
    __syncthreads();
  }

  if (am.out_ids) {
    // ---- fused arg-max: thread (r, j) holds the best over its column residue; reduce the 8
    // threads of each row, publish per block, the last block reduces the blocks
    float bv = best_v;
    int bi = best_i;
    if (tid < 128) {  // R <= 16 rows x 8 threads, aligned groups of 8 lanes
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (tid < R * 8 && (tid & 7) == 0) {
        am.blk_val[(long long)blockIdx.x * R + (tid >> 3)] = bv;
        am.blk_idx[(long long)blockIdx.x * R + (tid >> 3)] = bi;
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned tk = atomicAdd(am.ticket, 1u);
      flag_s = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (flag_s) {
      __threadfence();
      for (int row = warp; row < R; row += kGW) {
        float bv2 = -INFINITY;
        int ix = 0x7fffffff;
        for (int b = lane; b < (int)gridDim.x; b += 32) {
          const float ov = __ldcg(am.blk_val + (long long)b * R + row);
          const int oi = __ldcg(am.blk_idx + (long long)b * R + row);
          if (ov > bv2 || (ov == bv2 && oi < ix)) {
            bv2 = ov;
            ix = oi;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv2, o);
          const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
          if (ov > bv2 || (ov == bv2 && oi < ix)) {
            bv2 = ov;
            ix = oi;
          }
        }
        if (lane == 0) am.out_ids[row] = ix;
      }
      if (tid == 0) {
        *am.ticket = 0;
        for (int cc = 0; cc < am.n_counters; ++cc) am.counters[cc] += 1;
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

// ------------------------------------------------------------------------------------------
// decode_attn_fused_kernel — one launch per layer for the whole attention block of a decode step:
// RoPE of the new q and k (same fp16 rounding as rope_append_kernel), append of k/v to the cache,
// split-KV attention for all q heads of a kv group, and the cross-split combine by the last
// block of each kv head (atomic ticket).  Context length comes from device memory.
//   phase 1  thread = cached position: dot of the 256-byte K row with the G rotated q rows
//            (shared-memory broadcast reads), scores to shared memory        -> reads K once
//   phase 2  warp = q head: max / exp2 / sum over the block's positions
//   phase 3  thread = (dim, position half): P·V with coalesced V reads        -> reads V once
// ------------------------------------------------------------------------------------------
template <int HD, int G>
__global__ void __launch_bounds__(256, 2)
decode_attn_fused_kernel(const __half* __restrict__ qkv, const float* __restrict__ inv_freq,
                         float theta, __half* __restrict__ k_cache, __half* __restrict__ v_cache,
                         long long max_pos, const int* __restrict__ len_dev, float scale_log2,
                         int n_q, int n_kv, int per_max, float* __restrict__ part_m,
                         float* __restrict__ part_l, float* __restrict__ part_o,
                         unsigned* __restrict__ tickets, __half* __restrict__ out, long long ld_qkv,
                         long long cache_seq_stride, const char* __restrict__ next_w,
                         long long next_w_bytes) {
  constexpr int HALF = HD / 2;
  extern __shared__ __align__(16) float dsm[];
  __half* q_h = reinterpret_cast<__half*>(dsm);   // [8][HD] rotated q (fp16, rows >= G zero): the mma B operand
  float* s_s = dsm + 4 * HD;        // [G][per_max] scores (x scale*log2 e) -> exp2 values; per_max = 4 mod 16
  __shared__ float m_s[G], l_s[G];
  __shared__ float m_sp[G * 64], l_sp[G * 64];   // per-split statistics for the combine (n_splits <= 64)
  __shared__ float o_s[8][G][HD];   // per-warp partial outputs
  __shared__ unsigned last_flag;
  const int hk = blockIdx.x, split = blockIdx.y, n_splits = gridDim.y;
  const int seq = blockIdx.z;       // independent sequences (batched captions): own row, cache, length
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  qkv += (long long)seq * ld_qkv;
  k_cache += (long long)seq * cache_seq_stride;
  v_cache += (long long)seq * cache_seq_stride;
  out += (long long)seq * n_q * HD;
  part_m += (long long)seq * n_q * n_splits;
  part_l += (long long)seq * n_q * n_splits;
  part_o += (long long)seq * n_q * n_splits * HD;
  tickets += seq * n_kv;
  // the host validates lengths; this guard keeps a bad one from writing outside the cache slot (length
  // clamped, append skipped)
  const int kv_raw = len_dev[seq];
  const bool kv_ok = kv_raw >= 1 && kv_raw <= max_pos;
  const int kv_len = max(1, min(kv_raw, (int)max_pos));
  const int pos = kv_len - 1;       // position of the token being decoded
  const int per = (kv_len + n_splits - 1) / n_splits;
  const int begin = split * per, end = min(kv_len, begin + per);
  const int n_local = min(per_max, max(0, end - begin));   // per_max bounds the score buffer (host guarantees kv_len <= ctx_cap)

  // ---- cos/sin of the new position once per block (HALF angles), then RoPE of the G q rows
  __shared__ float cs_s[HALF], sn_s[HALF];
  if (tid < HALF) {
    const float f = inv_freq ? inv_freq[tid] : powf(theta, -(float)(2 * tid) / (float)HD);
    const float ang = (float)pos * f;
    cs_s[tid] = __half2float(__float2half_rn(cosf(ang)));
    sn_s[tid] = __half2float(__float2half_rn(sinf(ang)));
  }
  __syncthreads();
  for (int idx = tid; idx < G * HALF; idx += 256) {
    const int h = idx / HALF, i = idx % HALF;
    const float c = cs_s[i], sn = sn_s[i];
    const __half* qp = qkv + (long long)(hk * G + h) * HD;
    const float x1 = __half2float(qp[i]), x2 = __half2float(qp[i + HALF]);
    const float a1 = __half2float(__float2half_rn(x1 * c)), b1 = __half2float(__float2half_rn(-x2 * sn));
    const float a2 = __half2float(__float2half_rn(x2 * c)), b2 = __half2float(__float2half_rn(x1 * sn));
    q_h[h * HD + i] = __float2half_rn(a1 + b1);
    q_h[h * HD + i + HALF] = __float2half_rn(a2 + b2);
  }
  for (int idx = G * HD + tid; idx < 8 * HD; idx += 256) q_h[idx] = __float2half_rn(0.f);   // unused mma columns
  // ---- the block that owns `pos` rotates and appends the new k, copies the new v
  if (kv_ok && pos >= begin && pos < end) {
    const __half* kp = qkv + (long long)(n_q + hk) * HD;
    const __half* vp = qkv + (long long)(n_q + n_kv + hk) * HD;
    __half* kc = k_cache + ((long long)hk * max_pos + pos) * HD;
    __half* vc = v_cache + ((long long)hk * max_pos + pos) * HD;
    for (int i = tid; i < HALF; i += 256) {
      const float c = cs_s[i], sn = sn_s[i];
      const float x1 = __half2float(kp[i]), x2 = __half2float(kp[i + HALF]);
      const float a1 = __half2float(__float2half_rn(x1 * c)), b1 = __half2float(__float2half_rn(-x2 * sn));
      const float a2 = __half2float(__float2half_rn(x2 * c)), b2 = __half2float(__float2half_rn(x1 * sn));
      kc[i] = __float2half_rn(a1 + b1);
      kc[i + HALF] = __float2half_rn(a2 + b2);
    }
    for (int i = tid; i < HD; i += 256) vc[i] = vp[i];
  }
  __syncthreads();   // q_s ready; the appended row is visible to this block

  // ---- phase 1: scores on the tensor pipe.  A warp owns 16 positions per tile: the K rows are the A
  // operand of mma.m16n8k16 (M = positions, K = dims, N = the group's q heads), loaded straight from
  // the cache as 16-byte pieces — the dot product does not care which dims share a k-slot, so thread
  // (g,t) feeds dims [32c+8t, 32c+8t+8) of rows g / g+8 and the matching 16 bytes of q.  8 x 16 B in
  // flight per thread.
  {
    constexpr int NC = HD / 32;
    const int g = lane >> 2, t = lane & 3;
    uint4 qf[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) qf[c] = *reinterpret_cast<const uint4*>(q_h + g * HD + c * 32 + t * 8);
    const __half* kb = k_cache + ((long long)hk * max_pos + begin) * HD + t * 8;
    for (int p0 = warp * 16; p0 < n_local; p0 += 128) {
      const int r0 = p0 + g, r1 = r0 + 8;
      const bool ok0 = r0 < n_local, ok1 = r1 < n_local;
      uint4 k0[NC], k1[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        k0[c] = ok0 ? *reinterpret_cast<const uint4*>(kb + (long long)r0 * HD + c * 32) : make_uint4(0, 0, 0, 0);
        k1[c] = ok1 ? *reinterpret_cast<const uint4*>(kb + (long long)r1 * HD + c * 32) : make_uint4(0, 0, 0, 0);
      }
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        mma16816(d0, k0[c].x, k1[c].x, k0[c].y, k1[c].y, qf[c].x, qf[c].y);
        mma16816(d1, k0[c].z, k1[c].z, k0[c].w, k1[c].w, qf[c].z, qf[c].w);
      }
      const int h0 = 2 * t;
      if (h0 < G) {
        if (ok0) s_s[h0 * per_max + r0] = (d0[0] + d1[0]) * scale_log2;
        if (ok1) s_s[h0 * per_max + r1] = (d0[2] + d1[2]) * scale_log2;
      }
      if (h0 + 1 < G) {
        if (ok0) s_s[(h0 + 1) * per_max + r0] = (d0[1] + d1[1]) * scale_log2;
        if (ok1) s_s[(h0 + 1) * per_max + r1] = (d0[3] + d1[3]) * scale_log2;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: per-head softmax statistics (warp h)
  if (warp < G) {
    float mx = -INFINITY;
    for (int pl = lane; pl < n_local; pl += 32) mx = fmaxf(mx, s_s[warp * per_max + pl]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int pl = lane; pl < n_local; pl += 32) {
      const float e = exp2f(s_s[warp * per_max + pl] - mx);
      s_s[warp * per_max + pl] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      m_s[warp] = mx;
      l_s[warp] = sum;
    }
  }
  __syncthreads();

  // ---- phase 3: P·V on the tensor pipe, O^T = V^T · P^T: M = dims, K = positions, N = q heads.
  // Thread (g,t) loads 16 bytes (dims 8g..8g+7) of positions 2t, 2t+1, 2t+8, 2t+9 of the tile; a byte
  // permute of two such rows yields the (dim, position-pair) A registers — rows g / g+8 of m-tile j are
  // dims 8g+2j / 8g+2j+1.  The probabilities are rounded to fp16 for the B operand (as the reference's
  // fp16 softmax output is); accumulation is fp32.  8 x 16 B in flight per thread.
  {
    constexpr int NH = HD / 64;
    const int g = lane >> 2, t = lane & 3;
    float acc[NH][4][4];
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[hh][j][e] = 0.f;
    const __half* vb = v_cache + ((long long)hk * max_pos + begin) * HD + g * 8;
    const float* er = s_s + (g < G ? g : 0) * per_max;
    for (int p0 = warp * 16; p0 < n_local; p0 += 128) {
      const int pa = p0 + 2 * t, pb = pa + 8;
      const bool oka0 = pa < n_local, oka1 = pa + 1 < n_local, okb0 = pb < n_local, okb1 = pb + 1 < n_local;
      uint4 v[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const __half* vp = vb + hh * 64;
        v[hh][0] = oka0 ? *reinterpret_cast<const uint4*>(vp + (long long)pa * HD) : make_uint4(0, 0, 0, 0);
        v[hh][1] = oka1 ? *reinterpret_cast<const uint4*>(vp + (long long)(pa + 1) * HD) : make_uint4(0, 0, 0, 0);
        v[hh][2] = okb0 ? *reinterpret_cast<const uint4*>(vp + (long long)pb * HD) : make_uint4(0, 0, 0, 0);
        v[hh][3] = okb1 ? *reinterpret_cast<const uint4*>(vp + (long long)(pb + 1) * HD) : make_uint4(0, 0, 0, 0);
      }
      float2 ea = make_float2(0.f, 0.f), eb = make_float2(0.f, 0.f);
      if (g < G) {
        if (oka0) ea = *reinterpret_cast<const float2*>(er + pa);
        if (okb0) eb = *reinterpret_cast<const float2*>(er + pb);
        if (!oka1) ea.y = 0.f;
        if (!okb1) eb.y = 0.f;
      }
      const uint32_t b0 = f2_to_h2(ea.x, ea.y), b1 = f2_to_h2(eb.x, eb.y);
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const uint32_t w0[4] = {v[hh][0].x, v[hh][0].y, v[hh][0].z, v[hh][0].w};
        const uint32_t w1[4] = {v[hh][1].x, v[hh][1].y, v[hh][1].z, v[hh][1].w};
        const uint32_t w2[4] = {v[hh][2].x, v[hh][2].y, v[hh][2].z, v[hh][2].w};
        const uint32_t w3[4] = {v[hh][3].x, v[hh][3].y, v[hh][3].z, v[hh][3].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma16816(acc[hh][j], __byte_perm(w0[j], w1[j], 0x5410), __byte_perm(w0[j], w1[j], 0x7632),
                   __byte_perm(w2[j], w3[j], 0x5410), __byte_perm(w2[j], w3[j], 0x7632), b0, b1);
      }
    }
    const int h0 = 2 * t;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d0 = hh * 64 + g * 8 + 2 * j;
        if (h0 < G) {
          o_s[warp][h0][d0] = acc[hh][j][0];
          o_s[warp][h0][d0 + 1] = acc[hh][j][2];
        }
        if (h0 + 1 < G) {
          o_s[warp][h0 + 1][d0] = acc[hh][j][1];
          o_s[warp][h0 + 1][d0 + 1] = acc[hh][j][3];
        }
      }
  }
  __syncthreads();
  for (int idx = tid; idx < G * HD; idx += 256) {
    const int h = idx / HD, d = idx % HD;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += o_s[w][h][d];
    const long long pi = (long long)(hk * G + h) * n_splits + split;
    part_o[pi * HD + d] = v;
    if (d == 0) {
      part_m[pi] = m_s[h];
      part_l[pi] = l_s[h];
    }
  }
  // ---- last block of this kv head combines the splits
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned tk = atomicAdd(&tickets[hk], 1u);
    last_flag = (tk == (unsigned)n_splits - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!last_flag) return;
  __threadfence();
  // stage the per-split statistics of the G heads (one round trip), turn them into weights
  float* w_s = s_s;                      // [G][n_splits] reuse of the score buffer (n_splits <= per_max*...)
  __shared__ float inv_l[G];
  for (int idx = tid; idx < G * n_splits; idx += 256) {
    const int h = idx / n_splits, sp = idx % n_splits;
    const long long base = (long long)(hk * G + h) * n_splits + sp;
    m_sp[idx] = __ldcg(part_m + base);
    l_sp[idx] = __ldcg(part_l + base);
  }
  __syncthreads();
  if (warp < G) {
    float mm = -INFINITY;
    for (int sp = lane; sp < n_splits; sp += 32) mm = fmaxf(mm, m_sp[warp * n_splits + sp]);
    mm = warp_max(mm);
    float ll = 0.f;
    for (int sp = lane; sp < n_splits; sp += 32) {
      const float pm = m_sp[warp * n_splits + sp];
      const float scl = (pm == -INFINITY) ? 0.f : exp2f(pm - mm);
      w_s[warp * n_splits + sp] = scl;
      ll = fmaf(l_sp[warp * n_splits + sp], scl, ll);
    }
    ll = warp_sum(ll);
    if (lane == 0) inv_l[warp] = ll > 0.f ? 1.f / ll : 0.f;
  }
  __syncthreads();
  for (int idx = tid; idx < G * HD; idx += 256) {
    const int h = idx / HD, d = idx % HD;
    const float* po = part_o + ((long long)(hk * G + h) * n_splits) * HD + d;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    int sp = 0;
    for (; sp + 8 <= n_splits; sp += 8) {    // eight independent L2 loads in flight
      float a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = __ldcg(po + (long long)(sp + u) * HD);
#pragma unroll
      for (int u = 0; u < 8; u += 4) {
        o0 = fmaf(a[u + 0], w_s[h * n_splits + sp + u + 0], o0);
        o1 = fmaf(a[u + 1], w_s[h * n_splits + sp + u + 1], o1);
        o2 = fmaf(a[u + 2], w_s[h * n_splits + sp + u + 2], o2);
        o3 = fmaf(a[u + 3], w_s[h * n_splits + sp + u + 3], o3);
      }
    }
    for (; sp < n_splits; ++sp) o0 = fmaf(__ldcg(po + (long long)sp * HD), w_s[h * n_splits + sp], o0);
    out[(long long)(hk * G + h) * HD + d] = __float2half_rn(((o0 + o1) + (o2 + o3)) * inv_l[h]);
  }
  if (tid == 0) tickets[hk] = 0;
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

// workspace layout: [0,256) arg-max ticket | per-block arg-max (value,index) for 2048 blocks x 8 rows
//                   | group tickets (kMaxSplitGroups u32) | K-chunk partials
constexpr int kMaxSplitGroups = 4096;
constexpr int kMaxChunks = 8;
constexpr size_t kWsArgmax = 256 + (size_t)2048 * 16 * (sizeof(float) + sizeof(int));
constexpr size_t kWsTickets = (size_t)kMaxSplitGroups * sizeof(unsigned);
constexpr int kGWDefault = 8;   // sweep on B200 (tools/gemv_sweep.py): 8 warps per block, no K chunking
constexpr size_t kWsPartial = (size_t)kMaxSplitGroups * kMaxChunks * 128 * sizeof(float);

size_t scb_gemv_workspace_bytes(int32_t rows) {
  if (rows < 1 || rows > 16) return 0;
  return kWsArgmax + kWsTickets + kWsPartial;
}

int scb_gemv(const scb_gemv_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_gemv: args is NULL");
  SCB_REQUIRE(a->a && a->w && a->out, SCB_ERR_INVALID_ARG, "scb_gemv: NULL buffer");
  SCB_REQUIRE(a->rows >= 1 && a->rows <= 16, SCB_ERR_UNSUPPORTED,
              "scb_gemv: rows=%d (1..16; larger batches use scb_linear)", a->rows);
  SCB_REQUIRE(a->n >= 1 && a->k >= 8 && a->k % 8 == 0, SCB_ERR_INVALID_ARG,
              "scb_gemv: n=%lld k=%lld (k must be a multiple of 8)", (long long)a->n, (long long)a->k);
  SCB_REQUIRE(a->prologue >= PRO_NONE && a->prologue <= PRO_SILU_MUL, SCB_ERR_UNSUPPORTED,
              "scb_gemv: prologue %d unsupported", a->prologue);
  SCB_REQUIRE(a->prologue != PRO_RMSNORM || a->gamma, SCB_ERR_INVALID_ARG,
              "scb_gemv: RMSNorm prologue needs gamma");
  const long long min_lda = a->prologue == PRO_SILU_MUL ? 2 * a->k : a->k;
  SCB_REQUIRE(a->lda >= min_lda && a->ldw >= a->k && (a->ldo >= a->n || a->epilogue == 1) && a->lda % 8 == 0 &&
                  a->ldw % 8 == 0,
              SCB_ERR_INVALID_ARG, "scb_gemv: bad leading dimensions");
  SCB_REQUIRE(al16(a->a) && al16(a->w) && (!a->gamma || al16(a->gamma)), SCB_ERR_INVALID_ARG,
              "scb_gemv: a / w / gamma must be 16-byte aligned");
  SCB_REQUIRE(!a->residual || a->ldr >= a->n, SCB_ERR_INVALID_ARG, "scb_gemv: ldr too small");
  SCB_REQUIRE(a->epilogue == 0 || a->epilogue == 1, SCB_ERR_UNSUPPORTED, "scb_gemv: epilogue %d", a->epilogue);
  SCB_REQUIRE(a->epilogue == 0 || (a->k % 32 == 0 && a->n % 2 == 0 && !a->bias && !a->residual &&
                                   a->out_dtype == SCB_F16 && !a->argmax_out && a->ldo >= a->n / 2),
              SCB_ERR_INVALID_ARG,
              "scb_gemv: the SiLU-pair epilogue needs k%%32==0, even n, fp16 out, no bias/residual/arg-max");
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
  p.silu_pair = a->epilogue == 1;
  p.eps = a->eps;
  p.pn_gamma = reinterpret_cast<const __half*>(a->post_norm_gamma);
  p.pn_out = reinterpret_cast<__half*>(a->post_norm_out);
  p.pn_ld = a->post_norm_ld;
  p.pn_eps = a->post_norm_eps;
  p.done_ticket = nullptr;
  p.next_w = reinterpret_cast<const char*>(a->prefetch);
  p.next_w_bytes = a->prefetch_bytes;
  const int sms = scb_num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  if (a->k % 32 == 0 && a->prologue == PRO_NONE) {
    // ---- tensor-core streaming path (activations already normalised / gated by the caller)
    const long long n_groups = (a->n + 7) / 8;
    const int ktiles = (int)(a->k / 32);
    static int env_w = -1, env_c = -1;
    if (env_w < 0) {
      const char* e = getenv("SCB_GEMV_W");
      env_w = e ? atoi(e) : 0;
      e = getenv("SCB_GEMV_C");
      env_c = e ? atoi(e) : 0;
    }
    const int GW = (env_w == 8 || env_w == 16 || env_w == 32) ? env_w : kGWDefault;
    static int env_bps = -1;
    if (env_bps < 0) {
      const char* e = getenv("SCB_GEMV_BPS");
      env_bps = e ? atoi(e) : 0;
    }
    // blocks per SM: fewer than the register file allows leaves room for the NEXT kernel's blocks, whose
    // weight loads then start under programmatic dependent launch while this one is still streaming
    const int bps = (env_bps >= 1 && env_bps <= 32 / GW) ? env_bps : 32 / GW;
    const long long wave = (long long)sms * bps;
    // K chunks: fill whole waves (skinny layers), keep >= 2 batches of 8 tiles per warp
    // K chunking (C > 1) only pays when a layer has far fewer row groups than block slots; on the
    // Qwen2-7B shapes the ticket round trip costs more than it gains (profiles/r01_gemv_sweep.txt)
    int C = 1;
    if (n_groups * 4 < wave && n_groups <= kMaxSplitGroups && a->workspace) {
      while (C < kMaxChunks && n_groups * C * 2 <= wave && ktiles / (2 * C * GW) >= kMU) C *= 2;
    }
    if (env_c >= 1 && env_c <= kMaxChunks && n_groups <= kMaxSplitGroups && a->workspace) C = env_c;
    ArgmaxFuse am{};
    float* partial = nullptr;
    unsigned* gtickets = nullptr;
    if (a->workspace) {
      const size_t need = scb_gemv_workspace_bytes(a->rows);
      SCB_REQUIRE(a->workspace_bytes >= need, SCB_ERR_WORKSPACE,
                  "scb_gemv: workspace %zu B < required %zu B", a->workspace_bytes, need);
      char* ws = reinterpret_cast<char*>(a->workspace);
      gtickets = reinterpret_cast<unsigned*>(ws + kWsArgmax);
      partial = reinterpret_cast<float*>(ws + kWsArgmax + kWsTickets);
      p.done_ticket = reinterpret_cast<unsigned*>(ws + 128);
      if (a->argmax_out) {
        am.ticket = reinterpret_cast<unsigned*>(ws);
        am.blk_val = reinterpret_cast<float*>(ws + 256);
        am.blk_idx = reinterpret_cast<int*>(ws + 256 + (size_t)2048 * 16 * sizeof(float));
        am.out_ids = reinterpret_cast<long long*>(a->argmax_out);
        am.counters = a->counters;
        am.n_counters = a->counters ? a->n_counters : 0;
      }
    } else {
      SCB_REQUIRE(!a->argmax_out && !a->post_norm_out, SCB_ERR_WORKSPACE,
                  "scb_gemv: fused arg-max / post-norm need the workspace");
    }
    SCB_REQUIRE(!a->post_norm_out || (a->post_norm_gamma && a->out_dtype == SCB_F16 && a->epilogue == 0 &&
                                      a->n % 8 == 0 && a->ldo % 8 == 0 && a->post_norm_ld % 8 == 0),
                SCB_ERR_INVALID_ARG, "scb_gemv: post-norm needs gamma, fp16 out, n %% 8 == 0, aligned rows");
    const long long units = n_groups * C;
    long long cap = wave < 2048 ? wave : 2048;
    const int grid = (int)(units < cap ? units : cap);
#define SCB_GEMV_GO(GWV, HI)                                                                          \
  SCB_CUDA_OK(scb_launch_pdl(gemv_mma_kernel<GWV, HI>, dim3(grid), dim3(GWV * 32), 0, st, p, (int)a->rows, \
                             C, partial, gtickets, am))
    const bool hi = a->rows > 8;
    if (GW == 4) { if (hi) SCB_GEMV_GO(4, true); else SCB_GEMV_GO(4, false); }
    else if (GW == 8) { if (hi) SCB_GEMV_GO(8, true); else SCB_GEMV_GO(8, false); }
    else if (GW == 16) { if (hi) SCB_GEMV_GO(16, true); else SCB_GEMV_GO(16, false); }
    else { if (hi) SCB_GEMV_GO(32, true); else SCB_GEMV_GO(32, false); }
#undef SCB_GEMV_GO
    SCB_LAUNCH_OK("gemv_mma_kernel");
    return SCB_OK;
  }

  // ---- SIMT path for k not a multiple of 32
  SCB_REQUIRE(a->rows == 1 || a->rows == 2 || a->rows == 4, SCB_ERR_UNSUPPORTED,
              "scb_gemv: the fused-prologue / k %% 32 != 0 path supports rows 1, 2 or 4 only");
  SCB_REQUIRE(!a->argmax_out && a->epilogue == 0 && !a->post_norm_out, SCB_ERR_UNSUPPORTED,
              "scb_gemv: fused arg-max / SiLU-pair / post-norm need k %% 32 == 0 and no prologue");
  const size_t smem = (size_t)a->rows * a->k * sizeof(__half);
  SCB_REQUIRE(smem <= 200 * 1024, SCB_ERR_UNSUPPORTED,
              "scb_gemv: rows*k=%lld does not fit in shared memory", (long long)a->rows * a->k);
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)a->n + kGemvWarps - 1) / kGemvWarps;
  long long cap = (long long)sms * per_sm;
  int grid = (int)(want < cap ? want : cap);
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


size_t scb_decode_attention_fused_workspace_bytes(int32_t n_seqs, int32_t n_q_heads,
                                                  int32_t n_kv_heads, int32_t head_dim,
                                                  int32_t n_splits) {
  if (n_seqs < 1 || n_q_heads < 1 || n_kv_heads < 1 || head_dim < 1 || n_splits < 1) return 0;
  if (n_seqs * n_kv_heads > 256) return 0;  // tickets live in the first 1024 bytes
  return 1024 + sizeof(float) * (size_t)n_seqs * n_q_heads * n_splits * (head_dim + 2);
}

int scb_decode_attention_fused(const scb_decode_attention_fused_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_decode_attention_fused: args is NULL");
  SCB_REQUIRE(a->qkv && a->k_cache && a->v_cache && a->out && a->workspace && a->kv_len_dev,
              SCB_ERR_INVALID_ARG, "scb_decode_attention_fused: NULL buffer");
  SCB_REQUIRE(a->head_dim == 128 || a->head_dim == 64, SCB_ERR_UNSUPPORTED,
              "scb_decode_attention_fused: head_dim=%d (64 or 128)", a->head_dim);
  SCB_REQUIRE(a->n_kv_heads >= 1 && a->n_kv_heads <= 64 && a->n_q_heads % a->n_kv_heads == 0,
              SCB_ERR_INVALID_ARG, "scb_decode_attention_fused: bad head counts");
  const int G = a->n_q_heads / a->n_kv_heads;
  SCB_REQUIRE(G == 1 || G == 2 || G == 4 || G == 7 || G == 8, SCB_ERR_UNSUPPORTED,
              "scb_decode_attention_fused: group size %d unsupported (1,2,4,7,8)", G);
  SCB_REQUIRE(a->n_splits >= 1 && a->n_splits <= 64, SCB_ERR_INVALID_ARG,
              "scb_decode_attention_fused: n_splits out of range (1..64)");
  SCB_REQUIRE(a->n_seqs >= 1 && a->n_seqs * a->n_kv_heads <= 256, SCB_ERR_INVALID_ARG,
              "scb_decode_attention_fused: n_seqs=%d out of range (n_seqs * n_kv_heads <= 256)", a->n_seqs);
  const size_t need = scb_decode_attention_fused_workspace_bytes(a->n_seqs, a->n_q_heads, a->n_kv_heads,
                                                                 a->head_dim, a->n_splits);
  SCB_REQUIRE(a->workspace_bytes >= need, SCB_ERR_WORKSPACE,
              "scb_decode_attention_fused: workspace %zu B < %zu B", a->workspace_bytes, need);
  char* ws = reinterpret_cast<char*>(a->workspace);
  unsigned* tickets = reinterpret_cast<unsigned*>(ws);   // zeroed by scb_decode_attention_fused_init
  float* part_m = reinterpret_cast<float*>(ws + 1024);
  float* part_l = part_m + (size_t)a->n_seqs * a->n_q_heads * a->n_splits;
  float* part_o = part_l + (size_t)a->n_seqs * a->n_q_heads * a->n_splits;
  // score-buffer sizing: the caller may bound the live context (ctx_cap <= max_pos) so that short
  // contexts do not pay shared memory (and occupancy) for the cache capacity
  const long long cap = (a->ctx_cap > 0 && a->ctx_cap < a->max_pos) ? a->ctx_cap : a->max_pos;
  const int per_need = (int)((cap + a->n_splits - 1) / a->n_splits);
  const int per_min = per_need > a->n_splits ? per_need : a->n_splits;   // the combine reuses the score buffer
  const int per_max = ((per_min + 11) / 16) * 16 + 4;   // row stride = 4 mod 16 floats: conflict-free mma stores
  const size_t smem = sizeof(float) * ((size_t)4 * a->head_dim + (size_t)G * per_max);
  SCB_REQUIRE(smem <= 160 * 1024, SCB_ERR_UNSUPPORTED,
              "scb_decode_attention_fused: max_pos/n_splits too large for shared memory");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(a->n_kv_heads, a->n_splits, a->n_seqs);
  const float sl2 = a->scale * 1.4426950408889634f;
#define SCB_DAF(HD, GG)                                                                           \
  do {                                                                                            \
    auto kern = decode_attn_fused_kernel<HD, GG>;                                                 \
    if (smem > 8 * 1024) /* static o_s takes up to 32 KB of the default 48 KB */                  \
      SCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                       (int)smem));                                               \
    SCB_CUDA_OK(scb_launch_pdl(kern, grid, dim3(256), smem, st,                                   \
                               reinterpret_cast<const __half*>(a->qkv), a->inv_freq, a->theta,    \
                               reinterpret_cast<__half*>(a->k_cache),                             \
                               reinterpret_cast<__half*>(a->v_cache), (long long)a->max_pos,      \
                               a->kv_len_dev, sl2, (int)a->n_q_heads, (int)a->n_kv_heads, per_max,\
                               part_m, part_l, part_o, tickets,                                   \
                               reinterpret_cast<__half*>(a->out), (long long)a->ld_qkv,           \
                               (long long)a->cache_seq_stride,                                    \
                               reinterpret_cast<const char*>(a->prefetch),                        \
                               (long long)a->prefetch_bytes));                                    \
  } while (0)
  if (a->head_dim == 128) {
    switch (G) {
      case 1: SCB_DAF(128, 1); break;
      case 2: SCB_DAF(128, 2); break;
      case 4: SCB_DAF(128, 4); break;
      case 7: SCB_DAF(128, 7); break;
      default: SCB_DAF(128, 8); break;
    }
  } else {
    switch (G) {
      case 1: SCB_DAF(64, 1); break;
      case 2: SCB_DAF(64, 2); break;
      case 4: SCB_DAF(64, 4); break;
      case 7: SCB_DAF(64, 7); break;
      default: SCB_DAF(64, 8); break;
    }
  }
#undef SCB_DAF
  SCB_LAUNCH_OK("decode_attn_fused_kernel");
  return SCB_OK;
}

}  // extern "C"
