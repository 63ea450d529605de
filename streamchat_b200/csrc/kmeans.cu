// kmeans.cu — one Lloyd iteration (assign + weighted update + centre shift), HBM-bound.
//
// Replaces the ATen broadcast formulation of weighted_kmeans_feature (reference
// utiles.py:297-317: an [N,K,D] temporary per iteration plus a Python loop over clusters)
// and torch_kmeans' _assign / group_by_label_mean (torch_kmeans/clustering/kmeans.py:572-597,
// torch_kmeans/utils/utils.py:33-66: (BS*num_init*N*K) x D expanded rows).
//
// Data layout: points [N, D] row-major fp16/fp32 (read exactly twice per iteration — once by
// assign, once by update), centroids [B, K, D] fp32 (L2-resident: K*D*4 bytes).
// Algorithmic bytes per iteration: 2 * N*D*sizeof(T)  (+ 3 * B*K*D*4 for the centroids).
//
//   assign_partial : grid (column tiles, point splits, B).  The block stages its centroid
//                    tile in shared memory; a thread owns 16-byte vector columns, keeps
//                    4 points x K accumulators in registers and does the (x-c)^2 sums with
//                    packed fp32x2 FADD2/FFMA2.  Partials are reduced in a fixed order
//                    (deterministic labels; no float atomics).
//   assign_labels  : warp per (b, point): sum slices in order, sqrt, first-min argmin.
//   assign_sort    : per (b): weighted counts, counting sort of the points by label.
//   update_partial : grid (column chunks, B).  Thread owns a vector column, walks the sorted
//                    member list cluster by cluster (single accumulator set), writes the new
//                    centroid and a partial of |c_old - c_new|^2.
//   update_final   : shift[b,k] = sqrt(sum of partials).
#include "scb_common.cuh"
#include <math.h>

namespace {

constexpr int kThreads = 256;
constexpr int kPG = 4;        // points per group (accumulator rows per thread)
constexpr int kKT = 8;        // clusters per register tile
constexpr int kTileBytes = 96 * 1024;  // centroid tile in shared memory (2 blocks / SM)
constexpr int kMaxClusters = 64;

typedef unsigned long long u64;

SCB_DEVICE u64 pack2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
SCB_DEVICE float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
SCB_DEVICE u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
SCB_DEVICE u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
SCB_DEVICE void sqacc2(u64& acc, u64 d) {
  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d));
}

template <typename T>
struct PVec;  // one 16-byte vector of points -> fp32 pairs
template <>
struct PVec<float> {
  static constexpr int kElems = 4, kPairs = 2;
  SCB_DEVICE static void unpack(const uint4& v, u64* p) {
    p[0] = ((u64)v.y << 32) | v.x;
    p[1] = ((u64)v.w << 32) | v.z;
  }
  SCB_DEVICE static void to_float(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  SCB_DEVICE static float load1(const float* p) { return *p; }
};
template <>
struct PVec<__half> {
  static constexpr int kElems = 8, kPairs = 4;
  SCB_DEVICE static void unpack(const uint4& v, u64* p) {
    float2 a = h2_to_f2(v.x), b = h2_to_f2(v.y), c = h2_to_f2(v.z), d = h2_to_f2(v.w);
    p[0] = pack2(a.x, a.y); p[1] = pack2(b.x, b.y);
    p[2] = pack2(c.x, c.y); p[3] = pack2(d.x, d.y);
  }
  SCB_DEVICE static void to_float(const uint4& v, float* f) {
    float2 a = h2_to_f2(v.x), b = h2_to_f2(v.y), c = h2_to_f2(v.z), d = h2_to_f2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  SCB_DEVICE static float load1(const __half* p) { return __half2float(*p); }
};

// ------------------------------------------------------------------------------------------
// assign_partial
//   partial[b][tile][n][k] = sum over the tile's columns of (x[n][d] - c[b][k][d] + eps)^2
// The block stages its [kt, tile_cols] centroid tile in shared memory once and re-uses it for
// every point it owns, so L2 sees the centroids only (point splits) times.
// kVec: D % E == 0 and 16-byte aligned rows -> 128-bit loads; else one column per thread.
// ------------------------------------------------------------------------------------------
template <typename T, bool kVec, bool kEps>
__global__ void __launch_bounds__(kThreads, 2)
kmeans_assign_partial(const T* __restrict__ x, const float* __restrict__ cent, int n_points,
                      int n_clusters, long long dim, int tile_cols, int points_per_split, float eps,
                      float* __restrict__ partial) {
  constexpr int E = kVec ? PVec<T>::kElems : 1;
  constexpr int NP = kVec ? PVec<T>::kPairs : 1;
  extern __shared__ __align__(16) float cs[];  // [kKT][tile_cols]
  __shared__ float red[kThreads / 32][kPG][kKT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.z;
  const int n_tiles = gridDim.x;
  const long long col0 = (long long)tile * tile_cols;
  const int ncols = (int)min((long long)tile_cols, dim - col0);
  const int n_begin = blockIdx.y * points_per_split;
  const int n_end = min(n_points, n_begin + points_per_split);
  const float* cb = cent + (long long)b * n_clusters * dim + col0;
  const u64 eps2 = pack2(eps, eps);

  for (int k0 = 0; k0 < n_clusters; k0 += kKT) {
    const int kt = min(kKT, n_clusters - k0);
    // ---- stage the centroid tile
    if constexpr (kVec) {
      const int n4 = ncols >> 2;
      for (int i = tid; i < kt * n4; i += kThreads) {
        const int k = i / n4, c = (i - k * n4) << 2;
        *reinterpret_cast<float4*>(cs + k * tile_cols + c) =
            *reinterpret_cast<const float4*>(cb + (long long)(k0 + k) * dim + c);
      }
    } else {
      for (int i = tid; i < kt * ncols; i += kThreads) {
        const int k = i / ncols, c = i - k * ncols;
        cs[k * tile_cols + c] = cb[(long long)(k0 + k) * dim + c];
      }
    }
    __syncthreads();

    for (int p0 = n_begin; p0 < n_end; p0 += kPG) {
      u64 acc[kPG][kKT];  // packed (even, odd) partial sums
#pragma unroll
      for (int p = 0; p < kPG; ++p)
#pragma unroll
        for (int k = 0; k < kKT; ++k) acc[p][k] = 0ull;

      for (int c = tid * E; c < ncols; c += kThreads * E) {
        u64 xv[kPG][NP];
        if constexpr (kVec) {
          uint4 raw[kPG];
#pragma unroll
          for (int p = 0; p < kPG; ++p) {
            int n = p0 + p;
            if (n >= n_end) n = n_end - 1;  // clamped duplicate, discarded below
            raw[p] = ld_stream_u4(x + (long long)n * dim + col0 + c);
          }
#pragma unroll
          for (int p = 0; p < kPG; ++p) PVec<T>::unpack(raw[p], xv[p]);
        } else {
#pragma unroll
          for (int p = 0; p < kPG; ++p) {
            int n = p0 + p;
            if (n >= n_end) n = n_end - 1;
            xv[p][0] = pack2(PVec<T>::load1(x + (long long)n * dim + col0 + c), 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < kKT; ++k) {
          if (k < kt) {
            u64 cv[NP];
            const float* cp = cs + k * tile_cols + c;
            if constexpr (kVec) {
#pragma unroll
              for (int j = 0; j < NP; j += 2) {
                ulonglong2 t = *reinterpret_cast<const ulonglong2*>(cp + j * 2);
                cv[j] = t.x;
                cv[j + 1] = t.y;
              }
            } else {
              // unused upper half: (0 - eps) + eps == 0 exactly, so it never contributes
              cv[0] = pack2(*cp, kEps ? eps : 0.f);
            }
#pragma unroll
            for (int p = 0; p < kPG; ++p)
#pragma unroll
              for (int j = 0; j < NP; ++j) {
                u64 d = sub2(xv[p][j], cv[j]);
                if constexpr (kEps) d = add2(d, eps2);
                sqacc2(acc[p][k], d);
              }
          }
        }
      }

      // block reduction in a fixed order: lanes (xor tree), then warps 0..7
#pragma unroll
      for (int p = 0; p < kPG; ++p)
#pragma unroll
        for (int k = 0; k < kKT; ++k) {
          if (k < kt) {
            float2 a = unpack2(acc[p][k]);
            float v = warp_sum(a.x + a.y);
            if (lane == 0) red[warp][p][k] = v;
          }
        }
      __syncthreads();
      if (tid < kPG * kKT) {
        const int p = tid / kKT, k = tid % kKT;
        const int n = p0 + p;
        if (n < n_end && k < kt) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kThreads / 32; ++w) v += red[w][p][k];
          partial[(((long long)b * n_tiles + tile) * n_points + n) * n_clusters + k0 + k] = v;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// assign_partial_async — the vector path at HBM speed.  Same tiling as above, but the point
// vectors travel global -> shared with cp.async (16 B per thread, L1 bypass) through a private
// ring of kST stages per thread: three stages (12 loads of 16 B per thread, 96 KB per SM) are in
// flight while the fourth is consumed, without spending registers on them; a thread only ever
// reads back its own slots, so the ring needs cp.async.wait_group and no block barrier.  A warp
// owns whole point groups and reduces over its lanes once per group:
//   partial[b][tile][n][k]                   (summed in a fixed order by assign_final)
// ------------------------------------------------------------------------------------------
constexpr int kST = 4;

SCB_DEVICE void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
SCB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SCB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, bool kEps, int KT>
__global__ void __launch_bounds__(kThreads, 2)
kmeans_assign_partial_async(const T* __restrict__ x, const float* __restrict__ cent, int n_points,
                            int n_clusters, long long dim, int tile_cols, int points_per_split,
                            float eps, float* __restrict__ partial) {
  constexpr int E = PVec<T>::kElems;
  constexpr int NP = PVec<T>::kPairs;
  constexpr int NW = kThreads / 32;
  extern __shared__ __align__(16) float cs[];          // [KT][tile_cols] then the x ring
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.z;
  const int n_tiles = gridDim.x;
  uint4* xs = reinterpret_cast<uint4*>(cs + (size_t)KT * tile_cols);   // [kST][kPG][kThreads]
  const long long col0 = (long long)tile * tile_cols;
  const int ncols = (int)min((long long)tile_cols, dim - col0);
  const int n_begin = blockIdx.y * points_per_split;
  const int n_end = min(n_points, n_begin + points_per_split);
  const float* cb = cent + (long long)b * n_clusters * dim + col0;
  const u64 eps2 = pack2(eps, eps);
  // a WARP owns whole point groups (4 points) and walks the tile's columns in `passes` steps of
  // 32 lanes x 16 B, so the lane reduction happens once per point group, not once per load
  const int passes = (ncols + 32 * E - 1) / (32 * E);
  const int n_pg = (n_end - n_begin + kPG - 1) / kPG;
  const int n_pg_w = n_pg > warp ? (n_pg - warp + NW - 1) / NW : 0;
  const int n_it = n_pg_w * passes;

  auto issue = [&](int it) {
    if (it < n_it) {
      const int pgi = it / passes, pass = it - pgi * passes;
      const int c = (pass * 32 + lane) * E;
      if (c < ncols) {
        const int p0 = n_begin + (warp + pgi * NW) * kPG;
#pragma unroll
        for (int p = 0; p < kPG; ++p) {
          int n = p0 + p;
          if (n >= n_end) n = n_end - 1;   // clamped duplicate, discarded below
          cp_async16(xs + ((it % kST) * kPG + p) * kThreads + tid, x + (long long)n * dim + col0 + c);
        }
      }
    }
    cp_async_commit();
  };

  for (int k0 = 0; k0 < n_clusters; k0 += KT) {
    const int kt = min(KT, n_clusters - k0);
#pragma unroll
    for (int s = 0; s < kST - 1; ++s) issue(s);        // the points do not wait for the centroids
    {
      const int n4 = ncols >> 2;
      for (int i = tid; i < kt * n4; i += kThreads) {
        const int k = i / n4, c = (i - k * n4) << 2;
        *reinterpret_cast<float4*>(cs + k * tile_cols + c) =
            *reinterpret_cast<const float4*>(cb + (long long)(k0 + k) * dim + c);
      }
    }
    __syncthreads();

    u64 acc[kPG][KT];
#pragma unroll
    for (int p = 0; p < kPG; ++p)
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[p][k] = 0ull;

    for (int it = 0; it < n_it; ++it) {
      issue(it + kST - 1);
      cp_async_wait<kST - 1>();                          // stage `it` has landed (own slots only)
      const int pgi = it / passes, pass = it - pgi * passes;
      const int c = (pass * 32 + lane) * E;
      if (c < ncols) {
        u64 xv[kPG][NP];
#pragma unroll
        for (int p = 0; p < kPG; ++p)
          PVec<T>::unpack(xs[((it % kST) * kPG + p) * kThreads + tid], xv[p]);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < kt) {
            u64 cv[NP];
            const float* cp = cs + k * tile_cols + c;
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
              ulonglong2 t = *reinterpret_cast<const ulonglong2*>(cp + j * 2);
              cv[j] = t.x;
              cv[j + 1] = t.y;
            }
#pragma unroll
            for (int p = 0; p < kPG; ++p)
#pragma unroll
              for (int j = 0; j < NP; ++j) {
                u64 d = sub2(xv[p][j], cv[j]);
                if constexpr (kEps) d = add2(d, eps2);
                sqacc2(acc[p][k], d);
              }
          }
        }
      }
      if (pass == passes - 1) {
        // point group finished for this warp: xor-tree over the lanes, lane 0 writes
        const int p0 = n_begin + (warp + pgi * NW) * kPG;
#pragma unroll
        for (int p = 0; p < kPG; ++p)
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            float2 a = unpack2(acc[p][k]);
            const float v = warp_sum(a.x + a.y);
            acc[p][k] = 0ull;
            if (lane == 0 && k < kt && p0 + p < n_end)
              partial[(((long long)b * n_tiles + tile) * n_points + p0 + p) * n_clusters + k0 + k] = v;
          }
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // centroid tile reuse by the next cluster tile
  }
}

// ------------------------------------------------------------------------------------------
// assign_dot_mma — the assign pass for LARGE fp16 problems that only need labels (the memory-tree merge:
// 400 frames x 516 096 dims).  The packed-fp32 (x-c)^2 kernel above is instruction-bound (~440 warp
// instructions per 2 KB of points: 2.6 TB/s); argmin_k |x-c_k|^2 == argmin_k (|c_k|^2 - 2 x.c_k), and the
// dot products are a k-reduction the tensor pipe does as a side job: mma.sync.m16n8k16 with 16 POINTS x 16
// dims straight from HBM as the A operand (fp16 as stored — no conversion, no staging) and the centroids
// as the B operand.  Centroids are fp32 means, so each one enters as TWO fp16 terms (hi = fp16(c),
// lo = fp16(c - hi)): the fp16 products are exact in the fp32 accumulator and the pair carries 22 bits
// of the centroid (absolute floor 2^-25 from fp16 subnormals), i.e. the dot products are as accurate as the fp32 FMA chain.  ~24 warp instructions
// per 2 KB.  Not used when the caller wants distances (cancellation near a point's own centroid) or the
// eps-shifted torch distance: those keep the exact kernel.
//   partial[b][n][k][tile*4 + quarter] = sum over the quarter's columns of x[n][d] * c[b][k][d]
//                                        (slice index contiguous: the label kernel reads whole lines)
//   cnorm_partial[b][tile][k]          = sum over the tile's columns of c[b][k][d]^2
// ------------------------------------------------------------------------------------------
constexpr int kDotTileCols = 2048;
constexpr int kDotThreads = 256;   // (512-column tiles / 128-thread blocks were measured slower: 119 vs 113 us)
constexpr int kDotRowBytes = kDotTileCols * 2 + 64;      // +64 B: clusters g and g+1 land on different banks
constexpr int kDotSmem = 2 * 8 * kDotRowBytes;           // hi and lo tiles, 8 cluster rows each

SCB_DEVICE void mma16816_km(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                            uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kDotThreads, 2)
kmeans_assign_dot_mma(const __half* __restrict__ x, const float* __restrict__ cent, int n_points,
                      int n_clusters, long long dim, float* __restrict__ partial,
                      float* __restrict__ cnorm_partial) {
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* c_hi = dsm;
  uint8_t* c_lo = dsm + 8 * kDotRowBytes;
  __shared__ float cn_red[kDotThreads / 32][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.y, n_tiles = gridDim.x;
  const long long col0 = (long long)tile * kDotTileCols;
  const int ncols = (int)min((long long)kDotTileCols, dim - col0);
  const float* cb = cent + (long long)b * n_clusters * dim + col0;

  // ---- stage the centroid tile as fp16 (hi, lo) pairs; rows >= n_clusters and columns >= ncols are zero
  {
    float cn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cn[k] = 0.f;
    for (int i = tid; i < 8 * (kDotTileCols / 4); i += kDotThreads) {
      const int k = i / (kDotTileCols / 4), c = (i - k * (kDotTileCols / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < n_clusters && c < ncols) v = *reinterpret_cast<const float4*>(cb + (long long)k * dim + c);   // ncols % 8 == 0
      const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
      const __half l0 = __float2half_rn(v.x - __half2float(h0)), l1 = __float2half_rn(v.y - __half2float(h1));
      const __half l2 = __float2half_rn(v.z - __half2float(h2)), l3 = __float2half_rn(v.w - __half2float(h3));
      uint2 hv, lv;
      hv.x = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      hv.y = (uint32_t)__half_as_ushort(h2) | ((uint32_t)__half_as_ushort(h3) << 16);
      lv.x = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      lv.y = (uint32_t)__half_as_ushort(l2) | ((uint32_t)__half_as_ushort(l3) << 16);
      *reinterpret_cast<uint2*>(c_hi + k * kDotRowBytes + c * 2) = hv;
      *reinterpret_cast<uint2*>(c_lo + k * kDotRowBytes + c * 2) = lv;
      const float sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) cn[kk] += (kk == k) ? sq : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = warp_sum(cn[k]);
      if (lane == 0) cn_red[warp][k] = v;
    }
  }
  __syncthreads();
  if (tid < n_clusters) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDotThreads / 32; ++w) v += cn_red[w][tid];
    cnorm_partial[((long long)b * n_tiles + tile) * n_clusters + tid] = v;
  }

  // ---- work items = (group of 16 points, quarter of the tile's columns), handed to the warps round-robin
  const int g = lane >> 2, t = lane & 3;
  const int n_groups = (n_points + 15) >> 4;
  const uint32_t hi_s = smem_u32(c_hi) + (uint32_t)(g * kDotRowBytes + t * 16);
  const uint32_t lo_s = smem_u32(c_lo) + (uint32_t)(g * kDotRowBytes + t * 16);
  constexpr int QCOLS = kDotTileCols / 4, QSTEPS = QCOLS / 32;
  for (int item = warp; item < n_groups * 4; item += kDotThreads / 32) {
    const int grp = item >> 2, q = item & 3;
    const int r0 = min(grp * 16 + g, n_points - 1), r1 = min(grp * 16 + g + 8, n_points - 1);   // clamped rows are discarded
    const __half* p0 = x + (long long)r0 * dim + col0 + q * QCOLS + t * 8;
    const __half* p1 = x + (long long)r1 * dim + col0 + q * QCOLS + t * 8;
    const int lim = ncols - q * QCOLS - t * 8;     // columns left from this thread's first element
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s0 = 0; s0 < QSTEPS; s0 += 4) {
      uint4 a0[4], a1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {          // 8 x 16 B in flight per thread
        const bool ok = (s0 + u) * 32 < lim;
        a0[u] = ok ? ld_stream_u4(p0 + (s0 + u) * 32) : make_uint4(0, 0, 0, 0);
        a1[u] = ok ? ld_stream_u4(p1 + (s0 + u) * 32) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t off = (uint32_t)((q * QCOLS + (s0 + u) * 32) * 2);
        uint4 bh, bl;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(bh.x), "=r"(bh.y), "=r"(bh.z), "=r"(bh.w) : "r"(hi_s + off));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(bl.x), "=r"(bl.y), "=r"(bl.z), "=r"(bl.w) : "r"(lo_s + off));
        mma16816_km(d0, a0[u].x, a1[u].x, a0[u].y, a1[u].y, bh.x, bh.y);
        mma16816_km(d1, a0[u].z, a1[u].z, a0[u].w, a1[u].w, bh.z, bh.w);
        mma16816_km(d0, a0[u].x, a1[u].x, a0[u].y, a1[u].y, bl.x, bl.y);
        mma16816_km(d1, a0[u].z, a1[u].z, a0[u].w, a1[u].w, bl.z, bl.w);
      }
    }
    // D[m][n]: thread (g,t) holds points g / g+8, clusters 2t / 2t+1
    const long long n_slices = (long long)n_tiles * 4;
    float* pp = partial + (long long)b * n_points * n_clusters * n_slices + (tile * 4 + q);
    const int n0 = grp * 16 + g, n1 = n0 + 8, k0 = 2 * t;
    if (n0 < n_points) {
      if (k0 < n_clusters) pp[((long long)n0 * n_clusters + k0) * n_slices] = d0[0] + d1[0];
      if (k0 + 1 < n_clusters) pp[((long long)n0 * n_clusters + k0 + 1) * n_slices] = d0[1] + d1[1];
    }
    if (n1 < n_points) {
      if (k0 < n_clusters) pp[((long long)n1 * n_clusters + k0) * n_slices] = d0[2] + d1[2];
      if (k0 + 1 < n_clusters) pp[((long long)n1 * n_clusters + k0 + 1) * n_slices] = d0[3] + d1[3];
    }
  }
}

// labels from the dot products: argmin_k (|c_k|^2 - 2 x.c_k), first minimum wins
__global__ void __launch_bounds__(kThreads)
kmeans_assign_labels_dot(const float* __restrict__ partial, const float* __restrict__ cnorm_partial,
                         int n_points, int n_clusters, int n_slices, int n_tiles, int* __restrict__ labels) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * (kThreads / 32) + warp;
  if (n >= n_points) return;
  const float* pp = partial + ((long long)b * n_points + n) * n_clusters * (long long)n_slices;
  float best = INFINITY;
  int bi = 0;
  for (int k = 0; k < n_clusters; ++k) {
    const float* pk = pp + (long long)k * n_slices;
    float s = 0.f, cn = 0.f;
    // 32 independent coalesced loads in flight per lane (one round trip per 1024 slices), fixed order
    for (int s0 = 0; s0 < n_slices; s0 += 1024) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int sl = s0 + j * 32 + lane;
        v[j] = sl < n_slices ? pk[sl] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) s += v[j];
    }
    for (int tl = lane; tl < n_tiles; tl += 32) cn += cnorm_partial[((long long)b * n_tiles + tl) * n_clusters + k];
    s = warp_sum(s);
    cn = warp_sum(cn);
    const float score = cn - 2.f * s;
    if (k == 0 || score < best) {
      best = score;
      bi = k;
    }
  }
  if (lane == 0) labels[(long long)b * n_points + n] = bi;
}

// ------------------------------------------------------------------------------------------
// assign_final: two small kernels.
//   assign_labels : one WARP per (batch entry, point), 8 points per block: the K x n_slices partial sums of
//                   a point are summed in slice order with every load of a cluster in flight at once.
//                   (The first version ran all points of a batch entry through ONE block, a serial chain
//                   of ~250 dependent L2 round trips: 320 us of the 617 us Lloyd iteration at 400 x 516 096.)
//   assign_sort   : one block per batch entry: stable counting sort of the points by label + weight sums
//                   from a shared-memory copy of the labels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
kmeans_assign_labels(const float* __restrict__ partial, int n_points, int n_clusters, int n_slices,
                     int* __restrict__ labels, float* __restrict__ out_dist) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * (kThreads / 32) + warp;
  if (n >= n_points) return;
  const float* pp = partial + ((long long)b * n_slices * n_points + n) * n_clusters;
  const long long slice_stride = (long long)n_points * n_clusters;
  float best = INFINITY;
  int bi = 0;
  for (int k = 0; k < n_clusters; ++k) {
    // lanes stride over the slices; fixed order per lane, then the xor tree -> deterministic sums
    float s = 0.f;
    int sl = lane;
    for (; sl + 96 < n_slices; sl += 128) {
      const float a0 = pp[(long long)sl * slice_stride + k], a1 = pp[(long long)(sl + 32) * slice_stride + k];
      const float a2 = pp[(long long)(sl + 64) * slice_stride + k], a3 = pp[(long long)(sl + 96) * slice_stride + k];
      s += a0; s += a1; s += a2; s += a3;
    }
    for (; sl < n_slices; sl += 32) s += pp[(long long)sl * slice_stride + k];
    s = warp_sum(s);
    const float d = sqrtf(s);
    if (k == 0 || d < best) {  // first minimum wins (torch.argmin; NaN-free inputs assumed)
      best = d;
      bi = k;
    }
  }
  if (lane == 0) {
    labels[(long long)b * n_points + n] = bi;
    if (out_dist) out_dist[(long long)b * n_points + n] = best;
  }
}

__global__ void __launch_bounds__(kThreads)
kmeans_assign_sort(const int* __restrict__ labels, const float* __restrict__ weights, int n_points,
                   int n_clusters, float* __restrict__ wsum, int* __restrict__ order,
                   int* __restrict__ starts) {
  const int b = blockIdx.x, tid = threadIdx.x;
  extern __shared__ int lab_s[];            // [n_points]
  __shared__ int cnt[kMaxClusters + 1];
  for (int i = tid; i < n_points; i += kThreads) lab_s[i] = labels[(long long)b * n_points + i];
  if (tid <= n_clusters) cnt[tid] = 0;
  __syncthreads();
  if (weights == nullptr) {
    // Unweighted (every caller on the hot path): warp w owns the contiguous point range [w*per, (w+1)*per);
    // ballots count / place the members of each cluster in ascending point order (stable), counts are exact
    // integers so wsum is bit-identical to the serial sum.
    __shared__ int wcnt[kThreads / 32][kMaxClusters];
    const int lane = tid & 31, warp = tid >> 5;
    const int per = (((n_points + kThreads / 32 - 1) / (kThreads / 32)) + 31) & ~31;
    const int n0 = warp * per, n1 = min(n_points, n0 + per);
    for (int k = lane; k < n_clusters; k += 32) wcnt[warp][k] = 0;
    __syncwarp();
    for (int c0 = n0; c0 < n1; c0 += 32) {
      const int lab = (c0 + lane < n1) ? lab_s[c0 + lane] : -1;
      for (int k = 0; k < n_clusters; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, lab == k);
        if (lane == 0 && m) wcnt[warp][k] += __popc(m);
      }
    }
    __syncthreads();
    if (tid < n_clusters) {
      int c = 0;
      for (int w = 0; w < kThreads / 32; ++w) {
        const int v = wcnt[w][tid];
        wcnt[w][tid] = c;              // exclusive prefix over the warps: this warp's first slot inside cluster tid
        c += v;
      }
      cnt[tid + 1] = c;
      wsum[(long long)b * n_clusters + tid] = (float)c;
    }
    __syncthreads();
    if (tid == 0) {
      for (int k = 0; k < n_clusters; ++k) cnt[k + 1] += cnt[k];
      for (int k = 0; k <= n_clusters; ++k) starts[(long long)b * (n_clusters + 1) + k] = cnt[k];
    }
    __syncthreads();
    for (int c0 = n0; c0 < n1; c0 += 32) {
      const int lab = (c0 + lane < n1) ? lab_s[c0 + lane] : -1;
      for (int k = 0; k < n_clusters; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, lab == k);
        if (lab == k) order[(long long)b * n_points + cnt[k] + wcnt[warp][k] + __popc(m & ((1u << lane) - 1u))] = c0 + lane;
        __syncwarp();
        if (lane == 0 && m) wcnt[warp][k] += __popc(m);
        __syncwarp();
      }
    }
    return;
  }
  // weighted: serial per cluster owner, weight sums in ascending point order (deterministic)
  if (tid < n_clusters) {
    int c = 0;
    float ws = 0.f;
    for (int n = 0; n < n_points; ++n)
      if (lab_s[n] == tid) {
        ++c;
        ws += weights[n];
      }
    cnt[tid + 1] = c;
    wsum[(long long)b * n_clusters + tid] = ws;
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < n_clusters; ++k) cnt[k + 1] += cnt[k];
    for (int k = 0; k <= n_clusters; ++k) starts[(long long)b * (n_clusters + 1) + k] = cnt[k];
  }
  __syncthreads();
  if (tid < n_clusters) {
    int pos = cnt[tid];
    for (int n = 0; n < n_points; ++n)
      if (lab_s[n] == tid) order[(long long)b * n_points + pos++] = n;
  }
}

// ------------------------------------------------------------------------------------------
// update_partial
// ------------------------------------------------------------------------------------------
template <typename T, bool kVec, int TH>
__global__ void __launch_bounds__(TH)
kmeans_update_partial(const T* __restrict__ x, const float* __restrict__ weights,
                      const float* __restrict__ cent, const int* __restrict__ order,
                      const int* __restrict__ starts, const float* __restrict__ wsum, int n_points,
                      int n_clusters, long long dim, float* __restrict__ new_cent,
                      float* __restrict__ shift_partial) {
  constexpr int E = kVec ? PVec<T>::kElems : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int n_chunks = gridDim.x;
  extern __shared__ int sm_order[];  // [n_points] member list of this batch entry
  __shared__ int sm_starts[kMaxClusters + 1];
  __shared__ float red[TH / 32];
  for (int i = tid; i < n_points; i += TH) sm_order[i] = order[(long long)b * n_points + i];
  if (tid <= n_clusters) sm_starts[tid] = starts[(long long)b * (n_clusters + 1) + tid];
  __syncthreads();

  const long long col = ((long long)blockIdx.x * TH + tid) * E;
  const bool active = col < dim;
  for (int k = 0; k < n_clusters; ++k) {
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const int beg = sm_starts[k], end = sm_starts[k + 1];
    float sq = 0.f;
    if (active) {
      int i = beg;
      if constexpr (kVec) {
        // the grid has only D/2048 blocks (~14 warps per SM at 516 096 dims), so every thread keeps
        // 16 independent 16-byte loads in flight (~110 KB per SM) to cover the HBM latency
        for (; i + 16 <= end; i += 16) {
          uint4 r[16];
          float w[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int n = sm_order[i + u];
            r[u] = ld_stream_u4(x + (long long)n * dim + col);
            w[u] = weights ? weights[n] : 1.f;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float f[E];
            PVec<T>::to_float(r[u], f);
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(w[u], f[e], acc[e]);
          }
        }
        for (; i + 4 <= end; i += 4) {  // four independent 16-byte loads in flight
          uint4 r[4];
          float w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = sm_order[i + u];
            r[u] = ld_stream_u4(x + (long long)n * dim + col);
            w[u] = weights ? weights[n] : 1.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[E];
            PVec<T>::to_float(r[u], f);
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(w[u], f[e], acc[e]);
          }
        }
      }
      for (; i < end; ++i) {
        const int n = sm_order[i];
        const float w = weights ? weights[n] : 1.f;
        if constexpr (kVec) {
          uint4 r = ld_stream_u4(x + (long long)n * dim + col);
          float f[E];
          PVec<T>::to_float(r, f);
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = fmaf(w, f[e], acc[e]);
        } else {
          acc[0] = fmaf(w, PVec<T>::load1(x + (long long)n * dim + col), acc[0]);
        }
      }
      const float ws = wsum[(long long)b * n_clusters + k];
      const long long off = ((long long)b * n_clusters + k) * dim + col;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float nc = ws > 0.f ? acc[e] / ws : 0.f;
        const float d = cent[off + e] - nc;
        sq = fmaf(d, d, sq);
        new_cent[off + e] = nc;
      }
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < TH / 32; ++w) v += red[w];
      shift_partial[((long long)b * n_clusters + k) * n_chunks + blockIdx.x] = v;
    }
    __syncthreads();
  }
}

__global__ void kmeans_update_final(const float* __restrict__ shift_partial, int n_chunks,
                                    float* __restrict__ shift) {
  // one warp per (b, k); fixed-order strided sum then xor tree
  const int bk = blockIdx.x, lane = threadIdx.x;
  float v = 0.f;
  for (int c = lane; c < n_chunks; c += 32) v += shift_partial[(long long)bk * n_chunks + c];
  v = warp_sum(v);
  if (lane == 0) shift[bk] = sqrtf(v);
}

struct Plan {
  bool vec;
  int tile_cols, n_tiles;         // assign: centroid tile (columns) and tile count
  int points_per_split, n_splits; // assign: points per block
  size_t assign_smem;
  int n_slices;                   // partial-sum slices summed by assign_final
  int kt_async = 8;
  int n_chunks;                   // update: column chunks
  int dot_tiles;                  // assign_dot_mma: column tiles of 2048 (0 = path not applicable)
  size_t partial_off, order_off, starts_off, shiftp_off, cnorm_off, total;
};

Plan make_plan(int n_points, int n_clusters, int n_batch, long long dim, int esize, bool aligned) {
  Plan p;
  const int E = 16 / esize;
  p.vec = aligned && (dim % E == 0);
  const int ce = p.vec ? E : 1;
  const long long cols_per_pass = (long long)kThreads * ce;
  const long long passes = (dim + cols_per_pass - 1) / cols_per_pass;
  const int kt = n_clusters < kKT ? n_clusters : kKT;
  const long long tile_budget = p.vec ? (40 * 1024) : kTileBytes;   // vec: tile + 64 KB ring, 2 blocks / SM
  long long tile_passes = tile_budget / ((long long)kt * 4 * cols_per_pass);
  if (tile_passes < 1) tile_passes = 1;
  if (tile_passes > passes) tile_passes = passes;
  p.tile_cols = (int)(tile_passes * cols_per_pass);
  p.n_tiles = (int)((dim + p.tile_cols - 1) / p.tile_cols);
  p.assign_smem = (size_t)kt * p.tile_cols * sizeof(float);
  p.n_slices = p.n_tiles;
  if (p.vec) {   // cp.async kernel: exact cluster tile, per-thread x ring, per-warp partial slices
    p.kt_async = n_clusters <= 4 ? 4 : (n_clusters == 5 ? 5 : 8);
    p.assign_smem = (size_t)p.kt_async * p.tile_cols * sizeof(float) + (size_t)kST * kPG * kThreads * 16;
  }
  // split the points so that ~4 blocks per SM exist (each split re-reads the centroid tile from L2)
  const int n_groups = (n_points + kPG - 1) / kPG;
  long long splits = (148LL * 4 + (long long)p.n_tiles * n_batch - 1) / ((long long)p.n_tiles * n_batch);
  if (splits < 1) splits = 1;
  if (splits > n_groups) splits = n_groups;
  const int groups_per_split = (int)((n_groups + splits - 1) / splits);
  p.points_per_split = groups_per_split * kPG;
  p.n_splits = (n_points + p.points_per_split - 1) / p.points_per_split;
  p.n_chunks = (int)passes;
  // tensor-pipe dot-product assign: fp16 points, 16-byte rows, <= 8 clusters, big enough to matter
  p.dot_tiles = (esize == 2 && p.vec && n_clusters <= 8 && dim % 8 == 0 && (long long)n_points * dim >= (1ll << 24))
                    ? (int)((dim + kDotTileCols - 1) / kDotTileCols) : 0;
  size_t off = 0;
  p.partial_off = off;
  size_t n_part = (size_t)n_batch * p.n_slices * n_points * n_clusters;
  const size_t n_part_dot = (size_t)n_batch * p.dot_tiles * 4 * n_points * n_clusters;
  if (n_part_dot > n_part) n_part = n_part_dot;
  off = scb_align_up(off + sizeof(float) * n_part, 256);
  p.order_off = off;
  off = scb_align_up(off + sizeof(int) * (size_t)n_batch * n_points, 256);
  p.starts_off = off;
  off = scb_align_up(off + sizeof(int) * (size_t)n_batch * (n_clusters + 1), 256);
  p.shiftp_off = off;
  off = scb_align_up(off + sizeof(float) * (size_t)n_batch * n_clusters * p.n_chunks * (kThreads / 64), 256);
  p.cnorm_off = off;
  off = scb_align_up(off + sizeof(float) * (size_t)n_batch * (p.dot_tiles + 1) * n_clusters, 256);
  p.total = off;
  return p;
}

template <typename T>
int run(const scb_kmeans_step_args* a, cudaStream_t st) {
  const bool aligned = ((uintptr_t)a->points % 16 == 0) && ((uintptr_t)a->centroids % 16 == 0) &&
                       ((a->dim * sizeof(T)) % 16 == 0) && ((a->dim * sizeof(float)) % 16 == 0) &&
                       (!a->out_centroids || (uintptr_t)a->out_centroids % 16 == 0);
  Plan p = make_plan(a->n_points, a->n_clusters, a->n_batch, a->dim, (int)sizeof(T), aligned);
  SCB_REQUIRE(a->workspace_bytes >= p.total, SCB_ERR_WORKSPACE,
              "scb_kmeans_step: workspace %zu B < required %zu B", a->workspace_bytes, p.total);
  SCB_REQUIRE(p.n_splits <= 65535 && a->n_batch <= 65535, SCB_ERR_INVALID_ARG,
              "scb_kmeans_step: grid too large");
  char* ws = reinterpret_cast<char*>(a->workspace);
  float* partial = reinterpret_cast<float*>(ws + p.partial_off);
  int* order = reinterpret_cast<int*>(ws + p.order_off);
  int* starts = reinterpret_cast<int*>(ws + p.starts_off);
  float* shiftp = reinterpret_cast<float*>(ws + p.shiftp_off);
  const T* x = reinterpret_cast<const T*>(a->points);

  const bool use_dot = p.dot_tiles > 0 && a->dist_eps == 0.f && a->out_dist == nullptr && sizeof(T) == 2 &&
                       a->n_batch <= 65535;
  if (use_dot) {
    float* cnorm_p = reinterpret_cast<float*>(ws + p.cnorm_off);
    SCB_SET_FUNC_ATTR_ONCE_PER_DEVICE(kmeans_assign_dot_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kDotSmem);
    kmeans_assign_dot_mma<<<dim3(p.dot_tiles, a->n_batch), kDotThreads, kDotSmem, st>>>(
        reinterpret_cast<const __half*>(a->points), a->centroids, a->n_points, a->n_clusters, a->dim, partial, cnorm_p);
    SCB_LAUNCH_OK("kmeans_assign_dot_mma");
    dim3 gl((a->n_points + kThreads / 32 - 1) / (kThreads / 32), a->n_batch);
    kmeans_assign_labels_dot<<<gl, kThreads, 0, st>>>(partial, cnorm_p, a->n_points, a->n_clusters, p.dot_tiles * 4,
                                                      p.dot_tiles, a->out_labels);
    SCB_LAUNCH_OK("kmeans_assign_labels_dot");
  } else {
  dim3 g1(p.n_tiles, p.n_splits, a->n_batch);
  const bool eps = a->dist_eps != 0.f;
#define SCB_KM_ASSIGN(VEC, EPS)                                                                    \
  do {                                                                                             \
    auto kern = kmeans_assign_partial<T, VEC, EPS>;                                                \
    SCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                     (int)p.assign_smem));                                         \
    kern<<<g1, kThreads, p.assign_smem, st>>>(x, a->centroids, a->n_points, a->n_clusters, a->dim, \
                                              p.tile_cols, p.points_per_split, a->dist_eps,        \
                                              partial);                                            \
  } while (0)
#define SCB_KM_ASYNC(EPS, KT)                                                                      \
  do {                                                                                             \
    auto kern = kmeans_assign_partial_async<T, EPS, KT>;                                           \
    SCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                     (int)p.assign_smem));                                         \
    kern<<<g1, kThreads, p.assign_smem, st>>>(x, a->centroids, a->n_points, a->n_clusters, a->dim, \
                                              p.tile_cols, p.points_per_split, a->dist_eps,        \
                                              partial);                                            \
  } while (0)
  if (p.vec) {
    if (eps) {
      if (p.kt_async == 4) SCB_KM_ASYNC(true, 4); else if (p.kt_async == 5) SCB_KM_ASYNC(true, 5); else SCB_KM_ASYNC(true, 8);
    } else {
      if (p.kt_async == 4) SCB_KM_ASYNC(false, 4); else if (p.kt_async == 5) SCB_KM_ASYNC(false, 5); else SCB_KM_ASYNC(false, 8);
    }
  } else {
    if (eps) SCB_KM_ASSIGN(false, true); else SCB_KM_ASSIGN(false, false);
  }
#undef SCB_KM_ASSIGN
#undef SCB_KM_ASYNC
  SCB_LAUNCH_OK("kmeans_assign_partial");

  dim3 gl((a->n_points + kThreads / 32 - 1) / (kThreads / 32), a->n_batch);
  kmeans_assign_labels<<<gl, kThreads, 0, st>>>(partial, a->n_points, a->n_clusters, p.n_slices, a->out_labels,
                                                a->out_dist);
  SCB_LAUNCH_OK("kmeans_assign_labels");
  }
  if (!a->do_update) return SCB_OK;
  kmeans_assign_sort<<<a->n_batch, kThreads, sizeof(int) * (size_t)a->n_points, st>>>(
      a->out_labels, a->weights, a->n_points, a->n_clusters, a->out_wsum, order, starts);
  SCB_LAUNCH_OK("kmeans_assign_sort");

  // (64-thread blocks — ~1000 blocks, all resident — were measured slower than 256-thread ones: 153 vs 122 us)
  const bool small_blocks = false;
  const int upd_threads = small_blocks ? 64 : kThreads;
  const int upd_chunks = small_blocks ? p.n_chunks * (kThreads / 64) : p.n_chunks;
  dim3 g2(upd_chunks, a->n_batch);
  size_t smem = sizeof(int) * (size_t)a->n_points;
  if (small_blocks)
    kmeans_update_partial<T, true, 64><<<g2, upd_threads, smem, st>>>(
        x, a->weights, a->centroids, order, starts, a->out_wsum, a->n_points, a->n_clusters, a->dim,
        a->out_centroids, shiftp);
  else if (p.vec)
    kmeans_update_partial<T, true, kThreads><<<g2, kThreads, smem, st>>>(
        x, a->weights, a->centroids, order, starts, a->out_wsum, a->n_points, a->n_clusters, a->dim,
        a->out_centroids, shiftp);
  else
    kmeans_update_partial<T, false, kThreads><<<g2, kThreads, smem, st>>>(
        x, a->weights, a->centroids, order, starts, a->out_wsum, a->n_points, a->n_clusters, a->dim,
        a->out_centroids, shiftp);
  SCB_LAUNCH_OK("kmeans_update_partial");
  kmeans_update_final<<<a->n_batch * a->n_clusters, 32, 0, st>>>(shiftp, upd_chunks, a->out_shift);
  SCB_LAUNCH_OK("kmeans_update_final");
  return SCB_OK;
}

}  // namespace

extern "C" {

size_t scb_kmeans_step_workspace_bytes(int32_t n_points, int32_t n_clusters, int32_t n_batch,
                                       int64_t dim) {
  if (n_points < 1 || n_clusters < 1 || n_batch < 1 || dim < 1) return 0;
  // upper bound over both element sizes and both load paths
  size_t m = 0;
  for (int esize = 2; esize <= 4; esize += 2)
    for (int al = 0; al < 2; ++al) {
      size_t t = make_plan(n_points, n_clusters, n_batch, dim, esize, al != 0).total;
      if (t > m) m = t;
    }
  return m;
}

int scb_kmeans_step(const scb_kmeans_step_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_kmeans_step: args is NULL");
  SCB_REQUIRE(a->n_points >= 1 && a->n_points <= 12000, SCB_ERR_INVALID_ARG,
              "scb_kmeans_step: n_points=%d not in 1..12000", a->n_points);
  SCB_REQUIRE(a->n_clusters >= 1 && a->n_clusters <= kMaxClusters, SCB_ERR_INVALID_ARG,
              "scb_kmeans_step: n_clusters=%d not in 1..%d", a->n_clusters, kMaxClusters);
  SCB_REQUIRE(a->n_batch >= 1 && a->n_batch <= 65535, SCB_ERR_INVALID_ARG,
              "scb_kmeans_step: n_batch=%d not in 1..65535", a->n_batch);
  SCB_REQUIRE(a->dim >= 1, SCB_ERR_INVALID_ARG, "scb_kmeans_step: dim < 1");
  SCB_REQUIRE(a->points && a->centroids && a->out_labels && a->workspace, SCB_ERR_INVALID_ARG,
              "scb_kmeans_step: NULL buffer");
  SCB_REQUIRE(!a->do_update || (a->out_centroids && a->out_wsum && a->out_shift),
              SCB_ERR_INVALID_ARG, "scb_kmeans_step: do_update needs out_centroids/out_wsum/out_shift");
  SCB_REQUIRE(a->points_dtype == SCB_F32 || a->points_dtype == SCB_F16, SCB_ERR_UNSUPPORTED,
              "scb_kmeans_step: points dtype %d unsupported", a->points_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->points_dtype == SCB_F32) return run<float>(a, st);
  return run<__half>(a, st);
}

// the name SURVEY.md §8(b) lists for this op
int scb_kmeans_assign_update(const scb_kmeans_step_args* a, void* stream) { return scb_kmeans_step(a, stream); }

}  // extern "C"
