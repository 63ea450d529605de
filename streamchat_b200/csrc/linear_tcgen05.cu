// linear_tcgen05.cu — out = act(alpha * A·Wᵀ + bias) + residual on the 5th-gen tensor cores.
//
// Replaces every nn.Linear the reference reaches through HF transformers on the hot path
// (CLIP q/k/v/out/fc1/fc2 via longva/model/multimodal_encoder/clip_encoder.py:76, the
// mm_projector longva/model/multimodal_projector/builder.py:41-48, Qwen2 q/k/v/o/gate/up/down
// and lm_head via longva/model/language_model/llava_qwen.py:155).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   : TMA producer — cp.async.bulk.tensor 3-D tiles (k, rows, batch) of A [128 x 64]
//              and W [BN x 64] fp16 into a kStages-deep shared-memory ring (128-byte swizzle).
//   warp 1   : MMA issuer — one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//              K=16) x4 per stage; accumulators live in TMEM (2 stages x BN fp32 columns), smem
//              slots are released with tcgen05.commit -> mbarrier.
//   warps 2-5: epilogue — tcgen05.ld 32 lanes x 32 columns, alpha/bias/activation/residual in
//              fp32 registers, 16-byte stores; overlaps the next tile's main loop through the
//              double-buffered TMEM accumulator.
// FLOPs per launch: 2*m*n*k*batch.  Bytes are secondary (operands stream through L2).
#include "scb_common.cuh"
#include <math.h>
#include <mutex>
#include <stdlib.h>
#include <unordered_map>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int kGemmThreads = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue (two per TMEM lane quadrant)

constexpr int kStoreBoxBytes = 32 * 64 * 2;

// STAGED: fp16 rows leave through shared-memory boxes + TMA stores (one ring stage is traded for the staging buffers);
// otherwise (fp32 split-K partials, SiLU-pair outputs) every thread stores its own row pieces and the ring keeps 6 stages.
template <int BN, bool CG2 = false, bool STAGED = false>
struct Cfg {
  // CG2 (CTA pair, cta_group::2 MMA): a CTA stages only ITS half of the W tile -> 32 KB per stage, 6 stages
  static constexpr int kStages = (BN == 256 && !CG2) ? 4 : (STAGED ? 5 : 6);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (CG2 ? BN / 2 : BN) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32)
  // output staging: every epilogue warp owns kStoreBufs boxes of 32 rows x 64 columns fp16 (4 KB, 128-byte swizzle)
  static constexpr int kStoreBufs = (BN == 256 && !CG2) ? 1 : 2;
  static constexpr int kStoreBytes = STAGED ? 8 * kStoreBufs * kStoreBoxBytes : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Split-K with the reduction fused into the GEMM (TAIL instantiation; decode steps of wide batches).  Every CTA owns
// ONE (n tile, K slice) and stores its fp32 partial; the CTA that arrives LAST at the tile's ticket sums the slices in
// index order (deterministic), applies bias / residual, stores the fp16 rows and the tile's per-row sums of squares.
// With a post-norm, the finishing CTAs of a row block then meet at a second counter (they are all resident: the host
// launches at most one CTA per SM and one tile per CTA) and each normalises ITS columns of the finished rows
// (Qwen2RMSNorm, modeling_qwen2.py:249-263) — what took a GEMM launch + a one-block-per-row reduce launch is one launch.
struct SplitKTail {
  __half* out;
  const __half* bias;
  const __half* residual;
  const __half* gamma;     // post-norm weight or nullptr
  __half* pn_out;
  long long ldo, ldr, pn_ld;
  unsigned* tile_tickets;  // [tiles_m * tiles_n] zero before the first launch; left zero
  unsigned* row_arrive;    // [tiles_m]            "
  unsigned* row_depart;    // [tiles_m]            "
  float* ssq;              // [tiles_m * BM, tiles_n] sums of squares of the stored fp16 values
  float eps;
};

struct EpiParams {
  SplitKTail tail;
  const __half* bias;
  const __half* residual;
  void* out;
  long long m, n, k;
  long long ldr, ldo;
  long long stride_r, stride_o;
  int act;
  int out_f32;
  int batch;
  int res_vec, out_vec;  // 16-byte vector access is legal for residual / out rows
  int gm;                // rasterisation: M tiles per group (tm fastest inside a group, then tn, then the next group)
  float alpha;
};

SCB_DEVICE float apply_act(float v, int act) {
  switch (act) {
    case SCB_ACT_QUICK_GELU:
      return __fdividef(v, 1.f + __expf(-1.702f * v));   // IEEE division made the epilogue the bottleneck
    case SCB_ACT_GELU_ERF:
      return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case SCB_ACT_SILU:
      return __fdividef(v, 1.f + __expf(-v));
    default:
      return v;
  }
}

// Tile order.  Consecutive tile indices run over the M tiles of a GROUP of gm M-tiles first, then over N,
// then to the next group: the ~148 tiles in flight share gm A row-panels and ~148/gm W panels through L2.
// gm = tiles_m (one group) streams W once and re-reads A per N tile — right when A fits L2 (prefill
// gate|up); a small gm reads A once and re-reads W per group — right when A is the big operand (down_proj,
// CLIP fc2 at 64 frames: the A re-reads of the one-group order made those HBM-bound at ~5 TB/s).
SCB_DEVICE void tile_coords(int r, int tiles_m, int tiles_n, int gm, int& tm, int& tn) {
  const int per_group = gm * tiles_n;
  const int grp = r / per_group, rem = r - grp * per_group;
  const int m0 = grp * gm;
  const int gsize = min(gm, tiles_m - m0);
  tn = rem / gsize;
  tm = m0 + rem - tn * gsize;
}

// ---- CTA-pair variant (CG2): the two CTAs of a cluster compute ONE 256 x BN tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256): each CTA stages its own 128 rows of A and its own HALF of the W
// tile, the leader CTA (rank 0) issues the MMAs, which read A from both CTAs' shared memory and each half of
// W once — per CTA and k-block 32 KB land in shared memory and 8 KB are read per MMA instead of 48 / 12, which
// is what the single-CTA tile is bound by (128 B/clk of shared-memory bandwidth against 96 B/clk of operand
// reads + 96 B/clk of TMA writes).  Accumulator rows stay with their CTA (128 TMEM lanes each), so the
// epilogue is unchanged.  Cross-CTA signalling: TMA completions of both CTAs count on the LEADER's full
// barrier, tcgen05.commit multicasts slot releases / accumulator-ready to both CTAs, the peer's epilogue
// warps arrive on the leader's accumulator-empty barrier through the cluster address space.
SCB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SCB_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
SCB_DEVICE uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
SCB_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load of a CTA pair member: data into THIS CTA's shared memory, bytes counted on `bar_cluster_addr`
// (a shared::cluster address — the leader's barrier)
SCB_DEVICE void tma_load_3d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0,
                                int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
SCB_DEVICE void umma_f16_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
SCB_DEVICE void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
SCB_DEVICE void tmem_alloc_cg2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
SCB_DEVICE void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
SCB_DEVICE void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

SCB_DEVICE void tma_store_3d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
SCB_DEVICE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
SCB_DEVICE void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
SCB_DEVICE void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

SCB_DEVICE void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps only
SCB_DEVICE unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BN, bool PAIR, bool TAIL = false, bool STAGED = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_o, const EpiParams ep) {
  using C = Cfg<BN, PAIR, STAGED>;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte aligned tile bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint8_t* smem_o = smem + C::kStages * C::kStageBytes;      // [8 warps][kStoreBufs][32 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + C::kStoreBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + C::kStages;         // [kStages]
  uint64_t* tfull = bars + 2 * C::kStages;     // [2]
  uint64_t* tempty = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m_real = (int)((ep.m + BM - 1) / BM);
  // PAIR: the schedule runs over PAIRS of M tiles (one per CTA of the cluster); a missing second tile of an odd
  // count reads zero rows and stores nothing
  const int tiles_m = PAIR ? (tiles_m_real + 1) / 2 : tiles_m_real;
  const int tiles_n = (int)((ep.n + BN - 1) / BN);
  const long long tiles_per_batch = (long long)tiles_m * tiles_n;
  const long long n_tiles = tiles_per_batch * ep.batch;
  const int k_blocks = (int)((ep.k + BK - 1) / BK);
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const long long t_first = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long t_step = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (STAGED) tma_prefetch_desc(&map_o);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], PAIR ? 16 : 8);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_cg2(tmem_slot, C::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global memory and may overlap the tail of the previous kernel on the stream; nothing below runs before that
  // kernel has completed.  In the decode step's launch chain (8 launches per layer for wide batches) this hides the
  // launch latency of every GEMM behind its predecessor.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = t_first; t < n_tiles; t += t_step) {
        const int b = (int)(t / tiles_per_batch);
        const int r = (int)(t - (long long)b * tiles_per_batch);
        int tm, tn;
        tile_coords(r, tiles_m, tiles_n, ep.gm, tm, tn);
        if (PAIR) tm = 2 * tm + crank;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (PAIR) {
            // both CTAs' bytes count on the leader's barrier (armed by the leader for the pair)
            if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            tma_load_3d_cg2(smem_a + stage * C::kABytes, &map_a, lbar, kb * BK, tm * BM, b);
            tma_load_3d_cg2(smem_b + stage * C::kBBytes, &map_w, lbar, kb * BK, tn * BN + crank * (BN / 2), b);
          } else {
            mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
            tma_load_3d(smem_a + stage * C::kABytes, &map_a, &full[stage], kb * BK, tm * BM, b);
            tma_load_3d(smem_b + stage * C::kBBytes, &map_w, &full[stage], kb * BK, tn * BN, b);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!PAIR || crank == 0)) {
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long long t = t_first; t < n_tiles; t += t_step) {
        mbar_wait(&tempty[as], aphase ^ 1);  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * C::kABytes));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * C::kBBytes));
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the >>4 address field
            if (PAIR)
              umma_f16_ss_cg2(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
            else
              umma_f16_ss(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          if (PAIR) umma_commit_cg2(&empty[stage], (uint16_t)3);   // frees the slot in BOTH CTAs
          else umma_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (PAIR) umma_commit_cg2(&tfull[as], (uint16_t)3);   // accumulator complete -> both CTAs' epilogues
        else umma_commit(&tfull[as]);  // accumulator complete -> epilogue
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quadrant, each draining one half of the tile's columns: with one warp per quadrant the
    // epilogue of a K = 1024 tile (CLIP qkv / out / fc1: bias, quick-GELU or residual over 128 x 256 outputs per
    // CTA) took as long as its main loop and the tensor pipe idled 40-60 % of the time (ncu: 37-58 % active against
    // 94 % for K = 4096).
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;  // column half of the tile this warp drains
    int n_box = 0;                      // staging boxes this warp has handed to TMA stores
    int as = 0;
    uint32_t aphase = 0;
    for (long long t = t_first; t < n_tiles; t += t_step) {
      const int b = (int)(t / tiles_per_batch);
      const int r = (int)(t - (long long)b * tiles_per_batch);
      int tm, tn;
      tile_coords(r, tiles_m, tiles_n, ep.gm, tm, tn);
      if (PAIR) tm = 2 * tm + crank;
      const long long row = (long long)tm * BM + quad * 32 + lane;
      const bool row_ok = row < ep.m;
      // Residual rows are fetched one 32-column chunk AHEAD (the first one before the accumulator is even complete):
      // loaded where they are used, each chunk exposed a full L2 round trip inside the tile's epilogue chain and the
      // residual GEMMs of the CLIP tower (out_proj, fc2) were epilogue-bound by it.
      const bool res_ahead = ep.residual != nullptr && ep.res_vec && row_ok;
      const __half* res_row = ep.residual ? ep.residual + (long long)b * ep.stride_r + row * ep.ldr + (long long)tn * BN : nullptr;
      uint4 rcur[4], rnxt[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) rcur[q] = rnxt[q] = make_uint4(0u, 0u, 0u, 0u);
      if (res_ahead && (long long)tn * BN + chalf * (BN / 2) + 32 <= ep.n) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rcur[q] = *reinterpret_cast<const uint4*>(res_row + chalf * (BN / 2) + q * 8);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        const long long col0 = (long long)tn * BN + c0;
        if (col0 >= ep.n) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c0, v);
        if (res_ahead && c0 + 32 < (chalf + 1) * (BN / 2) && col0 + 64 <= ep.n) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rnxt[q] = *reinterpret_cast<const uint4*>(res_row + c0 + 32 + q * 8);
        }
        if (STAGED && (c0 & 32) == 0) {
          // first half of a 64-column box: its staging buffer must have been read by the store issued kStoreBufs boxes ago
          if (n_box >= C::kStoreBufs && lane == 0) bulk_wait_group_read<C::kStoreBufs - 1>();
          __syncwarp();
        }
        tmem_ld_wait();
        if (row_ok || STAGED) {
          const bool full_chunk = col0 + 32 <= ep.n;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * ep.alpha;
          if (ep.bias) {
            if (full_chunk) {
              const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 bv = __ldg(bp + q);
                float2 b0 = h2_to_f2(bv.x), b1 = h2_to_f2(bv.y), b2 = h2_to_f2(bv.z),
                       b3 = h2_to_f2(bv.w);
                f[q * 8 + 0] += b0.x; f[q * 8 + 1] += b0.y; f[q * 8 + 2] += b1.x;
                f[q * 8 + 3] += b1.y; f[q * 8 + 4] += b2.x; f[q * 8 + 5] += b2.y;
                f[q * 8 + 6] += b3.x; f[q * 8 + 7] += b3.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.n) f[j] += __half2float(ep.bias[col0 + j]);
            }
          }
          // one tight loop per activation: a per-element switch unrolled 32x (with erff inlined in
          // every copy) overflowed the instruction cache and made the epilogue 9x slower than the MMA
          if (ep.act == SCB_ACT_QUICK_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.f + __expf(-1.702f * f[j]));
          } else if (ep.act == SCB_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752440f));
          } else if (ep.act == SCB_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.f + __expf(-f[j]));
          }
          if (ep.residual && row_ok) {
            const __half* rp = ep.residual + (long long)b * ep.stride_r + row * ep.ldr + col0;
            if (full_chunk && ep.res_vec) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 rv = rcur[q];
                float2 r0 = h2_to_f2(rv.x), r1 = h2_to_f2(rv.y), r2 = h2_to_f2(rv.z),
                       r3 = h2_to_f2(rv.w);
                f[q * 8 + 0] += r0.x; f[q * 8 + 1] += r0.y; f[q * 8 + 2] += r1.x;
                f[q * 8 + 3] += r1.y; f[q * 8 + 4] += r2.x; f[q * 8 + 5] += r2.y;
                f[q * 8 + 6] += r3.x; f[q * 8 + 7] += r3.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.n) f[j] += __half2float(rp[j]);
            }
          }
          if (ep.act == SCB_ACT_SILU_PAIR) {
            // (gate, up) in adjacent columns -> 16 gated outputs per 32-column chunk
            __half* op = reinterpret_cast<__half*>(ep.out) + (long long)b * ep.stride_o +
                         row * ep.ldo + (col0 >> 1);
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float gt = __half2float(__float2half_rn(f[2 * j]));
              const float up = __half2float(__float2half_rn(f[2 * j + 1]));
              const float sg = __half2float(__float2half_rn(__fdividef(gt, 1.f + __expf(-gt))));
              o[j] = sg * up;
            }
            if (full_chunk && ep.out_vec) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint4 ov;
                ov.x = f2_to_h2(o[q * 8 + 0], o[q * 8 + 1]);
                ov.y = f2_to_h2(o[q * 8 + 2], o[q * 8 + 3]);
                ov.z = f2_to_h2(o[q * 8 + 4], o[q * 8 + 5]);
                ov.w = f2_to_h2(o[q * 8 + 6], o[q * 8 + 7]);
                *reinterpret_cast<uint4*>(op + q * 8) = ov;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + 2 * j + 1 < ep.n) op[j] = __float2half_rn(o[j]);
            }
          } else if (ep.out_f32) {
            float* op = reinterpret_cast<float*>(ep.out) + (long long)b * ep.stride_o +
                        row * ep.ldo + col0;
            if (full_chunk && ep.out_vec) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(op + q * 4) =
                    make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.n) op[j] = f[j];
            }
          } else if (STAGED) {
            // fp16 row piece (64 B) into the warp's staging box [32 rows][128 B], 16-byte chunks XOR-swizzled by the
            // row (the layout SWIZZLE_128B tensor maps expect; conflict-free for a warp storing one chunk per row).
            // Per-thread global stores of 16 B at a row stride of 2-8 KB made the stores the bottleneck of every
            // K = 1024 GEMM of the CLIP tower (qkv 302 us with, 234 us without them).
            uint8_t* box = smem_o + ((warp - 2) * C::kStoreBufs + (n_box % C::kStoreBufs)) * kStoreBoxBytes + lane * 128;
            const int ch0 = (c0 & 32) >> 3;      // first 16-byte chunk of this half inside the 128-byte row
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = f2_to_h2(f[q * 8 + 0], f[q * 8 + 1]);
              o.y = f2_to_h2(f[q * 8 + 2], f[q * 8 + 3]);
              o.z = f2_to_h2(f[q * 8 + 4], f[q * 8 + 5]);
              o.w = f2_to_h2(f[q * 8 + 6], f[q * 8 + 7]);
              *reinterpret_cast<uint4*>(box + (((ch0 + q) ^ (lane & 7)) << 4)) = o;
            }
          } else {
            __half* op = reinterpret_cast<__half*>(ep.out) + (long long)b * ep.stride_o +
                         row * ep.ldo + col0;
            if (full_chunk && ep.out_vec) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 o;
                o.x = f2_to_h2(f[q * 8 + 0], f[q * 8 + 1]);
                o.y = f2_to_h2(f[q * 8 + 2], f[q * 8 + 3]);
                o.z = f2_to_h2(f[q * 8 + 4], f[q * 8 + 5]);
                o.w = f2_to_h2(f[q * 8 + 6], f[q * 8 + 7]);
                *reinterpret_cast<uint4*>(op + q * 8) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.n) op[j] = __float2half_rn(f[j]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) rcur[q] = rnxt[q];
        if (STAGED && ((c0 & 32) != 0 || col0 + 32 >= ep.n)) {
          // box complete (or the row ends inside it): one TMA store of 32 rows x 64 columns; rows >= m and columns
          // >= n are clipped by the tensor map
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o, smem_o + ((warp - 2) * C::kStoreBufs + (n_box % C::kStoreBufs)) * kStoreBoxBytes,
                         (int)((long long)tn * BN + (c0 & ~63)), (int)((long long)tm * BM + quad * 32), b);
            bulk_commit_group();
          }
          ++n_box;
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): release the stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));   // the leader's MMA warp waits
        else mbar_arrive(&tempty[as]);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
      if constexpr (TAIL) {
        static_assert(!TAIL || (BN == 128 && !PAIR), "the fused split-K tail maps one warp onto the 128 columns of a tile");
        __shared__ unsigned tail_ticket;
        const SplitKTail& tk = ep.tail;
        const int e = (int)threadIdx.x - 64, ew = e >> 5;   // epilogue thread / warp index
        __threadfence();                                    // this warp's partial stores, device-wide
        epi_bar_sync();
        if (e == 0) tail_ticket = atomicAdd(&tk.tile_tickets[tm * tiles_n + tn], 1u);
        epi_bar_sync();
        if (tail_ticket == (unsigned)(ep.batch - 1)) {      // CTA-uniform: every slice of this tile is stored
          __threadfence();
          if (e == 0) tk.tile_tickets[tm * tiles_n + tn] = 0u;
          const int rows_here = (int)min((long long)BM, ep.m - (long long)tm * BM);
          const long long col = (long long)tn * BN + lane * 4;
          const bool col_ok = col < ep.n;                   // n is a multiple of 4
          const float* part = reinterpret_cast<const float*>(ep.out);
          uint2 keep[BM / 8];
#pragma unroll
          for (int i = 0; i < BM / 8; ++i) keep[i] = make_uint2(0u, 0u);
          // warp ew: rows ew, ew + 8, ...; lane: 4 columns.  Four rows x four slices of loads are in flight at a time
          // (one L2 round trip per 16 partial vectors; issued one by one they cost 16 round trips for a 25-row batch)
#pragma unroll
          for (int i0 = 0; i0 < BM / 8; i0 += 4) {
            if (ew + 8 * i0 < rows_here) {                  // warp-uniform
              float4 acc[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int sp0 = 0; sp0 < ep.batch; sp0 += 4) {
                float4 v[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int rl = min(ew + 8 * (i0 + q), rows_here - 1);          // clamped: loads stay unconditional
                  const long long rr = (long long)tm * BM + rl;
#pragma unroll
                  for (int d = 0; d < 4; ++d) {
                    const int sp = min(sp0 + d, ep.batch - 1);
                    v[q][d] = __ldcg(reinterpret_cast<const float4*>(part + (long long)sp * ep.stride_o + rr * ep.ldo +
                                                                     (col_ok ? col : 0)));
                  }
                }
#pragma unroll
                for (int d = 0; d < 4; ++d)                 // slices in index order
                  if (sp0 + d < ep.batch) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      acc[q].x += v[q][d].x; acc[q].y += v[q][d].y; acc[q].z += v[q][d].z; acc[q].w += v[q][d].w;
                    }
                  }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int rl = ew + 8 * (i0 + q);
                if (rl < rows_here) {                       // warp-uniform
                  const long long rr = (long long)tm * BM + rl;
                  float ss = 0.f;
                  if (col_ok) {
                    float4 a4 = acc[q];
                    if (tk.bias) {
                      const uint2 bq = *reinterpret_cast<const uint2*>(tk.bias + col);
                      const float2 b0 = h2_to_f2(bq.x), b1 = h2_to_f2(bq.y);
                      a4.x += b0.x; a4.y += b0.y; a4.z += b1.x; a4.w += b1.y;
                    }
                    if (tk.residual) {
                      const uint2 rq = *reinterpret_cast<const uint2*>(tk.residual + rr * tk.ldr + col);
                      const float2 r0 = h2_to_f2(rq.x), r1 = h2_to_f2(rq.y);
                      a4.x += r0.x; a4.y += r0.y; a4.z += r1.x; a4.w += r1.y;
                    }
                    uint2 o;
                    o.x = f2_to_h2(a4.x, a4.y);
                    o.y = f2_to_h2(a4.z, a4.w);
                    *reinterpret_cast<uint2*>(tk.out + rr * tk.ldo + col) = o;
                    keep[i0 + q] = o;
                    const float2 h0 = h2_to_f2(o.x), h1 = h2_to_f2(o.y);   // the norm sees the stored fp16 values
                    ss = h0.x * h0.x + h0.y * h0.y + h1.x * h1.x + h1.y * h1.y;
                  }
                  if (tk.pn_out) {
                    ss = warp_sum(ss);
                    if (lane == 0) tk.ssq[rr * tiles_n + tn] = ss;
                  }
                }
              }
            }
          }
          if (tk.pn_out) {
            __threadfence();
            epi_bar_sync();
            if (e == 0) {
              atomicAdd(&tk.row_arrive[tm], 1u);
              unsigned spins = 0;
              while (ld_acquire_gpu(&tk.row_arrive[tm]) < (unsigned)tiles_n)
                if (++spins > (1u << 27)) __trap();         // a finishing CTA never arrived: fail loudly, do not hang
            }
            epi_bar_sync();
#pragma unroll
            for (int i = 0; i < BM / 8; ++i) {
              const int rl = ew + 8 * i;
              if (rl < rows_here) {
                const long long rr = (long long)tm * BM + rl;
                float ss = 0.f;
                for (int j = lane; j < tiles_n; j += 32) ss += __ldcg(&tk.ssq[rr * tiles_n + j]);
                ss = warp_sum(ss);
                const float rstd = rsqrtf(ss / (float)ep.n + tk.eps);
                if (col_ok) {
                  const uint2 g = *reinterpret_cast<const uint2*>(tk.gamma + col);
                  const float2 x0 = h2_to_f2(keep[i].x), x1 = h2_to_f2(keep[i].y), g0 = h2_to_f2(g.x), g1 = h2_to_f2(g.y);
                  uint2 o;
                  o.x = f2_to_h2(__half2float(__float2half_rn(x0.x * rstd)) * g0.x, __half2float(__float2half_rn(x0.y * rstd)) * g0.y);
                  o.y = f2_to_h2(__half2float(__float2half_rn(x1.x * rstd)) * g1.x, __half2float(__float2half_rn(x1.y * rstd)) * g1.y);
                  *reinterpret_cast<uint2*>(tk.pn_out + rr * tk.pn_ld + col) = o;
                }
              }
            }
            epi_bar_sync();
            if (e == 0 && atomicAdd(&tk.row_depart[tm], 1u) == (unsigned)(tiles_n - 1)) {
              tk.row_arrive[tm] = 0u;                       // every finishing CTA of the row block is past its wait
              tk.row_depart[tm] = 0u;
            }
          }
        }
      }
    }
  }

  if (STAGED && warp >= 2 && lane == 0) bulk_wait_group<0>();   // staged rows are written before the CTA retires
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched through the runtime, so the
// library does not link libcuda) with a small cache keyed on the full descriptor contents.
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long rows, k, ld, stride;
  int batch, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && k == o.k && ld == o.ld && stride == o.stride &&
           batch == o.batch && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](long long v) { h ^= (size_t)v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.k); mix(k.ld); mix(k.stride); mix(k.batch); mix(k.box_rows);
    return h;
  }
};

// [batch, rows, k] fp16 tensor, k contiguous; box = (64, box_rows, 1); 128-byte swizzle;
// out-of-bounds elements read as zero (so ragged m / n / k tails need no special casing).
int make_map(CUtensorMap* out, const void* ptr, long long rows, long long k, long long ld,
             long long stride, int batch, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, k, ld, stride, batch, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return SCB_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  SCB_REQUIRE(fn != nullptr, SCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? stride : rows * ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_REQUIRE(r == CUDA_SUCCESS, SCB_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) rows=%lld k=%lld ld=%lld stride=%lld batch=%d",
              (int)r, rows, k, ld, stride, batch);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return SCB_OK;
}

// fp16 outputs with 16-byte aligned rows leave through shared memory and TMA box stores (SCB_LINEAR_TMA_STORE=0: every
// thread stores its own row pieces)
bool staged_store(const scb_linear_args* a) {
  static const int use_tma_store = [] { const char* e = getenv("SCB_LINEAR_TMA_STORE"); return e ? atoi(e) : 1; }();
  return use_tma_store && a->out_dtype == SCB_F16 && a->act != SCB_ACT_SILU_PAIR && a->ldo % 8 == 0 &&
         (a->batch == 1 || a->stride_o % 8 == 0);
}

template <int BN, bool PAIR, bool TAIL = false, bool STAGED = false>
int launch(const scb_linear_args* a, cudaStream_t st, const SplitKTail* tail = nullptr) {
  using C = Cfg<BN, PAIR, STAGED>;
  CUtensorMap ma, mw;
  int rc = make_map(&ma, a->a, a->m, a->k, a->lda, a->stride_a, a->batch, BM);
  if (rc != SCB_OK) return rc;
  rc = make_map(&mw, a->w, a->n, a->k, a->ldw, a->stride_w, a->batch, PAIR ? BN / 2 : BN);
  if (rc != SCB_OK) return rc;
  constexpr bool tma_store = STAGED;
  CUtensorMap mo = ma;
  if (tma_store) {
    rc = make_map(&mo, a->out, a->m, a->n, a->ldo, a->stride_o, a->batch, 32);
    if (rc != SCB_OK) return rc;
  }
  EpiParams ep;
  ep.tail = tail ? *tail : SplitKTail{};
  ep.bias = reinterpret_cast<const __half*>(a->bias);
  ep.residual = reinterpret_cast<const __half*>(a->residual);
  ep.out = a->out;
  ep.m = a->m; ep.n = a->n; ep.k = a->k;
  ep.ldr = a->ldr; ep.ldo = a->ldo;
  ep.stride_r = a->stride_r; ep.stride_o = a->stride_o;
  ep.act = a->act;
  ep.out_f32 = a->out_dtype == SCB_F32;
  ep.batch = a->batch;
  ep.alpha = a->alpha;
  {
    // A panel bytes decide the order (see tile_coords): one group while A stays L2-resident, groups of 8 otherwise
    static const int force_gm = [] { const char* e = getenv("SCB_LINEAR_GM"); return e ? atoi(e) : 0; }();   // tuning aid
    const long long tiles_m = PAIR ? (((a->m + BM - 1) / BM) + 1) / 2 : (a->m + BM - 1) / BM;
    const long long a_bytes = (long long)a->m * a->k * 2;
    int gm = (a_bytes <= (48ll << 20)) ? (int)tiles_m : (PAIR ? 4 : 8);
    if (force_gm > 0) gm = force_gm;
    if (gm > tiles_m) gm = (int)tiles_m;
    ep.gm = gm < 1 ? 1 : gm;
  }
  const int ov = ep.out_f32 ? 4 : 8;
  ep.out_vec = (a->ldo % ov == 0) && (a->batch == 1 || a->stride_o % ov == 0);
  ep.res_vec = (a->ldr % 8 == 0) && (a->batch == 1 || a->stride_r % 8 == 0);
  auto kern = linear_tcgen05_kernel<BN, PAIR, TAIL, STAGED>;
  SCB_SET_FUNC_ATTR_ONCE_PER_DEVICE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);  // per instantiation
  const long long tm_sched = PAIR ? (((a->m + BM - 1) / BM) + 1) / 2 : (a->m + BM - 1) / BM;
  const long long tiles = tm_sched * ((a->n + BN - 1) / BN) * a->batch;
  const int sms = scb_num_sms();
  if (TAIL) SCB_REQUIRE(tiles <= sms, SCB_ERR_UNSUPPORTED,
                        "scb_linear_splitk: %lld tiles need one resident CTA each, %d SMs (use scb_linear + scb_splitk_reduce)",
                        tiles, sms);
  if (PAIR) {
    const int clusters = (int)(tiles < sms / 2 ? tiles : sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = scb_pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    SCB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ma, mw, mo, ep));
  } else {
    const int grid = (int)(tiles < sms ? tiles : sms);
    SCB_CUDA_OK(scb_launch_pdl(kern, dim3(grid), dim3(kGemmThreads), C::kSmemBytes, st, ma, mw, mo, ep));
  }
  SCB_LAUNCH_OK("linear_tcgen05_kernel");
  return SCB_OK;
}

}  // namespace

extern "C" int scb_linear(const scb_linear_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_linear: args is NULL");
  SCB_REQUIRE(a->m >= 1 && a->n >= 1 && a->k >= 1 && a->batch >= 1, SCB_ERR_INVALID_ARG,
              "scb_linear: m=%lld n=%lld k=%lld batch=%d must be >= 1", (long long)a->m,
              (long long)a->n, (long long)a->k, a->batch);
  SCB_REQUIRE(a->a && a->w && a->out, SCB_ERR_INVALID_ARG, "scb_linear: NULL buffer");
  SCB_REQUIRE(a->lda >= a->k && a->ldw >= a->k &&
                  a->ldo >= (a->act == SCB_ACT_SILU_PAIR ? a->n / 2 : a->n),
              SCB_ERR_INVALID_ARG, "scb_linear: leading dimension too small");
  SCB_REQUIRE(a->act != SCB_ACT_SILU_PAIR ||
                  (a->n % 2 == 0 && !a->bias && !a->residual && a->out_dtype == SCB_F16),
              SCB_ERR_INVALID_ARG,
              "scb_linear: SILU_PAIR needs even n, fp16 out and no bias / residual");
  SCB_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0, SCB_ERR_INVALID_ARG,
              "scb_linear: lda/ldw must be multiples of 8 elements (TMA 16-byte strides)");
  SCB_REQUIRE((uintptr_t)a->a % 16 == 0 && (uintptr_t)a->w % 16 == 0 && (uintptr_t)a->out % 16 == 0,
              SCB_ERR_INVALID_ARG, "scb_linear: a/w/out must be 16-byte aligned");
  SCB_REQUIRE(a->batch == 1 || (a->stride_a % 8 == 0 && a->stride_w % 8 == 0), SCB_ERR_INVALID_ARG,
              "scb_linear: batch strides must be multiples of 8 elements");
  SCB_REQUIRE(!a->residual || a->ldr >= a->n, SCB_ERR_INVALID_ARG, "scb_linear: ldr too small");
  SCB_REQUIRE(!a->bias || (uintptr_t)a->bias % 16 == 0, SCB_ERR_INVALID_ARG,
              "scb_linear: bias must be 16-byte aligned");
  SCB_REQUIRE(!a->residual || (uintptr_t)a->residual % 16 == 0, SCB_ERR_INVALID_ARG,
              "scb_linear: residual must be 16-byte aligned");
  SCB_REQUIRE(a->out_dtype == SCB_F16 || a->out_dtype == SCB_F32, SCB_ERR_UNSUPPORTED,
              "scb_linear: out dtype %d unsupported", a->out_dtype);
  SCB_REQUIRE(a->act >= SCB_ACT_NONE && a->act <= SCB_ACT_SILU_PAIR, SCB_ERR_UNSUPPORTED,
              "scb_linear: activation %d unsupported", a->act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Narrow problems keep more CTAs busy with 128-wide tiles; wide ones amortise A with 256.
  const long long tiles256 = ((a->m + BM - 1) / BM) * ((a->n + 255) / 256) * a->batch;
  static const int force_bn = [] { const char* e = getenv("SCB_LINEAR_BN"); return e ? atoi(e) : 0; }();   // tuning aid
  static const int use_pair = [] { const char* e = getenv("SCB_LINEAR_PAIR"); return e ? atoi(e) : 1; }();   // 0: single-CTA tiles only
  const bool staged = staged_store(a);
  if (force_bn == 128 || a->n <= 128 || (tiles256 < scb_num_sms() && force_bn != 256))
    return staged ? launch<128, false, false, true>(a, st) : launch<128, false>(a, st);
  // CTA pairs (cta_group::2, UMMA M = 256) whenever there are at least two M tiles to pair up
  if (use_pair && a->m > BM) return staged ? launch<256, true, false, true>(a, st) : launch<256, true>(a, st);
  return staged ? launch<256, false, false, true>(a, st) : launch<256, false>(a, st);
}

// workspace of the fused tail in 4-byte words: tile tickets, row arrive / depart counters, per-(row, n tile) sums of squares
// Fixed offsets, whatever (m, n) a call has: the counters must never share words with the sums of squares an earlier
// call of another shape left behind (only the counters are guaranteed zero between calls).
constexpr long long kTailCounters = 256;   // >= SM count: one tile per CTA, one CTA per SM
static long long splitk_ws_words(long long m, long long n) {
  const long long tm = (m + BM - 1) / BM, tn = (n + 127) / 128;
  return 3 * kTailCounters + tm * BM * tn;
}

extern "C" int64_t scb_linear_splitk_workspace_bytes(int64_t m, int64_t n) {
  return m >= 1 && n >= 1 ? 4 * splitk_ws_words(m, n) : 0;
}

extern "C" int scb_linear_splitk(const scb_linear_splitk_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_linear_splitk: args is NULL");
  SCB_REQUIRE(a->m >= 1 && a->n >= 1 && a->k >= 1 && a->n_splits >= 1, SCB_ERR_INVALID_ARG,
              "scb_linear_splitk: m=%lld n=%lld k=%lld n_splits=%d must be >= 1", (long long)a->m, (long long)a->n,
              (long long)a->k, a->n_splits);
  SCB_REQUIRE(a->a && a->w && a->partial && a->out && a->workspace, SCB_ERR_INVALID_ARG, "scb_linear_splitk: NULL buffer");
  const long long ks = a->k / a->n_splits;
  SCB_REQUIRE(ks * a->n_splits == a->k && ks % 8 == 0, SCB_ERR_INVALID_ARG,
              "scb_linear_splitk: k=%lld cannot be cut into %d slices of a multiple of 8", (long long)a->k, a->n_splits);
  SCB_REQUIRE(a->n % 4 == 0 && a->lda >= a->k && a->ldw >= a->k && a->lda % 8 == 0 && a->ldw % 8 == 0 && a->ldo >= a->n &&
                  a->ldo % 4 == 0 && (!a->residual || (a->ldr >= a->n && a->ldr % 4 == 0)) &&
                  (!a->post_norm_out || (a->post_norm_gamma && a->post_norm_ld >= a->n && a->post_norm_ld % 4 == 0)),
              SCB_ERR_INVALID_ARG, "scb_linear_splitk: n and the leading dimensions must keep rows 8-byte aligned");
  SCB_REQUIRE((uintptr_t)a->a % 16 == 0 && (uintptr_t)a->w % 16 == 0 && (uintptr_t)a->partial % 16 == 0 &&
                  (uintptr_t)a->out % 8 == 0 && (uintptr_t)a->workspace % 4 == 0,
              SCB_ERR_INVALID_ARG, "scb_linear_splitk: misaligned buffer");
  SCB_REQUIRE(a->workspace_bytes >= scb_linear_splitk_workspace_bytes(a->m, a->n), SCB_ERR_INVALID_ARG,
              "scb_linear_splitk: workspace of %lld bytes, %lld needed", (long long)a->workspace_bytes,
              (long long)scb_linear_splitk_workspace_bytes(a->m, a->n));
  const long long tm = (a->m + BM - 1) / BM, tn = (a->n + 127) / 128;
  scb_linear_args la = {};
  la.a = a->a; la.w = a->w; la.out = a->partial;
  la.m = a->m; la.n = a->n; la.k = ks;
  la.lda = a->lda; la.ldw = a->ldw; la.ldo = a->n;
  la.act = SCB_ACT_NONE; la.out_dtype = SCB_F32; la.batch = a->n_splits;
  la.stride_a = ks; la.stride_w = ks; la.stride_o = a->m * a->n;
  la.alpha = 1.f;
  SplitKTail tk = {};
  tk.out = reinterpret_cast<__half*>(a->out);
  tk.bias = reinterpret_cast<const __half*>(a->bias);
  tk.residual = reinterpret_cast<const __half*>(a->residual);
  tk.gamma = reinterpret_cast<const __half*>(a->post_norm_gamma);
  tk.pn_out = reinterpret_cast<__half*>(a->post_norm_out);
  tk.ldo = a->ldo; tk.ldr = a->ldr; tk.pn_ld = a->post_norm_ld;
  unsigned* ws = reinterpret_cast<unsigned*>(a->workspace);
  SCB_REQUIRE(tm * tn * a->n_splits <= kTailCounters, SCB_ERR_UNSUPPORTED,
              "scb_linear_splitk: %lld tiles, at most %lld (use scb_linear + scb_splitk_reduce)", tm * tn * a->n_splits,
              kTailCounters);
  tk.tile_tickets = ws;
  tk.row_arrive = ws + kTailCounters;
  tk.row_depart = ws + 2 * kTailCounters;
  tk.ssq = reinterpret_cast<float*>(ws + 3 * kTailCounters);
  tk.eps = a->post_norm_eps;
  return launch<128, false, true>(&la, reinterpret_cast<cudaStream_t>(stream), &tk);
}
