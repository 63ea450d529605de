// attention_tcgen05.cu — fused softmax(scale·Q·Kᵀ [+causal])·V on tcgen05 tensor cores.
//
// Replaces the SDPA calls the reference reaches through HF transformers: CLIP self-attention
// (non-causal, 16 heads x 64, 577 tokens; modeling_clip.py:282-337 via clip_encoder.py:76) and
// Qwen2 prefill attention (causal, GQA 28q/4kv x 128; modeling_qwen2.py:161-184,227-228 via
// llava_qwen.py:155).  Scores never touch HBM.
//
// One CTA per (128-query tile, q head, batch); 192 threads:
//   warp 0   : TMA producer — Q once, then K and V tiles of kBKV keys (4-D tensor maps
//              (d, token, head, batch), 128-byte swizzle, 64-column sub-tiles) into 2-stage rings.
//   warp 1   : MMA issuer — S = Q·Kᵀ (M=128, N=kBKV, both operands K-major) into TMEM cols
//              [0,kBKV); O_j = P·V (A = P from shared memory, B = V MN-major, N = head_dim) into
//              TMEM cols [128,128+D).  S_{j+1} is issued before P·V_j so it overlaps the softmax.
//   warps 2-5: softmax — thread r owns query row r (TMEM lane r): two passes of tcgen05.ld over
//              the S row (max, then exp2/sum), P written as fp16 into the swizzled A-operand
//              layout, running (m, l) and the fp32 output row kept in registers
//              (O = O*alpha + O_j), final 1/l scaling and 16-byte stores.
// FLOPs per launch: 4 * batch * heads * q_len * kv_len * head_dim (x 1/2 when causal, q_len==kv_len).
#include "scb_common.cuh"
#include <math.h>
#include <mutex>
#include <stdlib.h>
#include <unordered_map>

namespace {

constexpr int kBQ = 128;
constexpr int kAttnThreads = 192;
constexpr int kKVStages = 2;

SCB_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// MN-major (N contiguous) B operand made of 64-column sub-tiles [rows=K][64] with the 128-byte
// swizzle: LBO = byte distance between sub-tiles (next 64 of N), SBO = 1024 (next 8 K rows).
SCB_DEVICE uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

SCB_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A operand from tensor memory (rows = lanes, two fp16 along K per 32-bit column), B from shared memory
SCB_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct AttnParams {
  __half* out;
  long long o_stride_b, o_stride_t, o_stride_h;
  int q_len, kv_len;
  int n_q_heads, n_kv_heads;
  int causal;
  float scale_log2;  // scale * log2(e)
};

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// NSB = number of S buffers in TMEM.  2: S_{j+1} is computed while the softmax works on S_j (one CTA per SM,
// D = 128).  1: S + O fit 256 columns and the tiles 112 KB, so TWO CTAs share an SM and fill each other's
// softmax / MMA gaps with 8 softmax warps per SM instead of 4 (D = 64: the CLIP tower, where the 4-warp
// softmax of a lone CTA left the tensor pipe idle 85 % of the time).
// PT: P stays in tensor memory — the softmax thread of a row writes fp16 P over the first half of the S columns it
// has just read (tcgen05.st) and P·V takes its A operand from there.  The S buffers alternate by block, so P_{j+1} is
// written while P·V_j still reads P_j: the shared-memory P tile was single-buffered, which put every P·V into the
// per-block dependency chain (switching the whole softmax arithmetic off shortened the kernel by only 8 %).
template <int D, int kBKV, int NSB, int KS, bool PT = false>
struct ACfg {
  static constexpr int kSub = D / 64;                  // 64-column sub-tiles per row
  static constexpr int kQBytes = kBQ * D * 2;
  static constexpr int kKBytes = kBKV * D * 2;
  static constexpr int kPBytes = PT ? 0 : kBQ * kBKV * 2;
  static_assert(!PT || KS == 1, "P in tensor memory: one thread per row");
  static constexpr int kThreads = (2 + 4 * KS) * 32;   // TMA warp, MMA warp, 4*KS softmax warps
  static constexpr int kXBytes = KS == 2 ? 3 * 2 * kBQ * 4 : 0;   // max (x2 parities) / sum exchange between the halves of a row
  static constexpr int kSmemBytes = kQBytes + 2 * kKVStages * kKBytes + kPBytes + 192 + kXBytes;
  //   S buffers: [sb*kBKV, (sb+1)*kBKV)   O_j: [NSB*kBKV, NSB*kBKV + D)
  static constexpr int kOCol = NSB * kBKV;
  static constexpr int kTmemCols = pow2_cols(NSB * kBKV + D);
  static constexpr int kBySmem = (228 * 1024) / (kSmemBytes + 1024), kByTmem = 512 / kTmemCols;
  static constexpr int kMinBlocks = kBySmem < kByTmem ? (kBySmem < 3 ? kBySmem : 3) : (kByTmem < 3 ? kByTmem : 3);
};

template <int D, int kBKV, int NSB, int KS, bool PT>
__global__ void __launch_bounds__((ACfg<D, kBKV, NSB, KS, PT>::kThreads), (ACfg<D, kBKV, NSB, KS, PT>::kMinBlocks))
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q,
                         const __grid_constant__ CUtensorMap map_k,
                         const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  using C = ACfg<D, kBKV, NSB, KS, PT>;
  // 128-byte swizzle atoms need 1024-byte aligned tiles; there is no static shared memory in this
  // kernel, so the dynamic window starts aligned (checked below — two CTAs per SM leave no slack).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sq = smem;
  uint8_t* sk = sq + C::kQBytes;
  uint8_t* sv = sk + kKVStages * C::kKBytes;
  uint8_t* sp = sv + kKVStages * C::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp + C::kPBytes);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [2]
  uint64_t* k_empty = bars + 3;            // [2]
  uint64_t* v_full = bars + 5;             // [2]
  uint64_t* v_empty = bars + 7;            // [2]
  uint64_t* s_full = bars + 9;             // [2]
  uint64_t* s_empty = bars + 11;           // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* o_full = bars + 14;
  uint64_t* o_empty = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* xch = reinterpret_cast<float*>(sp + C::kPBytes + 192);   // [3][2][kBQ] (KS == 2 only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: the tiles with the most key blocks are scheduled first (no long tail of late heavy CTAs)
  const int q0 = (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * kBQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.n_q_heads / p.n_kv_heads);
  const int causal_off = p.kv_len - p.q_len;
  int n_blocks = (p.kv_len + kBKV - 1) / kBKV;
  if (p.causal) {
    const int last_key = min(p.kv_len - 1, q0 + kBQ - 1 + causal_off);
    n_blocks = last_key < 0 ? 0 : min(n_blocks, last_key / kBKV + 1);
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int sb = 0; sb < 2; ++sb) {
      mbar_init(&s_full[sb], 1);
      mbar_init(&s_empty[sb], 4 * KS);
    }
    mbar_init(p_full, 4 * KS);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4 * KS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // + kBKV * (j % NSB)
  const uint32_t tmem_o = tmem_base + (uint32_t)C::kOCol;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_blocks > 0) {
      mbar_arrive_expect_tx(q_full, C::kQBytes);
      for (int s = 0; s < C::kSub; ++s)
        tma_load_4d(sq + s * (kBQ * 128), &map_q, q_full, s * 64, q0, h, b);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % kKVStages;
        const uint32_t ph = (uint32_t)(j / kKVStages) & 1u;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], C::kKBytes);
        for (int s = 0; s < C::kSub; ++s)
          tma_load_4d(sk + st * C::kKBytes + s * (kBKV * 128), &map_k, &k_full[st], s * 64,
                      j * kBKV, hk, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], C::kKBytes);
        for (int s = 0; s < C::kSub; ++s)
          tma_load_4d(sv + st * C::kKBytes + s * (kBKV * 128), &map_v, &v_full[st], s * 64,
                      j * kBKV, hk, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_blocks > 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(kBQ, D, 0, 1);  // B = V is MN-major
      auto issue_s = [&](int j) {
        const int st = j % kKVStages, sb = j % NSB;
        mbar_wait(&k_full[st], (uint32_t)(j / kKVStages) & 1u);
        if (j >= NSB) mbar_wait(&s_empty[sb], (uint32_t)((j / NSB) - 1) & 1u);   // softmax done with S_{j-NSB}
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da =
              umma_desc_k_sw128(smem_u32(sq + (kk / 4) * (kBQ * 128))) + (uint64_t)((kk % 4) * 2);
          const uint64_t db =
              umma_desc_k_sw128(smem_u32(sk + st * C::kKBytes + (kk / 4) * (kBKV * 128))) +
              (uint64_t)((kk % 4) * 2);
          umma_f16_ss(tmem_s + (uint32_t)(sb * kBKV), da, db, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < NSB && j < n_blocks; ++j) issue_s(j);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % kKVStages;
        mbar_wait(&v_full[st], (uint32_t)(j / kKVStages) & 1u);
        mbar_wait(p_full, (uint32_t)j & 1u);     // P_j written (and O rescaled if the maximum jumped)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          const uint64_t db = umma_desc_mn_sw128(smem_u32(sv + st * C::kKBytes + kk * 2048),
                                                 (uint32_t)(kBKV * 128));
          if (PT) {
            // P_j lives in the columns of S_j: 16 keys = 8 columns per MMA
            umma_f16_ts(tmem_o, tmem_s + (uint32_t)((j % NSB) * kBKV + kk * 8), db, idesc_o, (j | kk) != 0 ? 1u : 0u);
          } else {
            const uint64_t da =
                umma_desc_k_sw128(smem_u32(sp + (kk / 4) * (kBQ * 128))) + (uint64_t)((kk % 4) * 2);
            umma_f16_ss(tmem_o, da, db, idesc_o, (j | kk) != 0 ? 1u : 0u);   // O accumulates over the blocks
          }
        }
        umma_commit(&v_empty[st]);
        umma_commit(o_full);
        if (j + NSB < n_blocks) issue_s(j + NSB);   // keeps the tensor pipe busy under the next softmax
      }
    }
  } else {
    // ===================== softmax + output (warps 2 .. 2+4*KS) =====================
    // KS threads share a query row: thread (r, half) owns kBKV/KS score columns and D/KS output dims, so
    // the per-block softmax work of a row is split over two warps of the same TMEM lane quadrant (KS = 2:
    // 8 softmax warps per CTA, two per scheduler — the lone warp per scheduler of KS = 1 could not hide its
    // own tcgen05.ld / MUFU latencies).  The halves exchange the block maximum (and the row sum at the end).
    constexpr int CW = kBKV / KS, DH = D / KS;
    const int quad = warp & 3;             // TMEM lane quadrant a warp may access = warp id mod 4
    const int half = (warp - 2) >> 2;      // 0 .. KS-1
    const int r = quad * 32 + lane;        // row inside the query tile == TMEM lane
    const int qpos = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_o_own = tmem_o + lane_addr + (uint32_t)(half * DH);
    // TMEM reads are the scarce resource of this kernel (64 B/clk/SM: reading S twice and O once per block
    // cost ~3000 of the ~4000 cycles a block took).  So: the S row is read ONCE into registers, and O stays
    // in TMEM as the accumulator of the P·V MMAs — it is only touched when the running maximum moved by more
    // than 2^8 (lazy rescale: P is computed against a slightly stale maximum, exact after the final 1/l).
    float m_ref = -INFINITY, l_run = 0.f;
    const int key_limit = p.causal ? min(p.kv_len - 1, qpos + causal_off) : p.kv_len - 1;

    for (int j = 0; j < n_blocks; ++j) {
      const int sb = j % NSB;
      const uint32_t tmem_sj = tmem_s + (uint32_t)(sb * kBKV) + lane_addr + (uint32_t)(half * CW);
      mbar_wait(&s_full[sb], (uint32_t)(j / NSB) & 1u);
      tc_fence_after();
      const int key0 = j * kBKV + half * CW;
      uint32_t sv[CW];
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 32) tmem_ld_32x32b_x32(tmem_sj + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&sv[c0]));
      tmem_ld_wait();
      // interior blocks (every key visible to every row of the warp) skip the mask
      if (!__all_sync(0xffffffffu, key0 + CW - 1 <= key_limit)) {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (key0 + i > key_limit) sv[i] = 0xff800000u;   // -inf
      }
      float m_blk = -INFINITY;
#pragma unroll
      for (int i = 0; i < CW; i += 2) m_blk = fmaxf(m_blk, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
      if (KS == 2) {
        // the two halves of a row agree on the block maximum (double-buffered by block parity)
        float* mx = xch + (j & 1) * (2 * kBQ);
        mx[half * kBQ + r] = m_blk;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        m_blk = fmaxf(m_blk, mx[(half ^ 1) * kBQ + r]);
      }
      const float m_new = fmaxf(m_ref, m_blk);
      // rescale O (and l) only when the maximum moved by more than 2^8; identical decision in both halves
      const bool need = (m_new - m_ref) * p.scale_log2 > 8.f;   // also true for m_ref == -inf < m_new
      bool o_seen = false;     // has this thread observed the completion of P·V_{j-1} in this block?
      if (j == 0) {
        m_ref = m_new;
      } else if (__any_sync(0xffffffffu, need)) {
        o_seen = true;
        const float alpha = need ? exp2f((m_ref - m_new) * p.scale_log2) : 1.f;   // m_ref == -inf -> 0 (O row is 0)
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);        // P·V_{j-1} has landed in O
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_o_own + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x32(tmem_o_own + (uint32_t)c0, v);
        }
        tmem_st_wait();
        l_run *= alpha;
        if (need) m_ref = m_new;
      }
      // P·V_{j-1} must have finished READING the P tile before block j overwrites it (this also observes every
      // phase of o_full in order — an mbarrier parity wait is only sound one phase behind)
      // (PT: P_j goes to the other S buffer, so the wait moves behind the P store — see there)
      if (!PT && j > 0 && !o_seen) mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
      // ---- P = exp2(S*scale - m), row sum, fp16 P into the swizzled A-operand tile
      unsigned long long sc2, mb2, l2 = 0ull;
      asm("mov.b64 %0, {%1, %1};" : "=l"(sc2) : "f"(p.scale_log2));
      asm("mov.b64 %0, {%1, %1};" : "=l"(mb2) : "f"(-mb));
      uint32_t pt[PT ? CW / 2 : 1];      // PT: the row's fp16 P values, two per register
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 32) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          unsigned long long x2, e2;
          asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "r"(sv[c0 + i]), "r"(sv[c0 + i + 1]));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x2) : "l"(x2), "l"(sc2), "l"(mb2));
          float x0, x1;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x2));
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);   // exp2(-inf) = 0 for masked keys
          asm("mov.b64 %0, {%1, %2};" : "=l"(e2) : "f"(p0), "f"(p1));
          asm("add.f32x2 %0, %0, %1;" : "+l"(l2) : "l"(e2));
          packed[i >> 1] = f2_to_h2(p0, p1);
        }
        if constexpr (PT) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pt[c0 / 2 + i] = packed[i];
        } else {
          const int cc = half * CW + c0;     // column inside the kBKV-wide P tile
          uint8_t* tile = sp + (cc / 64) * (kBQ * 128) + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((cc % 64) / 8 + q) ^ (r & 7);
            *reinterpret_cast<uint4*>(tile + chunk * 16) =
                make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
          }
        }
      }
      if constexpr (PT) {
        static_assert(!PT || CW == 64, "P in tensor memory: 64-key blocks (one 32-column store per row)");
        tmem_st_32x32b_x32(tmem_s + (uint32_t)(sb * kBKV) + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&pt[0]));
        tmem_st_wait();
        // P_j is PUBLISHED only after P·V_{j-1} has completed, i.e. after the issuer has consumed p_full phase j-1.
        // Without this the softmax warps — S_{j+1} is usually waiting for them — can finish block j+1 as well and
        // complete p_full TWICE before a delayed issuer (a late V tile) has polled it once; its parity wait then sees
        // the barrier two phases on, takes it for "not yet", and the CTA deadlocks.  P·V_{j-1} is long done by the
        // time a block's exponentials are, so the wait costs nothing; it also keeps every thread's view of o_full
        // one phase behind at most (rescale above, final read below).
        if (j > 0 && !o_seen) mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
      }
      {
        float la, lb;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(la), "=f"(lb) : "l"(l2));
        l_run += la + lb;
      }
      // S_j fully consumed -> the issuer may overwrite this buffer with S_{j+NSB}; P_j (and a rescaled O) ready
      tc_fence_before();
      if (!PT) fence_proxy_async_smem();  // P (generic-proxy stores) visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(p_full);
      }
    }
    if (KS == 2) {
      // row sum = sum of the halves' partial sums (both against the same reference maximum)
      float* ls = xch + 2 * (2 * kBQ);
      ls[half * kBQ + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      l_run += ls[(half ^ 1) * kBQ + r];
    }
    if (n_blocks > 0) {
      mbar_wait(o_full, (uint32_t)(n_blocks - 1) & 1u);
      tc_fence_after();
    }
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* op = p.out + (long long)b * p.o_stride_b + (long long)qpos * p.o_stride_t +
                 (long long)h * p.o_stride_h + half * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t v[32];
      if (n_blocks > 0) {
        tmem_ld_32x32b_x32(tmem_o_own + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (qpos < p.q_len) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 o;
          o.x = f2_to_h2(__uint_as_float(v[c]) * inv, __uint_as_float(v[c + 1]) * inv);
          o.y = f2_to_h2(__uint_as_float(v[c + 2]) * inv, __uint_as_float(v[c + 3]) * inv);
          o.z = f2_to_h2(__uint_as_float(v[c + 4]) * inv, __uint_as_float(v[c + 5]) * inv);
          o.w = f2_to_h2(__uint_as_float(v[c + 6]) * inv, __uint_as_float(v[c + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c0 + c) = o;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn attn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// (d, token, head, batch) fp16 view; box (64, box_rows, 1, 1); OOB rows read as zero.
int make_map4(CUtensorMap* out, const void* ptr, int d, long long len, int heads, int batch,
              long long stride_t, long long stride_h, long long stride_b, int box_rows) {
  EncodeTiledFn fn = attn_encode_fn();
  SCB_REQUIRE(fn != nullptr, SCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)len, (cuuint64_t)heads, (cuuint64_t)batch};
  // a size-1 dimension may carry any legal stride; keep them 16-byte multiples
  cuuint64_t strides[3] = {(cuuint64_t)stride_t * 2, (cuuint64_t)(heads > 1 ? stride_h : d) * 2,
                           (cuuint64_t)(batch > 1 ? stride_b : d) * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_REQUIRE(r == CUDA_SUCCESS, SCB_ERR_CUDA,
              "cuTensorMapEncodeTiled(4d) failed (%d): d=%d len=%lld heads=%d batch=%d strides=%lld/%lld/%lld",
              (int)r, d, len, heads, batch, stride_t, stride_h, stride_b);
  return SCB_OK;
}

template <int D, int kBKV, int NSB, int KS, bool PT = false>
int launch_attn(const scb_attention_args* a, cudaStream_t st) {
  using C = ACfg<D, kBKV, NSB, KS, PT>;
  CUtensorMap mq, mk, mv;
  int rc = make_map4(&mq, a->q, D, a->q_len, a->n_q_heads, a->batch, a->q_stride_t, a->q_stride_h,
                     a->q_stride_b, kBQ);
  if (rc != SCB_OK) return rc;
  rc = make_map4(&mk, a->k, D, a->kv_len, a->n_kv_heads, a->batch, a->k_stride_t, a->k_stride_h,
                 a->k_stride_b, kBKV);
  if (rc != SCB_OK) return rc;
  rc = make_map4(&mv, a->v, D, a->kv_len, a->n_kv_heads, a->batch, a->v_stride_t, a->v_stride_h,
                 a->v_stride_b, kBKV);
  if (rc != SCB_OK) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__half*>(a->out);
  p.o_stride_b = a->o_stride_b; p.o_stride_t = a->o_stride_t; p.o_stride_h = a->o_stride_h;
  p.q_len = (int)a->q_len; p.kv_len = (int)a->kv_len;
  p.n_q_heads = a->n_q_heads; p.n_kv_heads = a->n_kv_heads;
  p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  auto kern = attention_tcgen05_kernel<D, kBKV, NSB, KS, PT>;
  SCB_SET_FUNC_ATTR_ONCE_PER_DEVICE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  dim3 grid((unsigned)((a->q_len + kBQ - 1) / kBQ), (unsigned)a->n_q_heads, (unsigned)a->batch);
  kern<<<grid, C::kThreads, C::kSmemBytes, st>>>(mq, mk, mv, p);
  SCB_LAUNCH_OK("attention_tcgen05_kernel");
  return SCB_OK;
}

}  // namespace

extern "C" int scb_attention(const scb_attention_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_attention: args is NULL");
  SCB_REQUIRE(a->q && a->k && a->v && a->out, SCB_ERR_INVALID_ARG, "scb_attention: NULL buffer");
  SCB_REQUIRE(a->head_dim == 64 || a->head_dim == 128, SCB_ERR_UNSUPPORTED,
              "scb_attention: head_dim=%d (64 or 128 only)", a->head_dim);
  SCB_REQUIRE(a->batch >= 1 && a->batch <= 65535 && a->n_q_heads >= 1 && a->n_q_heads <= 65535 &&
                  a->n_kv_heads >= 1 && a->n_q_heads % a->n_kv_heads == 0,
              SCB_ERR_INVALID_ARG, "scb_attention: bad batch/head counts");
  SCB_REQUIRE(a->q_len >= 1 && a->kv_len >= 1 && a->q_len < (1ll << 30) && a->kv_len < (1ll << 30),
              SCB_ERR_INVALID_ARG, "scb_attention: bad sequence lengths");
  SCB_REQUIRE(!a->causal || a->kv_len >= a->q_len, SCB_ERR_INVALID_ARG,
              "scb_attention: causal needs kv_len >= q_len");
  const long long strides[] = {a->q_stride_t, a->q_stride_h, a->q_stride_b, a->k_stride_t,
                               a->k_stride_h, a->k_stride_b, a->v_stride_t, a->v_stride_h,
                               a->v_stride_b, a->o_stride_t, a->o_stride_h, a->o_stride_b};
  for (long long s : strides)
    SCB_REQUIRE(s % 8 == 0, SCB_ERR_INVALID_ARG,
                "scb_attention: strides must be multiples of 8 elements (got %lld)", s);
  SCB_REQUIRE(((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->out) % 16 == 0,
              SCB_ERR_INVALID_ARG, "scb_attention: buffers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // variants: 0 = 128-key blocks, one thread per row; 1 = 128-key blocks, two threads per row (8 softmax
  // warps); 2 = 64-key blocks, one thread per row, TWO CTAs per SM (256 TMEM columns, 112 KB each): the two
  // CTAs drift apart, so one reads S from TMEM (64 B/clk/SM) while the other one is in its exp2 phase;
  // 3 (default) = 2 with P kept in tensor memory (see ACfg): 285 -> 265 us for a 5 824-token causal prefill head
  // set, 217 -> 200 us for 64 CLIP frames.  (One S buffer and three CTAs per SM for D = 64 spills: 911 us.)
  static const int v64 = [] { const char* e = getenv("SCB_ATTN64"); return e ? atoi(e) : 3; }();
  static const int v128 = [] { const char* e = getenv("SCB_ATTN128"); return e ? atoi(e) : 3; }();
  if (a->head_dim == 64) {
    if (v64 == 3) return launch_attn<64, 64, 2, 1, true>(a, st);     // P in tensor memory
    if (v64 == 2) return launch_attn<64, 64, 2, 1>(a, st);
    return v64 == 1 ? launch_attn<64, 128, 2, 2>(a, st) : launch_attn<64, 128, 2, 1>(a, st);
  }
  if (v128 == 3) return launch_attn<128, 64, 2, 1, true>(a, st);     // P in tensor memory
  if (v128 == 2) return launch_attn<128, 64, 2, 1>(a, st);
  return v128 == 1 ? launch_attn<128, 128, 2, 2>(a, st) : launch_attn<128, 128, 2, 1>(a, st);
}
