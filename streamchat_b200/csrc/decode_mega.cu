// decode_mega.cu — one persistent kernel per cached decode step (scb_qwen2_decode_step).
//
// Replaces the per-token body of Qwen2ForCausalLM.generate reached from
// longva/model/language_model/llava_qwen.py:155 (transformers modeling_qwen2.py:269-311 per layer,
// :474-475 lm_head, greedy arg-max) for 1..4 sequences decoding in lock-step.  The step is
// HBM-bound: every weight byte and the live KV cache are read once (14.4 GB at LongVA-7B shapes),
// so the design goal is an HBM stream that never stops:
//
//   * one CTA per SM stays resident for the whole step (cooperative launch, 288 threads);
//   * warp 8 is the producer: it walks the step's weight matrices in execution order and streams
//     this CTA's share as 16-row x 512-k tiles (16 KB) into a shared-memory ring with TMA
//     (3-D tensor map (64 k, row, k-block), 128-byte swizzle).  Weights depend on nothing, so the
//     producer runs ahead of the layer dependencies and HBM keeps streaming while the consumers
//     sit in a grid barrier or in the attention stage;
//   * warps 0-7 are the consumers: warp w owns k-block w (64 k) of every tile and feeds it to
//     mma.sync.m16n8k16 (weights = A operand read with conflict-free 16-byte shared loads thanks to a
//     row permutation, activations = B operand from a shared-memory copy of the layer input); the
//     k-reduction across warps happens once per 16-row tile in a fixed order (deterministic);
//   * the stages of a layer (qkv | attention partials | attention combine | o_proj | gate|up | down)
//     are separated by grid barriers (one atomic + one flag, consumers only).  Skinny matrices
//     (o_proj, down_proj: 224 row tiles for 148 SMs) are cut along K so that every SM streams the
//     same number of bytes; the two partial sums of a tile are added in a fixed order by the next
//     stage's prologue, which also applies residual + RMSNorm (HF rounding points reproduced).
//
// Algorithmic bytes per step = sum of weight bytes + 2 * n_layers * n_kv * ctx * head_dim * 2 per row.
#include "scb_common.cuh"
#include <math.h>
#include <mutex>
#include <new>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kCWarps = 8;                 // consumer warps
constexpr int kCThreads = kCWarps * 32;    // 256
constexpr int kMegaThreads = kCThreads + 32;
constexpr int kUnitRows = 16;
constexpr int kUnitK = 512;
constexpr int kUnitBytes = kUnitRows * kUnitK * 2;   // 16 KB
constexpr int kRMax = 4;
constexpr int kMaxStages = 12;
constexpr int kMaxSplits = 64;
constexpr int kProfStamps = 16;

struct MegaLayer {          // device-resident, one per layer
  const __half* w[4];       // qkv, o, gate|up, down (for the L2 prefetch; the data path uses the tensor maps)
  const __half* qkv_b;
  const __half* ln1;
  const __half* ln2;
  __half* k_cache;
  __half* v_cache;
};

struct StageShape {         // one weight matrix
  int n_tiles;              // ceil(n / 16)
  int chunks;               // ceil(k / 512)
  int split;                // 1: units are (tile, chunk) pairs cut anywhere; 0: whole tiles per CTA
  int n, k;
  int g0;                   // > 0 (down_proj): CTAs [0, g0) stream k-chunks [0, nch0), CTAs [g0, G) the rest — a CTA then
                            // needs only ITS K-half of the activation rows in shared memory (see unit_plan)
};

struct MegaParams {
  const CUtensorMap* maps;  // [4 * n_layers + 1]: qkv, o, gu, down per layer, then lm_head
  const MegaLayer* layers;
  StageShape shp[5];        // qkv, o, gu, down, lm_head
  int n_layers, H, n_q, n_kv, I, V, qkv_dim;
  float eps, theta, scale_log2;
  const float* inv_freq;
  const __half* embed;
  const __half* final_norm;
  const __half* lm_head;
  long long max_ctx, cache_seq_stride;
  // per step
  int R, per_max, n_splits, n_stage, act_ld;
  long long* next_id;
  const int* kv_len;
  int* counters;
  int n_counters;
  __half* logits;
  long long ld_logits;
  // workspace (handle-owned)
  __half *x_buf, *x2_buf, *qkv_buf, *attn_buf, *gu_buf;
  float *part_o, *part_d;               // [slot][kRMax][H] fp32, zero between uses: o_proj 1 slot, down_proj 1 per K-half
  float *att_m, *att_l, *att_o;         // [kRMax][n_q][n_splits] (+ [HD])
  float* cand_val;                      // [grid][kRMax]
  int* cand_idx;
  unsigned* bar_flag;
  unsigned long long* bar_count;
  long long* prof;                      // optional [grid][n_layers][kProfStamps] clock64 stamps (debug)
};

SCB_DEVICE void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// fp16 value written by another CTA before the last grid barrier: read through L2
SCB_DEVICE float ldcg_h(const __half* p) { return __half2float(__ldcg(p)); }

// barrier among the 256 consumer threads only (the producer warp never joins)
SCB_DEVICE void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

SCB_DEVICE unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
SCB_DEVICE void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier for the consumer threads of every CTA.  count is monotonic (never reset): barrier number g
// completes when count == g * gridDim.x; flag publishes the last completed g.  The grid size is fixed per
// handle, so count == flag * gridDim.x holds between launches.  Release / acquire ride on the atomic and the
// polling load (bar.sync makes them cumulative over the CTA's threads); data written before a barrier by
// other CTAs is always read with ld.global.cg, so no L1 invalidation (gpu-scope fence) is needed.
// (A per-CTA flag array polled by every CTA was measured 2.5x slower: 148 x 148 acquire loads per round.)
SCB_DEVICE void grid_sync(const MegaParams& p, unsigned& gen) {
  cbar();
  if (threadIdx.x == 0) {
    gen += 1;
    unsigned long long old;
    asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(old) : "l"(p.bar_count) : "memory");
    if (old + 1ull == (unsigned long long)gen * gridDim.x) {
      st_release_u32(p.bar_flag, gen);
    } else {
      while ((int)(ld_acquire_u32(p.bar_flag) - gen) < 0) {
      }
    }
  }
  cbar();
}

struct Ring {
  uint8_t* buf;
  uint64_t* full;
  uint64_t* empty;
  int n_stage;
  int flat;        // units of this CTA's flattened sequence consumed by earlier stages (slot = index % n_stage)
};

// This CTA's units of a stage.  A unit = (16-row tile, 512-k chunk); units are numbered tile-major over the
// CTA's chunk range [ch0, ch0 + nch).  Three layouts:
//   whole tiles        (split 0): CTA c owns tiles [c*T/G, (c+1)*T/G) with all their chunks;
//   cut anywhere       (split 1): the T*chunks units are dealt out evenly; a tile spans at most two CTAs, whose
//                                 partial sums meet in ONE zero-initialised fp32 slot through red.add — 0 + a + b is
//                                 the same value in either order, so the result stays deterministic, and the next
//                                 prologue reads one slot per K-half instead of two (it is bound by those bytes:
//                                 every CTA re-reads every row);
//   K-halves           (g0 > 0):  CTAs [0, g0) work on chunks [0, nch0), the others on [nch0, chunks), each half cut
//                                 anywhere as above (partial slot = half).  The wide down_proj input
//                                 (18 944 values per sequence) then occupies HALF the shared memory per sequence, which
//                                 gives the weight ring back the stages 3- and 4-row steps were losing (6 -> 10, 4 -> 8).
struct UnitPlan {
  int u0, u1;      // unit range in the CTA's local numbering (tile = u / nch, local chunk = u % nch)
  int nch, ch0;    // chunk range of this CTA
  int pslot;       // partial slot of this CTA's K-half
};
SCB_DEVICE UnitPlan unit_plan(const StageShape& s) {
  const long long G = gridDim.x, c = blockIdx.x;
  UnitPlan pl;
  pl.nch = s.chunks;
  pl.ch0 = 0;
  pl.pslot = 0;
  if (s.g0 > 0) {
    const int nch0 = (s.chunks + 1) / 2;
    const bool hi = c >= s.g0;
    const long long cc = hi ? c - s.g0 : c, Gh = hi ? G - s.g0 : s.g0;
    pl.nch = hi ? s.chunks - nch0 : nch0;
    pl.ch0 = hi ? nch0 : 0;
    pl.pslot = hi ? 1 : 0;
    const long long U = (long long)s.n_tiles * pl.nch;
    pl.u0 = (int)(cc * U / Gh);
    pl.u1 = (int)((cc + 1) * U / Gh);
  } else if (s.split) {
    const long long U = (long long)s.n_tiles * s.chunks;
    pl.u0 = (int)(c * U / G);
    pl.u1 = (int)((c + 1) * U / G);
  } else {
    pl.u0 = (int)((c * s.n_tiles / G) * s.chunks);
    pl.u1 = (int)(((c + 1) * s.n_tiles / G) * s.chunks);
  }
  return pl;
}

SCB_DEVICE float block_sum_256(float v, float* red8) {   // consumers only; red8: 8 floats of scratch
  v = warp_sum(v);
  cbar();
  if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
  cbar();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kCWarps; ++w) t += red8[w];
  return t;
}


enum { ST_QKV = 0, ST_O = 1, ST_GU = 2, ST_DOWN = 3, ST_LM = 4 };

// ------------------------------------------------------------------------------------------
// One weight stage for the consumers: out-tile sums of act (shared, [R][act_ld] fp16) against this
// CTA's units, flushed per 16-row tile.
// ------------------------------------------------------------------------------------------
struct LmBest {
  float v;
  int i;
};

SCB_DEVICE uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// A warp owns WHOLE units (unit f of the CTA's flattened sequence belongs to warp f mod 8): the
// per-unit overhead (barrier wait, slot release, index math) is paid once per 16 KB instead of once per
// 2 KB, which keeps the consumers off the issue-bound regime (the first version spent ~500 issue cycles per
// unit and capped the stream at ~24 B/clk/SM).  The warps' partial tile sums meet in shared memory once
// per tile, in warp order.
template <int KIND>
SCB_DEVICE void run_stage(const MegaParams& p, const StageShape& s, const MegaLayer* ly, Ring& ring,
                          const __half* act, float* red, int& red_buf, LmBest& best) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);   // weight row (mod 8) fed to mma row g: keeps the 16-byte
                                               // reads of a quarter-warp on 8 distinct swizzled chunks
  const UnitPlan pl = unit_plan(s);
  const int u0 = pl.u0, u1 = pl.u1;
  const int R = p.R;
  const uint32_t ring_s = smem_u32(ring.buf);
  const uint32_t offa0 = (uint32_t)(rho * 128 + ((t ^ rho) << 4));          // k-step parity 0
  const uint32_t offa1 = (uint32_t)(rho * 128 + (((4 + t) ^ rho) << 4));    // k-step parity 1
  const uint32_t act_s = smem_u32(act) + (uint32_t)((g * p.act_ld + t * 8) * 2);
  const bool b_live = g < R;
  const int f0 = ring.flat;   // flattened index of unit u0
  const int chunks = pl.nch;            // chunks per tile in THIS CTA's numbering
  const int kglob0 = pl.ch0 * kUnitK;   // first k of the CTA's chunk range: act holds columns [kglob0, ...) of the input
  int tile = u0 / chunks;
  for (int ua = u0; ua < u1; ++tile) {
    const int ub = min(u1, (tile + 1) * chunks);
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
    // A unit belongs to the warp given by its RING SLOT (slot & 7), so consecutive uses of a slot have the
    // same owner: a warp never waits on a slot whose previous phase it has not consumed itself (an mbarrier
    // parity wait is only sound one phase ahead).
    int slot = (f0 + ua - u0) % ring.n_stage;
    uint32_t parity = (uint32_t)((f0 + ua - u0) / ring.n_stage) & 1u;
    for (int u = ua; u < ub; ++u, ++slot) {
      if (slot == ring.n_stage) {
        slot = 0;
        parity ^= 1u;
      }
      if ((slot & 7) != warp) continue;
      const int kbase = (u - tile * chunks) * kUnitK;
      mbar_wait(&ring.full[slot], parity);
      const uint32_t base = ring_s + (uint32_t)slot * kUnitBytes;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t off = (uint32_t)((j >> 1) * 2048) + ((j & 1) ? offa1 : offa0);
        const uint4 a0 = lds128(base + off);
        const uint4 a1 = lds128(base + off + 1024);
        uint4 b = make_uint4(0u, 0u, 0u, 0u);
        const int k = kbase + 32 * j;
        if (b_live && kglob0 + k + t * 8 < s.k) b = lds128(act_s + (uint32_t)(k * 2));
        mma16816(d0, a0.x, a1.x, a0.y, a1.y, b.x, b.y);
        mma16816(d1, a0.z, a1.z, a0.w, a1.w, b.z, b.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ring.empty[slot]);
    }
    {
      // ---- flush the tile: fixed-order sum over the 8 warps' partial sums
      float* r = red + red_buf * (kCWarps * kUnitRows * kRMax);
      if (t < 2) {
        float* q = r + (warp * kUnitRows + rho) * kRMax + 2 * t;
        q[0] = d0[0] + d1[0];
        q[1] = d0[1] + d1[1];
        q[8 * kRMax] = d0[2] + d1[2];
        q[8 * kRMax + 1] = d0[3] + d1[3];
      }
      cbar();
      red_buf ^= 1;
      if (KIND == ST_GU) {
        // rows (2i, 2i+1) = (gate_i, up_i): out = fp16(silu(fp16 gate)) * fp16 up  (modeling_qwen2.py:41-48)
        if (tid < 8 * kRMax) {
          const int i = tid >> 2, n = tid & 3;
          const int o = tile * 8 + i;
          if (n < R && 2 * o + 1 < s.n) {
            float sg = 0.f, su = 0.f;
#pragma unroll
            for (int w = 0; w < kCWarps; ++w) {
              sg += r[(w * kUnitRows + 2 * i) * kRMax + n];
              su += r[(w * kUnitRows + 2 * i + 1) * kRMax + n];
            }
            const float gt = __half2float(__float2half_rn(sg));
            const float up = __half2float(__float2half_rn(su));
            const float sl = __half2float(__float2half_rn(__fdividef(gt, 1.f + __expf(-gt))));
            p.gu_buf[(size_t)n * p.I + o] = __float2half_rn(sl * up);
          }
        }
      } else if (tid < kUnitRows * kRMax) {
        const int row = tid >> 2, n = tid & 3;
        const int o = tile * kUnitRows + row;
        if (n < R && o < s.n) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kCWarps; ++w) v += r[(w * kUnitRows + row) * kRMax + n];
          if (KIND == ST_QKV) {
            if (ly->qkv_b) v += __half2float(ly->qkv_b[o]);
            p.qkv_buf[(size_t)n * p.qkv_dim + o] = __float2half_rn(v);
          } else if (KIND == ST_O || KIND == ST_DOWN) {
            float* part = ((KIND == ST_O) ? p.part_o : p.part_d) + (size_t)pl.pslot * kRMax * p.H;
            atomicAdd(part + (size_t)n * p.H + o, v);   // at most two contributions per element onto a zeroed slot
          } else {  // ST_LM
            const __half hv = __float2half_rn(v);
            p.logits[(size_t)n * p.ld_logits + o] = hv;
            const float fv = __half2float(hv);   // the arg-max sees the stored (rounded) logits
            if (fv > best.v || (fv == best.v && o < best.i)) {
              best.v = fv;
              best.i = o;
            }
          }
        }
      }
    }
    ua = ub;
  }
  ring.flat += u1 - u0;
}

// ------------------------------------------------------------------------------------------
// Prologues: build the stage's input rows in shared memory (act [R][act_ld] fp16).
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Attention work plan: the S = n_splits CTAs of a kv head share the live sequences' KV positions.
//   R == 1: split j of the one sequence covers [j*L, (j+1)*L), L = ceil(len / S).
//   R  > 1: every CTA works on ONE (sequence, split) — the sequences are attended CONCURRENTLY, not one after the
//           other in every CTA (a short sequence's 37 tiny splits cost a full dependent-latency chain each:
//           5.4 us per sequence per layer).  With L = ceil(sum(len) / (S - R)) and n_s = ceil(len_s / L) splits for
//           sequence s, sum(n_s) < sum(len)/L + R <= S, so every split fits a CTA and no split is longer than L
//           (the host sizes the score buffer for L <= ceil(R * ctx_cap / (S - R))).
// Every CTA derives the same plan from kv_len[] (clamped as in attention_partial).
// ------------------------------------------------------------------------------------------
SCB_DEVICE int clamped_len(const MegaParams& p, int n) { return max(1, min(p.kv_len[n], (int)p.max_ctx)); }

SCB_DEVICE int plan_split_len(const MegaParams& p) {
  const int S = p.n_splits;
  if (p.R == 1 || S <= p.R) return (clamped_len(p, 0) + S - 1) / S;   // S <= R: legacy mapping (see attn_plan)
  int tot = 0;
  for (int n = 0; n < p.R; ++n) tot += clamped_len(p, n);
  return (tot + (S - p.R) - 1) / (S - p.R);
}

// number of splits of sequence n under the plan
SCB_DEVICE int plan_splits_of(const MegaParams& p, int n, int L) {
  if (p.R == 1 || p.n_splits <= p.R) return p.n_splits;
  return (clamped_len(p, n) + L - 1) / L;
}

// CTA `j` (0..S) of a kv head -> (sequence, local split, [begin, end)); seq = -1: idle.  Legacy mapping (R == 1, or
// fewer CTAs per head than sequences + 1): the caller loops over the sequences itself.
SCB_DEVICE void attn_plan(const MegaParams& p, int j, int L, int& seq, int& jl, int& begin, int& end) {
  seq = -1;
  jl = begin = end = 0;
  int first = 0;
  for (int n = 0; n < p.R; ++n) {
    const int len = clamped_len(p, n);
    const int ns = (len + L - 1) / L;
    if (j < first + ns) {
      seq = n;
      jl = j - first;
      begin = jl * L;
      end = min(len, begin + L);
      return;
    }
    first += ns;
  }
}

// Pull this CTA's part of the live K and V rows of one layer into L2 ahead of the attention stage (the addresses
// depend only on the context lengths).
SCB_DEVICE void prefetch_kv_l2(const MegaParams& p, const MegaLayer* ly, int hk, int split, int head_dim) {
  if (split >= p.n_splits) return;
  const int L = plan_split_len(p);
  const bool legacy = p.R == 1 || p.n_splits <= p.R;
  for (int n = 0; n < p.R; ++n) {
    int seq = n, begin, end, jl;
    if (legacy) {
      const int kv_len = clamped_len(p, n);
      const int per = (kv_len + p.n_splits - 1) / p.n_splits;
      begin = split * per;
      end = min(kv_len, begin + per);
    } else {
      if (n > 0) break;
      attn_plan(p, split, L, seq, jl, begin, end);
      if (seq < 0) break;
    }
    if (end <= begin) continue;
    const size_t off = (size_t)seq * p.cache_seq_stride + ((size_t)hk * p.max_ctx + begin) * head_dim;
    const long long bytes = (long long)(end - begin) * head_dim * 2;
    const char* kp = reinterpret_cast<const char*>(ly->k_cache + off);
    const char* vp = reinterpret_cast<const char*>(ly->v_cache + off);
    for (long long o = (long long)threadIdx.x * 128; o < bytes; o += (long long)kCThreads * 128) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + o));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + o));
    }
  }
}

// x = fp16(res + p0 + p1) (or the embedding rows), optional copy-out by CTA 0, then
// act = fp16(fp16(x * rstd) * gamma)  — RMSNorm in HF's order (modeling_qwen2.py:249-263).
// A thread owns 8-element chunks; every global load of a pass is issued before the first use
// (one L2 round trip per pass instead of one per element group).
SCB_DEVICE void prologue_norm(const MegaParams& p, __half* act, float* red8, const __half* res,
                              const float* part, int n_slots, const __half* embed_rows, __half* x_out,
                              const __half* gamma, const MegaLayer* kv_layer = nullptr, int hk = 0, int split = 0,
                              int head_dim = 0) {
  constexpr int CH = 2;                         // chunks per thread per pass
  const int tid = threadIdx.x;
  const int H = p.H, n_chunks = H >> 3;
  const size_t slot_stride = (size_t)kRMax * H;
  for (int n = 0; n < p.R; ++n) {
    __half* a = act + (size_t)n * p.act_ld;
    float ss = 0.f;
    const __half* er = embed_rows ? embed_rows + (size_t)p.next_id[n] * H : nullptr;
    for (int c0 = 0; c0 < n_chunks; c0 += kCThreads * CH) {
      uint4 rr[CH];
      float4 q0[CH][2], q1[CH][2];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = min(c0 + j * kCThreads + tid, n_chunks - 1);   // clamped: loads stay unconditional (registers)
        q0[j][0] = q0[j][1] = q1[j][0] = q1[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (er) {
          rr[j] = *reinterpret_cast<const uint4*>(er + c * 8);
        } else {
          rr[j] = __ldcg(reinterpret_cast<const uint4*>(res + (size_t)n * H + c * 8));
          const float4* s0 = reinterpret_cast<const float4*>(part + (size_t)n * H + c * 8);
          q0[j][0] = __ldcg(s0); q0[j][1] = __ldcg(s0 + 1);
          if (n_slots == 2) {      // K-halves (down_proj): slot 1 holds the upper half's partial sums
            const float4* s1 = reinterpret_cast<const float4*>(part + slot_stride + (size_t)n * H + c * 8);
            q1[j][0] = __ldcg(s1); q1[j][1] = __ldcg(s1 + 1);
          }
        }
      }
      // the attention stage's K / V rows start their trip to L2 behind this batch's loads
      if (kv_layer && n == 0 && c0 == 0) prefetch_kv_l2(p, kv_layer, hk, split, head_dim);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = c0 + j * kCThreads + tid;
        if (c < n_chunks) {
          float x[8];
          const float2 r0 = h2_to_f2(rr[j].x), r1 = h2_to_f2(rr[j].y), r2 = h2_to_f2(rr[j].z), r3 = h2_to_f2(rr[j].w);
          x[0] = r0.x; x[1] = r0.y; x[2] = r1.x; x[3] = r1.y; x[4] = r2.x; x[5] = r2.y; x[6] = r3.x; x[7] = r3.y;
          if (!er) {
            const float a0[8] = {q0[j][0].x, q0[j][0].y, q0[j][0].z, q0[j][0].w, q0[j][1].x, q0[j][1].y, q0[j][1].z, q0[j][1].w};
            const float a1[8] = {q1[j][0].x, q1[j][0].y, q1[j][0].z, q1[j][0].w, q1[j][1].x, q1[j][1].y, q1[j][1].z, q1[j][1].w};
            // fixed order: (lower K-half) + (upper K-half, absent without K-halves) + residual; each half is the
            // commutative sum of its (at most two) CTA contributions
#pragma unroll
            for (int e = 0; e < 8; ++e)
              x[e] = __half2float(__float2half_rn(n_slots == 2 ? (a0[e] + a1[e]) + x[e] : a0[e] + x[e]));
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) ss = fmaf(x[e], x[e], ss);
          uint4 o;
          o.x = f2_to_h2(x[0], x[1]); o.y = f2_to_h2(x[2], x[3]); o.z = f2_to_h2(x[4], x[5]); o.w = f2_to_h2(x[6], x[7]);
          *reinterpret_cast<uint4*>(a + c * 8) = o;
          if (x_out && blockIdx.x == 0) *reinterpret_cast<uint4*>(x_out + (size_t)n * H + c * 8) = o;
        }
      }
    }
    // gamma travels while the block reduces
    uint4 gm[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) gm[j] = *reinterpret_cast<const uint4*>(gamma + min(j * kCThreads + tid, n_chunks - 1) * 8);
    const float tot = block_sum_256(ss, red8);
    const float rstd = rsqrtf(tot / (float)H + p.eps);
    for (int c0 = 0; c0 < n_chunks; c0 += kCThreads * CH) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = c0 + j * kCThreads + tid;
        if (c < n_chunks) {
          const uint4 xv = *reinterpret_cast<const uint4*>(a + c * 8);   // own writes
          const uint4 g4 = (c0 == 0) ? gm[j] : *reinterpret_cast<const uint4*>(gamma + c * 8);
          const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {g4.x, g4.y, g4.z, g4.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xf = h2_to_f2(xs[e]), gf = h2_to_f2(gs[e]);
            o[e] = f2_to_h2(__half2float(__float2half_rn(xf.x * rstd)) * gf.x,
                            __half2float(__float2half_rn(xf.y * rstd)) * gf.y);
          }
          *reinterpret_cast<uint4*>(a + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  cbar();
}

// The partial slots go back to zero once EVERY CTA has read them (the caller sits behind the grid barrier that follows
// the reading prologue); each CTA clears its share of the R live rows.  The next red.add into them is several grid
// barriers away.
SCB_DEVICE void zero_partials(const MegaParams& p, float* part, int n_slots) {
  const size_t slot_stride = (size_t)kRMax * p.H;
  const int n = p.R * p.H;
  for (int sl = 0; sl < n_slots; ++sl)
    for (int i = blockIdx.x * kCThreads + threadIdx.x; i < n; i += gridDim.x * kCThreads) part[sl * slot_stride + i] = 0.f;
}

// act <- src rows ([R][len] fp16 in global, written by other CTAs before the last grid barrier);
// five independent 16-byte L2 loads in flight per thread.
SCB_DEVICE void prologue_copy(const MegaParams& p, __half* act, const __half* src, int row_len, int col0, int len) {
  // act[n][0 .. len) <- src[n][col0 .. col0 + len): the whole row (col0 = 0) or this CTA's K-half of it
  constexpr int U = 5;
  const int tid = threadIdx.x, n16 = len >> 3;
  for (int n = 0; n < p.R; ++n) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)n * row_len + col0);
    uint4* d4 = reinterpret_cast<uint4*>(act + (size_t)n * p.act_ld);
    for (int i0 = 0; i0 < n16; i0 += kCThreads * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcg(s4 + min(i0 + u * kCThreads + tid, n16 - 1));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * kCThreads + tid;
        if (i < n16) d4[i] = v[u];
      }
    }
  }
  cbar();
}

// ------------------------------------------------------------------------------------------
// Attention stage 1: this CTA = (kv head, split).  RoPE of the new q / k (same fp16 rounding points
// as rope_append), KV append by the split that owns the new position, split-KV attention of the
// group's q heads on mma.sync (as decode_attn_fused_kernel), partial (m, l, o) to the workspace.
// ------------------------------------------------------------------------------------------
template <int HD, int G>
SCB_DEVICE void attention_partial(const MegaParams& p, const MegaLayer* ly, uint8_t* scratch, int hk, int split,
                                  int seq, int begin, int end) {
  constexpr int HALF = HD / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_max = p.per_max, n_splits = p.n_splits;
  __half* q_h = reinterpret_cast<__half*>(scratch);                    // [8][HD] fp16
  float* cs_s = reinterpret_cast<float*>(scratch + 8 * HD * 2);          // [HALF]
  float* sn_s = cs_s + HALF;                                             // [HALF]
  float* m_s = sn_s + HALF;                                              // [8]
  float* l_s = m_s + 8;                                                  // [8]
  float* o_s = l_s + 8;                                                  // [8 warps][G][HD]
  float* s_s = o_s + 8 * G * HD;                                         // [G][per_max]
  const __half* qkv = p.qkv_buf + (size_t)seq * p.qkv_dim;
  __half* k_cache = ly->k_cache + (size_t)seq * p.cache_seq_stride;
  __half* v_cache = ly->v_cache + (size_t)seq * p.cache_seq_stride;
  const long long max_pos = p.max_ctx;
  // the host validates lengths (Qwen2Engine raises before a step could run past the cache); this guard keeps a
  // bad length from ever writing outside the slot: the length is clamped and the append skipped
  const int kv_raw = p.kv_len[seq];
  const bool kv_ok = kv_raw >= 1 && kv_raw <= max_pos;
  const int kv_len = max(1, min(kv_raw, (int)max_pos));
  const int pos = kv_len - 1;
  end = min(end, kv_len);
  const int n_local = min(per_max, max(0, end - begin));   // the plan bounds every range by per_max (see attn_plan)

  if (tid < HALF) {
    const float f = p.inv_freq ? p.inv_freq[tid] : powf(p.theta, -(float)(2 * tid) / (float)HD);
    const float ang = (float)pos * f;
    cs_s[tid] = __half2float(__float2half_rn(cosf(ang)));
    sn_s[tid] = __half2float(__float2half_rn(sinf(ang)));
  }
  cbar();
  for (int idx = tid; idx < G * HALF; idx += kCThreads) {
    const int h = idx / HALF, i = idx % HALF;
    const float c = cs_s[i], sn = sn_s[i];
    const __half* qp = qkv + (size_t)(hk * G + h) * HD;
    const float x1 = ldcg_h(qp + i), x2 = ldcg_h(qp + i + HALF);
    const float a1 = __half2float(__float2half_rn(x1 * c)), b1 = __half2float(__float2half_rn(-x2 * sn));
    const float a2 = __half2float(__float2half_rn(x2 * c)), b2 = __half2float(__float2half_rn(x1 * sn));
    q_h[h * HD + i] = __float2half_rn(a1 + b1);
    q_h[h * HD + i + HALF] = __float2half_rn(a2 + b2);
  }
  for (int idx = G * HD + tid; idx < 8 * HD; idx += kCThreads) q_h[idx] = __float2half_rn(0.f);
  if (kv_ok && pos >= begin && pos < end) {
    const __half* kp = qkv + (size_t)(p.n_q + hk) * HD;
    const __half* vp = qkv + (size_t)(p.n_q + p.n_kv + hk) * HD;
    __half* kc = k_cache + ((size_t)hk * max_pos + pos) * HD;
    __half* vc = v_cache + ((size_t)hk * max_pos + pos) * HD;
    for (int i = tid; i < HALF; i += kCThreads) {
      const float c = cs_s[i], sn = sn_s[i];
      const float x1 = ldcg_h(kp + i), x2 = ldcg_h(kp + i + HALF);
      const float a1 = __half2float(__float2half_rn(x1 * c)), b1 = __half2float(__float2half_rn(-x2 * sn));
      const float a2 = __half2float(__float2half_rn(x2 * c)), b2 = __half2float(__float2half_rn(x1 * sn));
      kc[i] = __float2half_rn(a1 + b1);
      kc[i + HALF] = __float2half_rn(a2 + b2);
    }
    for (int i = tid; i < HD; i += kCThreads)
      vc[i] = __ldcg(vp + i);
  }
  cbar();   // q_h ready; the appended row is visible to this CTA (block-scope ordering)

  // ---- scores: K rows are the A operand (M = positions, K = dims, N = the group's q heads)
  {
    constexpr int NC = HD / 32;
    const int g = lane >> 2, t = lane & 3;
    uint4 qf[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) qf[c] = *reinterpret_cast<const uint4*>(q_h + g * HD + c * 32 + t * 8);
    const __half* kb = k_cache + ((size_t)hk * max_pos + begin) * HD + t * 8;
    // two 128-position sweeps per iteration: all 16-byte loads of both are in flight before the first mma
    for (int p0 = warp * 16; p0 < n_local; p0 += 256) {
      uint4 k0[2][NC], k1[2][NC];
      bool ok0[2], ok1[2];
#pragma unroll
      for (int sw = 0; sw < 2; ++sw) {
        const int r0 = p0 + sw * 128 + g, r1 = r0 + 8;
        ok0[sw] = r0 < n_local;
        ok1[sw] = r1 < n_local;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          k0[sw][c] = ok0[sw] ? *reinterpret_cast<const uint4*>(kb + (size_t)r0 * HD + c * 32) : make_uint4(0, 0, 0, 0);
          k1[sw][c] = ok1[sw] ? *reinterpret_cast<const uint4*>(kb + (size_t)r1 * HD + c * 32) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int sw = 0; sw < 2; ++sw) {
        const int r0 = p0 + sw * 128 + g, r1 = r0 + 8;
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          mma16816(d0, k0[sw][c].x, k1[sw][c].x, k0[sw][c].y, k1[sw][c].y, qf[c].x, qf[c].y);
          mma16816(d1, k0[sw][c].z, k1[sw][c].z, k0[sw][c].w, k1[sw][c].w, qf[c].z, qf[c].w);
        }
        const int h0 = 2 * t;
        if (h0 < G) {
          if (ok0[sw]) s_s[h0 * per_max + r0] = (d0[0] + d1[0]) * p.scale_log2;
          if (ok1[sw]) s_s[h0 * per_max + r1] = (d0[2] + d1[2]) * p.scale_log2;
        }
        if (h0 + 1 < G) {
          if (ok0[sw]) s_s[(h0 + 1) * per_max + r0] = (d0[1] + d1[1]) * p.scale_log2;
          if (ok1[sw]) s_s[(h0 + 1) * per_max + r1] = (d0[3] + d1[3]) * p.scale_log2;
        }
      }
    }
  }
  cbar();
  // ---- per-head softmax statistics (warp h)
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
  cbar();
  // ---- O^T = V^T · P^T: M = dims, K = positions, N = q heads (probabilities rounded to fp16)
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
    const __half* vb = v_cache + ((size_t)hk * max_pos + begin) * HD + g * 8;
    const float* er = s_s + (g < G ? g : 0) * per_max;
    for (int p0 = warp * 16; p0 < n_local; p0 += 256) {
      uint4 v[2][NH][4];
#pragma unroll
      for (int sw = 0; sw < 2; ++sw) {
        const int pa = p0 + sw * 128 + 2 * t, pb = pa + 8;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const __half* vp = vb + hh * 64;
          v[sw][hh][0] = pa < n_local ? *reinterpret_cast<const uint4*>(vp + (size_t)pa * HD) : make_uint4(0, 0, 0, 0);
          v[sw][hh][1] = pa + 1 < n_local ? *reinterpret_cast<const uint4*>(vp + (size_t)(pa + 1) * HD) : make_uint4(0, 0, 0, 0);
          v[sw][hh][2] = pb < n_local ? *reinterpret_cast<const uint4*>(vp + (size_t)pb * HD) : make_uint4(0, 0, 0, 0);
          v[sw][hh][3] = pb + 1 < n_local ? *reinterpret_cast<const uint4*>(vp + (size_t)(pb + 1) * HD) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int sw = 0; sw < 2; ++sw) {
        const int pa = p0 + sw * 128 + 2 * t, pb = pa + 8;
        const bool oka0 = pa < n_local, oka1 = pa + 1 < n_local, okb0 = pb < n_local, okb1 = pb + 1 < n_local;
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
          const uint32_t w0[4] = {v[sw][hh][0].x, v[sw][hh][0].y, v[sw][hh][0].z, v[sw][hh][0].w};
          const uint32_t w1[4] = {v[sw][hh][1].x, v[sw][hh][1].y, v[sw][hh][1].z, v[sw][hh][1].w};
          const uint32_t w2[4] = {v[sw][hh][2].x, v[sw][hh][2].y, v[sw][hh][2].z, v[sw][hh][2].w};
          const uint32_t w3[4] = {v[sw][hh][3].x, v[sw][hh][3].y, v[sw][hh][3].z, v[sw][hh][3].w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma16816(acc[hh][j], __byte_perm(w0[j], w1[j], 0x5410), __byte_perm(w0[j], w1[j], 0x7632),
                     __byte_perm(w2[j], w3[j], 0x5410), __byte_perm(w2[j], w3[j], 0x7632), b0, b1);
        }
      }
    }
    const int h0 = 2 * t;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d0 = hh * 64 + g * 8 + 2 * j;
        if (h0 < G) {
          o_s[(warp * G + h0) * HD + d0] = acc[hh][j][0];
          o_s[(warp * G + h0) * HD + d0 + 1] = acc[hh][j][2];
        }
        if (h0 + 1 < G) {
          o_s[(warp * G + h0 + 1) * HD + d0] = acc[hh][j][1];
          o_s[(warp * G + h0 + 1) * HD + d0 + 1] = acc[hh][j][3];
        }
      }
  }
  cbar();
  for (int idx = tid; idx < G * HD; idx += kCThreads) {
    const int h = idx / HD, d = idx % HD;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += o_s[(w * G + h) * HD + d];
    const size_t pi = ((size_t)seq * p.n_q + hk * G + h) * n_splits + split;
    p.att_o[pi * HD + d] = v;
    if (d == 0) {
      p.att_m[pi] = m_s[h];
      p.att_l[pi] = l_s[h];
    }
  }
  cbar();   // scratch reuse by the next sequence
}

// Attention stage 2: task = (sequence, q head, 32-dim slice): merge the splits' partials in split order.
// 256 threads = 32 dims x 8 split groups; a thread's <= 8 partial loads are issued before the split
// weights exist (they do not depend on them), so a task costs one L2 round trip.
template <int HD>
SCB_DEVICE void attention_combine(const MegaParams& p, uint8_t* scratch, int task, int L) {
  constexpr int QS = HD / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* w_s = reinterpret_cast<float*>(scratch);    // [kMaxSplits]
  float* inv_l = w_s + kMaxSplits;                   // [1]
  float* acc_s = inv_l + 8;                          // [8][32]
  const int head_task = task / QS, d = (task - head_task * QS) * 32 + lane;   // head_task = seq * n_q + head
  const size_t base = (size_t)head_task * p.n_splits;                         // slot stride stays n_splits
  const int n_splits = plan_splits_of(p, head_task / p.n_q, L);               // live partials of this sequence
  const float* po = p.att_o + base * HD + d;
  float ov[kMaxSplits / kCWarps];
#pragma unroll
  for (int j = 0; j < kMaxSplits / kCWarps; ++j) {
    const int sp = warp + j * kCWarps;
    ov[j] = sp < n_splits ? __ldcg(po + (size_t)sp * HD) : 0.f;
  }
  if (warp == 0) {
    float mm = -INFINITY;
    float pm[kMaxSplits / 32], pl[kMaxSplits / 32];
#pragma unroll
    for (int j = 0; j < kMaxSplits / 32; ++j) {
      const int sp = lane + j * 32;
      pm[j] = sp < n_splits ? __ldcg(p.att_m + base + sp) : -INFINITY;
      pl[j] = sp < n_splits ? __ldcg(p.att_l + base + sp) : 0.f;
      mm = fmaxf(mm, pm[j]);
    }
    mm = warp_max(mm);
    float ll = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxSplits / 32; ++j) {
      const int sp = lane + j * 32;
      const float scl = (pm[j] == -INFINITY) ? 0.f : exp2f(pm[j] - mm);
      if (sp < n_splits) w_s[sp] = scl;
      ll = fmaf(pl[j], scl, ll);
    }
    ll = warp_sum(ll);
    if (lane == 0) inv_l[0] = ll > 0.f ? 1.f / ll : 0.f;
  }
  cbar();
  {
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxSplits / kCWarps; ++j) {
      const int sp = warp + j * kCWarps;
      if (sp < n_splits) o = fmaf(ov[j], w_s[sp], o);
    }
    acc_s[warp * 32 + lane] = o;
  }
  cbar();
  if (warp == 0) {
    float o = 0.f;
#pragma unroll
    for (int gq = 0; gq < kCWarps; ++gq) o += acc_s[gq * 32 + lane];
    p.attn_buf[(size_t)head_task * HD + d] = __float2half_rn(o * inv_l[0]);   // [seq][n_q][HD]
  }
  cbar();
}

// ------------------------------------------------------------------------------------------
template <int HD, int G>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ring_buf = smem;
  uint8_t* scratch = ring_buf + (size_t)p.n_stage * kUnitBytes;            // act rows / attention scratch
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages];
  __shared__ float red[2 * kCWarps * kUnitRows * kRMax];
  __shared__ float red8[kRMax * kCWarps];
  __shared__ float lm_v[kUnitRows * kRMax];
  __shared__ int lm_i[kUnitRows * kRMax];
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.n_stage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);      // released by the one warp that owns the unit
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kCWarps) {
    // ===================== producer: every weight tile of the step, in execution order =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const int n_mats = 4 * p.n_layers + 1;
      for (int m = 0; m < n_mats; ++m) {
        const StageShape& s = p.shp[m == n_mats - 1 ? ST_LM : (m & 3)];
        const CUtensorMap* map = p.maps + m;
        const UnitPlan pl = unit_plan(s);
        int tile = pl.u0 / pl.nch, kc = pl.u0 - tile * pl.nch;
        for (int u = pl.u0; u < pl.u1; ++u, ++kc) {
          if (kc == pl.nch) {
            kc = 0;
            ++tile;
          }
          mbar_wait(&empty[slot], phase ^ 1u);
          mbar_arrive_expect_tx(&full[slot], kUnitBytes);
          tma_load_3d(ring_buf + (size_t)slot * kUnitBytes, map, &full[slot], 0, tile * kUnitRows, (pl.ch0 + kc) * 8);
          if (++slot == p.n_stage) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  Ring ring{ring_buf, full, empty, p.n_stage, 0};
  __half* act = reinterpret_cast<__half*>(scratch);
  unsigned gen = 0;
  if (tid == 0) gen = ld_acquire_u32(p.bar_flag);   // only thread 0 uses it
  int red_buf = 0;
  LmBest best{-INFINITY, 0x7fffffff};
  const int cta = blockIdx.x;
  const int hk = cta % p.n_kv, split = cta / p.n_kv;

  long long* prof = (p.prof && tid == 0) ? p.prof + (size_t)cta * p.n_layers * kProfStamps : nullptr;
#define SCB_STAMP(i) do { if (prof) prof[l * kProfStamps + (i)] = clock64(); } while (0)
  const int down_slots = p.shp[ST_DOWN].g0 > 0 ? 2 : 1;   // partial slots the down_proj stage fills (one per K-half)
  for (int l = 0; l < p.n_layers; ++l) {
    const MegaLayer* ly = p.layers + l;
    SCB_STAMP(0);
    // ---- qkv: x = embedding rows (layer 0) or x2 + down partials; act = RMSNorm(x) * ln1
    prologue_norm(p, act, red8, p.x2_buf, p.part_d, down_slots, l == 0 ? p.embed : nullptr, p.x_buf, ly->ln1, ly, hk, split, HD);
    SCB_STAMP(1);
    run_stage<ST_QKV>(p, p.shp[ST_QKV], ly, ring, act, red, red_buf, best);
    SCB_STAMP(2);
    grid_sync(p, gen);
    SCB_STAMP(3);
    zero_partials(p, p.part_d, down_slots);   // every CTA has read the previous layer's down_proj sums
    // ---- attention partials
    const int plan_L = plan_split_len(p);
    if (split < p.n_splits) {
      if (p.R == 1 || p.n_splits <= p.R) {
        for (int n = 0; n < p.R; ++n) {
          const int len = clamped_len(p, n), per = (len + p.n_splits - 1) / p.n_splits;
          attention_partial<HD, G>(p, ly, scratch, hk, split, n, split * per, min(len, split * per + per));
        }
      } else {
        int seq, jl, begin, end;
        attn_plan(p, split, plan_L, seq, jl, begin, end);
        if (seq >= 0) attention_partial<HD, G>(p, ly, scratch, hk, jl, seq, begin, end);
      }
    }
    SCB_STAMP(4);
    grid_sync(p, gen);
    SCB_STAMP(5);
    for (int task = cta; task < p.R * p.n_q * (HD / 32); task += gridDim.x)
      attention_combine<HD>(p, scratch, task, plan_L);
    SCB_STAMP(6);
    grid_sync(p, gen);
    SCB_STAMP(7);
    // ---- o_proj
    prologue_copy(p, act, p.attn_buf, p.n_q * HD, 0, p.n_q * HD);
    SCB_STAMP(8);
    run_stage<ST_O>(p, p.shp[ST_O], ly, ring, act, red, red_buf, best);
    SCB_STAMP(9);
    grid_sync(p, gen);
    SCB_STAMP(10);
    // ---- gate|up: x2 = x + o partials; act = RMSNorm(x2) * ln2
    prologue_norm(p, act, red8, p.x_buf, p.part_o, 1, nullptr, p.x2_buf, ly->ln2);
    SCB_STAMP(11);
    run_stage<ST_GU>(p, p.shp[ST_GU], ly, ring, act, red, red_buf, best);
    SCB_STAMP(12);
    grid_sync(p, gen);
    SCB_STAMP(13);
    zero_partials(p, p.part_o, 1);            // every CTA has read the o_proj sums
    // ---- down
    {
      // down_proj input: the whole gated row, or (K-halves) only the columns of this CTA's chunk range
      const UnitPlan dp = unit_plan(p.shp[ST_DOWN]);
      const int col0 = dp.ch0 * kUnitK;
      prologue_copy(p, act, p.gu_buf, p.I, col0, min(p.I - col0, dp.nch * kUnitK));
    }
    SCB_STAMP(14);
    run_stage<ST_DOWN>(p, p.shp[ST_DOWN], ly, ring, act, red, red_buf, best);
    SCB_STAMP(15);
    grid_sync(p, gen);
  }
#undef SCB_STAMP
  // ---- final norm + lm_head + arg-max
  prologue_norm(p, act, red8, p.x2_buf, p.part_d, down_slots, nullptr, nullptr, p.final_norm);
  run_stage<ST_LM>(p, p.shp[ST_LM], nullptr, ring, act, red, red_buf, best);
  if (tid < kUnitRows * kRMax) {
    lm_v[tid] = best.v;
    lm_i[tid] = best.i;
  }
  cbar();
  if (tid < p.R) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int row = 0; row < kUnitRows; ++row) {
      const float v = lm_v[row * kRMax + tid];
      const int i = lm_i[row * kRMax + tid];
      if (v > bv || (v == bv && i < bi)) {
        bv = v;
        bi = i;
      }
    }
    p.cand_val[(size_t)cta * kRMax + tid] = bv;
    p.cand_idx[(size_t)cta * kRMax + tid] = bi;
  }
  grid_sync(p, gen);
  zero_partials(p, p.part_d, down_slots);     // read by the final norm above: the next step starts from zeroed slots
  if (cta == 0) {
    if (warp < p.R) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < (int)gridDim.x; c += 32) {
        const float v = __ldcg(p.cand_val + (size_t)c * kRMax + warp);
        const int i = __ldcg(p.cand_idx + (size_t)c * kRMax + warp);
        if (v > bv || (v == bv && i < bi)) {
          bv = v;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) p.next_id[warp] = (bi == 0x7fffffff) ? 0 : (long long)bi;
    }
    if (tid == 0)
      for (int cc = 0; cc < p.n_counters; ++cc) p.counters[cc] += 1;
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn mega_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// W [n, k] fp16, k contiguous -> (64 k, row, k-block) view; box (64, 16, 8) lands in shared memory as
// [8 k-blocks][16 rows][128 B] with the 128-byte swizzle keyed on the row; OOB rows / k-blocks read zero.
int make_weight_map(CUtensorMap* out, const void* w, long long n, long long k) {
  EncodeTiledFn fn = mega_encode_fn();
  SCB_REQUIRE(fn != nullptr, SCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {64, (cuuint64_t)n, (cuuint64_t)(k / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)k * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)kUnitRows, 8};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_REQUIRE(r == CUDA_SUCCESS, SCB_ERR_CUDA, "cuTensorMapEncodeTiled(weight %lld x %lld) failed (%d)", n, k,
              (int)r);
  return SCB_OK;
}

struct MegaHandle {
  uint32_t magic;
  MegaParams base;
  int grid;
  int head_dim, group;
  void* dev_maps;
  void* dev_layers;
  void* dev_ws;
  void* dev_prof;
  int smem_limit;
  MegaLayer* host_layers;   // host copy of the per-layer pointer table (scb_qwen2_prefill)
  int hidden, inter, vocab, n_q, n_kv, qkv_dim, max_seqs;
};
constexpr uint32_t kMagic = 0x51323244u;

typedef void (*MegaKernel)(const MegaParams);
MegaKernel pick_kernel(int hd, int g) {
#define SCB_MK(HD, GG) \
  if (hd == HD && g == GG) return decode_mega_kernel<HD, GG>;
  SCB_MK(128, 1) SCB_MK(128, 2) SCB_MK(128, 4) SCB_MK(128, 7) SCB_MK(128, 8)
  SCB_MK(64, 1) SCB_MK(64, 2) SCB_MK(64, 4) SCB_MK(64, 7) SCB_MK(64, 8)
#undef SCB_MK
  return nullptr;
}

StageShape make_shape(long long n, long long k, int grid, bool may_split) {
  StageShape s;
  s.n = (int)n;
  s.k = (int)k;
  s.n_tiles = (int)((n + kUnitRows - 1) / kUnitRows);
  s.chunks = (int)((k + kUnitK - 1) / kUnitK);
  const long long U = (long long)s.n_tiles * s.chunks;
  // cutting tiles along K is only sound while a tile spans at most two CTAs (two partial slots)
  s.split = (may_split && U / grid >= s.chunks) ? 1 : 0;
  s.g0 = 0;
  return s;
}

size_t attn_scratch_bytes(int hd, int g, int per_max) {
  return (size_t)8 * hd * 2 + sizeof(float) * ((size_t)hd + 16 + (size_t)8 * g * hd + (size_t)g * per_max);
}

}  // namespace

extern "C" {

int scb_qwen2_decoder_max_rows(void) { return kRMax; }

int scb_qwen2_decoder_create(const scb_qwen2_decoder_desc* d, void** out_handle) {
  SCB_REQUIRE(d != nullptr && out_handle != nullptr, SCB_ERR_INVALID_ARG, "scb_qwen2_decoder_create: NULL argument");
  *out_handle = nullptr;
  SCB_REQUIRE(d->n_layers >= 1 && d->hidden >= 64 && d->n_q_heads >= 1 && d->n_kv_heads >= 1 && d->inter >= 64 &&
                  d->vocab >= 16 && d->max_seqs >= 1 && d->max_ctx >= 1,
              SCB_ERR_INVALID_ARG, "scb_qwen2_decoder_create: bad dimensions");
  SCB_REQUIRE(d->n_q_heads % d->n_kv_heads == 0, SCB_ERR_INVALID_ARG, "scb_qwen2_decoder_create: n_q %% n_kv != 0");
  const int G = d->n_q_heads / d->n_kv_heads;
  MegaKernel kern = pick_kernel(d->head_dim, G);
  SCB_REQUIRE(kern != nullptr, SCB_ERR_UNSUPPORTED,
              "scb_qwen2_decoder_create: head_dim=%d group=%d unsupported (64|128 x 1,2,4,7,8)", d->head_dim, G);
  const long long qkv_dim = (long long)(d->n_q_heads + 2 * d->n_kv_heads) * d->head_dim;
  const long long attn_dim = (long long)d->n_q_heads * d->head_dim;
  SCB_REQUIRE(d->hidden % 64 == 0 && d->inter % 64 == 0 && attn_dim % 64 == 0, SCB_ERR_UNSUPPORTED,
              "scb_qwen2_decoder_create: hidden / inter / n_q*head_dim must be multiples of 64");
  SCB_REQUIRE(d->hidden % 16 == 0 && qkv_dim % 16 == 0 && d->vocab % 16 == 0 && d->inter % 8 == 0,
              SCB_ERR_UNSUPPORTED, "scb_qwen2_decoder_create: output widths must be multiples of 16");
  SCB_REQUIRE(d->embed && d->final_norm && d->lm_head && d->qkv_w && d->o_w && d->gu_w && d->down_w && d->ln1 &&
                  d->ln2 && d->k_cache && d->v_cache,
              SCB_ERR_INVALID_ARG, "scb_qwen2_decoder_create: NULL pointer table");

  int dev = 0;
  SCB_CUDA_OK(cudaGetDevice(&dev));
  int coop = 0, smem_optin = 0;
  SCB_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  SCB_REQUIRE(coop != 0, SCB_ERR_UNSUPPORTED, "scb_qwen2_decoder_create: device lacks cooperative launch");
  SCB_CUDA_OK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int grid = scb_num_sms();

  MegaHandle* h = new (std::nothrow) MegaHandle();
  SCB_REQUIRE(h != nullptr, SCB_ERR_CUDA, "scb_qwen2_decoder_create: out of host memory");
  h->magic = kMagic;
  h->grid = grid;
  h->head_dim = d->head_dim;
  h->group = G;
  h->smem_limit = smem_optin;
  h->dev_maps = h->dev_layers = h->dev_ws = h->dev_prof = nullptr;
  h->host_layers = nullptr;
  h->hidden = d->hidden; h->inter = d->inter; h->vocab = d->vocab; h->n_q = d->n_q_heads; h->n_kv = d->n_kv_heads;
  h->qkv_dim = (int)qkv_dim; h->max_seqs = d->max_seqs;
  MegaParams& p = h->base;
  memset(&p, 0, sizeof(p));
  p.n_layers = d->n_layers; p.H = d->hidden; p.n_q = d->n_q_heads; p.n_kv = d->n_kv_heads;
  p.I = d->inter; p.V = d->vocab; p.qkv_dim = (int)qkv_dim;
  p.eps = d->rms_eps; p.theta = d->rope_theta;
  p.scale_log2 = (1.f / sqrtf((float)d->head_dim)) * 1.4426950408889634f;
  p.inv_freq = d->inv_freq;
  p.embed = reinterpret_cast<const __half*>(d->embed);
  p.final_norm = reinterpret_cast<const __half*>(d->final_norm);
  p.lm_head = reinterpret_cast<const __half*>(d->lm_head);
  p.max_ctx = d->max_ctx;
  p.cache_seq_stride = (long long)d->n_kv_heads * d->max_ctx * d->head_dim;
  p.shp[ST_QKV] = make_shape(qkv_dim, d->hidden, grid, false);
  p.shp[ST_O] = make_shape(d->hidden, attn_dim, grid, true);
  p.shp[ST_GU] = make_shape(2ll * d->inter, d->hidden, grid, false);
  p.shp[ST_DOWN] = make_shape(d->hidden, d->inter, grid, true);
  {
    // K-halves for down_proj (see unit_plan): CTAs split in proportion to the halves' chunk counts; sound while every
    // half has at least as many tiles as CTAs (a tile then spans at most two CTAs of a half)
    StageShape& sd = p.shp[ST_DOWN];
    const int nch0 = (sd.chunks + 1) / 2, nch1 = sd.chunks - nch0;
    int g0 = (int)(((long long)grid * nch0 + sd.chunks / 2) / sd.chunks);
    if (g0 < 1) g0 = 1;
    if (g0 > grid - 1) g0 = grid - 1;
    static const int use_khalf = [] { const char* e = getenv("SCB_MEGA_KHALF"); return e ? atoi(e) : 1; }();
    if (use_khalf && sd.split && nch1 >= 2 && grid >= 2 && sd.n_tiles >= g0 && sd.n_tiles >= grid - g0) sd.g0 = g0;
  }
  p.shp[ST_LM] = make_shape(d->vocab, d->hidden, grid, false);
  int ns = grid / d->n_kv_heads;
  if (ns > kMaxSplits) ns = kMaxSplits;
  if (ns < 1) {
    delete h;
    SCB_REQUIRE(false, SCB_ERR_UNSUPPORTED, "scb_qwen2_decoder_create: more kv heads than SMs");
  }
  p.n_splits = ns;

  // ---- tensor maps + layer table
  const int n_mats = 4 * d->n_layers + 1;
  CUtensorMap* maps = new (std::nothrow) CUtensorMap[n_mats];
  MegaLayer* layers = new (std::nothrow) MegaLayer[d->n_layers];
  int rc = (maps && layers) ? SCB_OK : SCB_ERR_CUDA;
  for (int l = 0; rc == SCB_OK && l < d->n_layers; ++l) {
    if (!d->qkv_w[l] || !d->o_w[l] || !d->gu_w[l] || !d->down_w[l] || !d->ln1[l] || !d->ln2[l] || !d->k_cache[l] ||
        !d->v_cache[l]) {
      scb_set_error("scb_qwen2_decoder_create: NULL pointer for layer %d", l);
      rc = SCB_ERR_INVALID_ARG;
      break;
    }
    rc = make_weight_map(&maps[4 * l + 0], d->qkv_w[l], qkv_dim, d->hidden);
    if (rc == SCB_OK) rc = make_weight_map(&maps[4 * l + 1], d->o_w[l], d->hidden, attn_dim);
    if (rc == SCB_OK) rc = make_weight_map(&maps[4 * l + 2], d->gu_w[l], 2ll * d->inter, d->hidden);
    if (rc == SCB_OK) rc = make_weight_map(&maps[4 * l + 3], d->down_w[l], d->hidden, d->inter);
    layers[l].w[0] = reinterpret_cast<const __half*>(d->qkv_w[l]);
    layers[l].w[1] = reinterpret_cast<const __half*>(d->o_w[l]);
    layers[l].w[2] = reinterpret_cast<const __half*>(d->gu_w[l]);
    layers[l].w[3] = reinterpret_cast<const __half*>(d->down_w[l]);
    layers[l].qkv_b = reinterpret_cast<const __half*>(d->qkv_b ? d->qkv_b[l] : nullptr);
    layers[l].ln1 = reinterpret_cast<const __half*>(d->ln1[l]);
    layers[l].ln2 = reinterpret_cast<const __half*>(d->ln2[l]);
    layers[l].k_cache = reinterpret_cast<__half*>(d->k_cache[l]);
    layers[l].v_cache = reinterpret_cast<__half*>(d->v_cache[l]);
  }
  if (rc == SCB_OK) rc = make_weight_map(&maps[n_mats - 1], d->lm_head, d->vocab, d->hidden);

  // ---- workspace
  const size_t H = d->hidden, I = d->inter;
  size_t off = 0;
  auto take = [&off](size_t bytes) {
    const size_t o = off;
    off = scb_align_up(off + bytes, 256);
    return o;
  };
  const size_t o_bar = take(256);
  const size_t o_cnt = take(256);
  const size_t o_x = take(kRMax * H * 2), o_x2 = take(kRMax * H * 2);
  const size_t o_qkv = take(kRMax * (size_t)qkv_dim * 2), o_attn = take(kRMax * (size_t)attn_dim * 2);
  const size_t o_gu = take(kRMax * I * 2);
  const size_t o_po = take(2 * kRMax * H * 4), o_pd = take(4 * kRMax * H * 4);   // down_proj: up to 4 partial slots (K-halves)
  const size_t n_part = (size_t)kRMax * d->n_q_heads * ns;
  const size_t o_am = take(n_part * 4), o_al = take(n_part * 4), o_ao = take(n_part * d->head_dim * 4);
  const size_t o_cv = take((size_t)grid * kRMax * 4), o_ci = take((size_t)grid * kRMax * 4);
  auto fail = [&](int code) {
    if (h->dev_maps) cudaFree(h->dev_maps);
    if (h->dev_layers) cudaFree(h->dev_layers);
    if (h->dev_ws) cudaFree(h->dev_ws);
    delete[] maps;
    delete[] layers;          // == h->host_layers once the handle has taken it over
    delete h;
    return code;
  };
  if (rc != SCB_OK) return fail(rc);
  if (cudaMalloc(&h->dev_maps, sizeof(CUtensorMap) * n_mats) != cudaSuccess ||
      cudaMalloc(&h->dev_layers, sizeof(MegaLayer) * d->n_layers) != cudaSuccess ||
      cudaMalloc(&h->dev_ws, off) != cudaSuccess) {
    scb_set_error("scb_qwen2_decoder_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(SCB_ERR_CUDA);
  }
  if (cudaMemcpy(h->dev_maps, maps, sizeof(CUtensorMap) * n_mats, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->dev_layers, layers, sizeof(MegaLayer) * d->n_layers, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemset(h->dev_ws, 0, off) != cudaSuccess) {
    scb_set_error("scb_qwen2_decoder_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(SCB_ERR_CUDA);
  }
  delete[] maps;
  maps = nullptr;             // fail() below must not free these twice
  h->host_layers = layers;
  char* ws = reinterpret_cast<char*>(h->dev_ws);
  p.maps = reinterpret_cast<const CUtensorMap*>(h->dev_maps);
  p.layers = reinterpret_cast<const MegaLayer*>(h->dev_layers);
  p.bar_flag = reinterpret_cast<unsigned*>(ws + o_bar);
  p.bar_count = reinterpret_cast<unsigned long long*>(ws + o_cnt);
  p.x_buf = reinterpret_cast<__half*>(ws + o_x);
  p.x2_buf = reinterpret_cast<__half*>(ws + o_x2);
  p.qkv_buf = reinterpret_cast<__half*>(ws + o_qkv);
  p.attn_buf = reinterpret_cast<__half*>(ws + o_attn);
  p.gu_buf = reinterpret_cast<__half*>(ws + o_gu);
  p.part_o = reinterpret_cast<float*>(ws + o_po);
  p.part_d = reinterpret_cast<float*>(ws + o_pd);
  p.att_m = reinterpret_cast<float*>(ws + o_am);
  p.att_l = reinterpret_cast<float*>(ws + o_al);
  p.att_o = reinterpret_cast<float*>(ws + o_ao);
  p.cand_val = reinterpret_cast<float*>(ws + o_cv);
  p.cand_idx = reinterpret_cast<int*>(ws + o_ci);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 8 * 1024) !=
      cudaSuccess) {
    scb_set_error("scb_qwen2_decoder_create: cannot opt in to %d B of shared memory", smem_optin - 8 * 1024);
    return fail(SCB_ERR_CUDA);
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMegaThreads, smem_optin - 8 * 1024) !=
          cudaSuccess ||
      per_sm < 1) {
    scb_set_error("scb_qwen2_decoder_create: kernel does not fit one CTA per SM");
    return fail(SCB_ERR_UNSUPPORTED);
  }
  *out_handle = h;
  return SCB_OK;
}

int scb_qwen2_decoder_destroy(void* handle) {
  MegaHandle* h = reinterpret_cast<MegaHandle*>(handle);
  SCB_REQUIRE(h != nullptr && h->magic == kMagic, SCB_ERR_INVALID_ARG, "scb_qwen2_decoder_destroy: bad handle");
  h->magic = 0;
  cudaFree(h->dev_maps);
  cudaFree(h->dev_layers);
  cudaFree(h->dev_ws);
  if (h->dev_prof) cudaFree(h->dev_prof);
  delete[] h->host_layers;
  delete h;
  return SCB_OK;
}

// Debug: per-CTA clock64 stamps of the LAST step (16 per layer: layer start, then after each prologue /
// stage / grid barrier).  First call arms the recording (returns 0 stamps); later calls copy
// [grid][n_layers][16] int64 to `host_out` (capacity in elements) and return the element count.
long long scb_qwen2_decoder_profile(void* handle, long long* host_out, long long capacity) {
  MegaHandle* h = reinterpret_cast<MegaHandle*>(handle);
  if (h == nullptr || h->magic != kMagic) return -1;
  const long long n = (long long)h->grid * h->base.n_layers * kProfStamps;
  if (!h->dev_prof) {
    if (cudaMalloc(&h->dev_prof, n * sizeof(long long)) != cudaSuccess) return -1;
    cudaMemset(h->dev_prof, 0, n * sizeof(long long));
    h->base.prof = reinterpret_cast<long long*>(h->dev_prof);
    return 0;
  }
  if (host_out == nullptr || capacity < n) return -1;
  if (cudaMemcpy(host_out, h->dev_prof, n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

int scb_qwen2_decode_step(const scb_qwen2_decode_step_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: args is NULL");
  MegaHandle* h = reinterpret_cast<MegaHandle*>(a->decoder);
  SCB_REQUIRE(h != nullptr && h->magic == kMagic, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: bad decoder handle");
  SCB_REQUIRE(a->rows >= 1 && a->rows <= kRMax, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: rows=%d not in 1..%d",
              a->rows, kRMax);
  SCB_REQUIRE(a->next_id && a->kv_len && a->logits, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: NULL buffer");
  SCB_REQUIRE(a->ld_logits >= h->base.V, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: ld_logits < vocab");
  SCB_REQUIRE(a->n_counters == 0 || a->counters, SCB_ERR_INVALID_ARG, "scb_qwen2_decode_step: counters is NULL");
  MegaParams p = h->base;
  p.R = a->rows;
  // K-halves for down_proj pay when the activation rows crowd the weight ring out of shared memory (3 and 4 rows:
  // 6 -> 10 and 4 -> 8 stages; measured 3.82 -> 3.74 and 4.39 -> 3.96 ms per step); one- and two-row steps keep the
  // whole-row layout (fewer tile flushes and partial slots per layer).  SCB_MEGA_KHALF=2 forces it for every width.
  static const int khalf_mode = [] { const char* e = getenv("SCB_MEGA_KHALF"); return e ? atoi(e) : 1; }();
  if (a->rows < 3 && khalf_mode != 2) p.shp[ST_DOWN].g0 = 0;
  p.next_id = reinterpret_cast<long long*>(a->next_id);
  p.kv_len = a->kv_len;
  p.counters = a->counters;
  p.n_counters = a->n_counters;
  p.logits = reinterpret_cast<__half*>(a->logits);
  p.ld_logits = a->ld_logits;
  const long long cap = (a->ctx_cap > 0 && a->ctx_cap < p.max_ctx) ? a->ctx_cap : p.max_ctx;
  // longest KV range one CTA attends (see attn_plan): ceil(cap / S) for one sequence, ceil(R * cap / (S - R)) when the
  // R sequences share the S CTAs of a kv head
  // ctx_sum (optional): an upper bound of the SUM of the rows' lengths, tighter than rows * cap when the batch mixes
  // a long context with short ones (the answer next to two summaries)
  long long sum_cap = (long long)a->rows * cap;
  if (a->ctx_sum > 0 && a->ctx_sum < sum_cap) sum_cap = a->ctx_sum;
  const int per_need = (a->rows == 1 || p.n_splits <= a->rows)
                           ? (int)((cap + p.n_splits - 1) / p.n_splits)
                           : (int)((sum_cap + (p.n_splits - a->rows) - 1) / (p.n_splits - a->rows));
  p.per_max = ((per_need + 11) / 16) * 16 + 4;   // row stride = 4 mod 16 floats: conflict-free score stores
  // shared-memory budget: ring stages + max(activation rows, attention scratch)
  int kmax = p.H;
  // widest activation row a CTA keeps in shared memory: down_proj's input — all of it, or one K-half (unit_plan)
  const int down_cols = p.shp[ST_DOWN].g0 > 0 ? ((p.shp[ST_DOWN].chunks + 1) / 2) * kUnitK : p.I;
  if (down_cols > kmax) kmax = down_cols < p.I ? down_cols : p.I;
  if (p.n_q * h->head_dim > kmax) kmax = p.n_q * h->head_dim;
  p.act_ld = kmax + 8;
  size_t scratch = (size_t)p.R * p.act_ld * 2;
  const size_t attn_b = attn_scratch_bytes(h->head_dim, h->group, p.per_max);
  if (attn_b > scratch) scratch = attn_b;
  const size_t comb = sizeof(float) * (kMaxSplits + 8 + (size_t)kCThreads);
  if (comb > scratch) scratch = comb;
  scratch = scb_align_up(scratch, 1024);
  const long long budget = (long long)h->smem_limit - 8 * 1024 - 1024 - (long long)scratch;
  int n_stage = (int)(budget / kUnitBytes);
  if (n_stage > kMaxStages) n_stage = kMaxStages;
  SCB_REQUIRE(n_stage >= 2, SCB_ERR_UNSUPPORTED,
              "scb_qwen2_decode_step: rows=%d x k=%d (+ctx_cap %lld) leaves no shared memory for the weight ring",
              a->rows, kmax, cap);
  p.n_stage = n_stage;
  const size_t smem = 1024 + (size_t)n_stage * kUnitBytes + scratch;
  MegaKernel kern = pick_kernel(h->head_dim, h->group);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h->grid);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SCB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  SCB_LAUNCH_OK("decode_mega_kernel");
  return SCB_OK;
}

// ------------------------------------------------------------------------------------------
// scb_qwen2_prefill — the long-context prefill of one sequence as ONE native call over the decoder
// handle: per layer RMSNorm -> qkv GEMM (+bias) -> RoPE + KV append -> fused causal attention over the
// cache -> o_proj GEMM (+residual) -> RMSNorm -> gate|up GEMM (SiLU gate) -> down GEMM (+residual), then
// final RMSNorm and lm_head on the LAST position only (modeling_qwen2.py:269-311, :474-475 reached from
// longva/model/language_model/llava_qwen.py:155).  Same kernels, same order and the same rounding points
// as the per-launch path of qwen2.Qwen2Engine.prefill; only the launch loop moves out of Python.
// ------------------------------------------------------------------------------------------
size_t scb_qwen2_prefill_workspace_bytes(void* handle, int64_t n_tokens) {
  MegaHandle* h = reinterpret_cast<MegaHandle*>(handle);
  if (h == nullptr || h->magic != kMagic || n_tokens < 1) return 0;
  const size_t T = (size_t)n_tokens;
  size_t off = scb_align_up(scb_gemv_workspace_bytes(1), 256);
  off += 3 * scb_align_up(T * h->hidden * 2, 256);                       // x ping-pong + normalised rows
  off += scb_align_up(T * (size_t)h->qkv_dim * 2, 256);
  off += scb_align_up(T * (size_t)h->n_q * h->head_dim * 2, 256);
  off += scb_align_up(T * (size_t)h->inter * 2, 256);
  return off;
}

// Several sequences in ONE pass: the rows of all prompts are stacked ([sum(n_tokens), hidden], sequence after
// sequence), every linear layer and norm runs once over the stack (a 5 816-token prompt alone leaves the last wave of
// the o_proj / down_proj GEMMs two-thirds empty: 322 CTA-pair tiles on 74 pairs; five prompts stacked fill it), RoPE +
// KV append and the causal attention run per sequence on its own rows, cache slot and start position.  Per row the
// arithmetic is that of the one-sequence call (rows of a GEMM are independent): results are bit-identical.
int scb_qwen2_prefill_batch(void* handle, const void* inputs_embeds, int32_t n_seqs, const int64_t* n_tokens_host,
                            const int32_t* slots_host, const int64_t* pos0_host, void* logits_rows, int64_t ld_logits,
                            void* workspace, size_t workspace_bytes, void* stream) {
  MegaHandle* h = reinterpret_cast<MegaHandle*>(handle);
  SCB_REQUIRE(h != nullptr && h->magic == kMagic, SCB_ERR_INVALID_ARG, "scb_qwen2_prefill: bad decoder handle");
  SCB_REQUIRE(inputs_embeds && logits_rows && workspace && n_tokens_host && slots_host && pos0_host, SCB_ERR_INVALID_ARG,
              "scb_qwen2_prefill: NULL buffer");
  SCB_REQUIRE(n_seqs >= 1 && n_seqs <= 64, SCB_ERR_INVALID_ARG, "scb_qwen2_prefill: n_seqs=%d not in 1..64", n_seqs);
  SCB_REQUIRE(n_seqs == 1 || ld_logits >= h->vocab, SCB_ERR_INVALID_ARG, "scb_qwen2_prefill: ld_logits < vocab");
  int64_t total = 0;
  for (int q = 0; q < n_seqs; ++q) {
    const int64_t n = n_tokens_host[q], p0 = pos0_host[q];
    const int slot = slots_host[q];
    SCB_REQUIRE(n >= 1 && slot >= 0 && slot < h->max_seqs && p0 >= 0 && p0 + n <= h->base.max_ctx, SCB_ERR_INVALID_ARG,
                "scb_qwen2_prefill: tokens=%lld slot=%d pos0=%lld out of range (max_ctx %lld)", (long long)n, slot,
                (long long)p0, h->base.max_ctx);
    for (int r = 0; r < q; ++r)
      SCB_REQUIRE(slots_host[r] != slot, SCB_ERR_INVALID_ARG, "scb_qwen2_prefill: cache slot %d used twice in one batch", slot);
    total += n;
  }
  const size_t need = scb_qwen2_prefill_workspace_bytes(handle, total);
  SCB_REQUIRE(workspace_bytes >= need, SCB_ERR_WORKSPACE, "scb_qwen2_prefill: workspace %zu B < required %zu B",
              workspace_bytes, need);
  const MegaParams& p = h->base;
  const size_t T = (size_t)total;
  const int H = h->hidden, I = h->inter, hd = h->head_dim, QKV = h->qkv_dim, A = h->n_q * hd;
  char* ws = reinterpret_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = ws + off;
    off += scb_align_up(bytes, 256);
    return q;
  };
  char* gemv_ws = take(scb_gemv_workspace_bytes(1));     // zero before the first call (ticket), left zero
  __half* xa = reinterpret_cast<__half*>(take(T * H * 2));
  __half* xb = reinterpret_cast<__half*>(take(T * H * 2));
  __half* y = reinterpret_cast<__half*>(take(T * H * 2));
  __half* qkv = reinterpret_cast<__half*>(take(T * (size_t)QKV * 2));
  __half* ctx = reinterpret_cast<__half*>(take(T * (size_t)A * 2));
  __half* gu = reinterpret_cast<__half*>(take(T * (size_t)I * 2));
  auto linear = [&](const void* a, int64_t k, const void* w, int64_t n, const void* bias, const void* res, void* out,
                    int act) {
    scb_linear_args la;
    memset(&la, 0, sizeof(la));
    la.a = a; la.w = w; la.bias = bias; la.residual = res; la.out = out;
    la.m = total; la.n = n; la.k = k;
    la.lda = k; la.ldw = k; la.ldr = n; la.ldo = act == SCB_ACT_SILU_PAIR ? n / 2 : n;
    la.act = act; la.out_dtype = SCB_F16; la.batch = 1; la.alpha = 1.f;
    return scb_linear(&la, stream);
  };
  const __half* x = reinterpret_cast<const __half*>(inputs_embeds);
  __half* xn = xa;          // output of the current layer's o_proj (x2), then of down (the next x)
  const float scale = 1.f / sqrtf((float)hd);
  int rc = SCB_OK;
  for (int l = 0; l < p.n_layers && rc == SCB_OK; ++l) {
    const MegaLayer& ly = h->host_layers[l];
    rc = scb_rmsnorm(x, H, ly.ln1, p.eps, total, H, y, H, stream);
    if (rc == SCB_OK) rc = linear(y, H, ly.w[0], QKV, ly.qkv_b, nullptr, qkv, SCB_ACT_NONE);
    int64_t row0 = 0;
    for (int q = 0; q < n_seqs && rc == SCB_OK; ++q) {
      const int64_t n = n_tokens_host[q], p0 = pos0_host[q];
      __half* kc = ly.k_cache + (size_t)slots_host[q] * p.cache_seq_stride;
      __half* vc = ly.v_cache + (size_t)slots_host[q] * p.cache_seq_stride;
      __half* qrows = qkv + (size_t)row0 * QKV;
      rc = scb_rope_append(qrows, QKV, n, h->n_q, h->n_kv, hd, nullptr, (int32_t)p0, p.theta, p.inv_freq, kc, vc,
                           p.max_ctx, stream);
      if (rc == SCB_OK) {
        scb_attention_args aa;
        memset(&aa, 0, sizeof(aa));
        aa.q = qrows; aa.k = kc; aa.v = vc; aa.out = ctx + (size_t)row0 * A;
        aa.batch = 1; aa.n_q_heads = h->n_q; aa.n_kv_heads = h->n_kv; aa.head_dim = hd;
        aa.q_len = n; aa.kv_len = p0 + n;
        aa.q_stride_t = QKV; aa.q_stride_h = hd;
        aa.k_stride_t = hd; aa.k_stride_h = p.max_ctx * hd;
        aa.v_stride_t = hd; aa.v_stride_h = p.max_ctx * hd;
        aa.o_stride_t = A; aa.o_stride_h = hd;
        aa.scale = scale; aa.causal = 1;
        rc = scb_attention(&aa, stream);
      }
      row0 += n;
    }
    __half* x2 = (x == xa) ? xb : xa;                       // never the buffer x lives in
    if (rc == SCB_OK) rc = linear(ctx, A, ly.w[1], H, nullptr, x, x2, SCB_ACT_NONE);
    if (rc == SCB_OK) rc = scb_rmsnorm(x2, H, ly.ln2, p.eps, total, H, y, H, stream);
    if (rc == SCB_OK) rc = linear(y, H, ly.w[2], 2ll * I, nullptr, nullptr, gu, SCB_ACT_SILU_PAIR);
    xn = (x2 == xa) ? xb : xa;
    if (rc == SCB_OK) rc = linear(gu, I, ly.w[3], H, nullptr, x2, xn, SCB_ACT_NONE);
    x = xn;
  }
  if (rc != SCB_OK) return rc;
  // final norm + lm_head on the last position of every sequence
  int64_t row0 = 0;
  for (int q = 0; q < n_seqs; ++q) {
    row0 += n_tokens_host[q];
    rc = scb_rmsnorm(x + (size_t)(row0 - 1) * H, H, p.final_norm, p.eps, 1, H, y, H, stream);
    if (rc != SCB_OK) return rc;
    scb_gemv_args ga;
    memset(&ga, 0, sizeof(ga));
    ga.a = y; ga.w = p.lm_head; ga.out = reinterpret_cast<__half*>(logits_rows) + (size_t)q * ld_logits;
    ga.n = h->vocab; ga.k = H; ga.lda = H; ga.ldw = H; ga.ldo = h->vocab;
    ga.rows = 1; ga.prologue = SCB_PRO_NONE; ga.out_dtype = SCB_F16;
    ga.workspace = gemv_ws; ga.workspace_bytes = scb_gemv_workspace_bytes(1);
    rc = scb_gemv(&ga, stream);
    if (rc != SCB_OK) return rc;
  }
  return SCB_OK;
}

int scb_qwen2_prefill(void* handle, const void* inputs_embeds, int64_t n_tokens, int32_t slot, int64_t pos0,
                      void* logits_row, void* workspace, size_t workspace_bytes, void* stream) {
  return scb_qwen2_prefill_batch(handle, inputs_embeds, 1, &n_tokens, &slot, &pos0, logits_row, 0, workspace,
                                 workspace_bytes, stream);
}

}  // extern "C"
