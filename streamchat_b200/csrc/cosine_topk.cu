// cosine_topk.cu — cosine-similarity top-k over a row-major embedding index (HBM-bound).
//
// Replaces sentence_transformers.util.cos_sim + running arg-max of the memory-tree search
// (reference utiles.py:732-742,764-772) and the FAISS IndexFlatL2 search behind
// memory_bank/memory_retrieval/local_doc_qa.py:140,270.
//
// One launch per call.  grid = (gx, n_queries); every warp streams whole rows with 128-bit
// L1-bypassing loads (8 x 16 B in flight per lane for a 1024-d fp32 row), reduces <x,q> and
// <x,x> with shuffles in one pass (rows are not pre-normalised) and keeps a lane-distributed
// sorted top-k list in registers.  Warps merge through shared memory, blocks through a
// candidate buffer in the workspace; the last block to finish (atomic ticket) produces the
// final (score desc, row asc) ordered answer, so no second launch is needed.
//
// Algorithmic bytes per launch: n_rows*dim*sizeof(T) + n_queries*dim*sizeof(Tq) + n_queries*k*12.
#include "scb_common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 32;
constexpr int kMaxGx = 1024;

struct TopkList {
  float s;       // this lane's entry (lane i = i-th best)
  long long i;
};

// true if candidate (s, i) ranks strictly before entry (es, ei): score desc, row asc.
SCB_DEVICE bool ranks_before(float s, long long i, float es, long long ei) {
  return (s > es) || (s == es && i < ei);
}

// Warp-collective insert of (s, idx) into the lane-distributed sorted list of length k.
SCB_DEVICE void topk_insert(TopkList& l, float s, long long idx, int k, int lane) {
  // entries that stay ahead of the candidate
  bool ahead = (lane < k) && !ranks_before(s, idx, l.s, l.i);
  unsigned m = __ballot_sync(0xffffffffu, ahead);
  int pos = __popc(m);
  if (pos >= k) return;  // warp-uniform
  float ps = __shfl_up_sync(0xffffffffu, l.s, 1);
  long long pi = __shfl_up_sync(0xffffffffu, l.i, 1);
  if (lane == pos) {
    l.s = s;
    l.i = idx;
  } else if (lane > pos) {
    l.s = ps;
    l.i = pi;
  }
}

// Compare-exchange step of a bitonic network over the 32 lanes: after it, in an ascending-rank ("best first")
// block the lower lane of each pair holds the entry that ranks first.  Entries with i < 0 are empty (-inf, -1).
SCB_DEVICE void bitonic_step(float& s, long long& i, int stride, bool best_first, int lane) {
  const float os = __shfl_xor_sync(0xffffffffu, s, stride);
  const long long oi = __shfl_xor_sync(0xffffffffu, i, stride);
  const bool lower = (lane & stride) == 0;
  const bool other_first = ranks_before(os, oi, s, i);
  const bool take = (lower == best_first) ? other_first : !other_first;
  if (take) {
    s = os;
    i = oi;
  }
}
// Full sort of 32 lane-distributed entries, best first (lane 0 = best).
SCB_DEVICE void bitonic_sort32(float& s, long long& i, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
    const bool best_first = (size == 32) || ((lane & size) == 0);
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) bitonic_step(s, i, stride, best_first, lane);
  }
}
// (ls, li): sorted best-first list of 32; (s, i): 32 new entries in any order.  Leaves the best 32 of the 64 in
// (ls, li), sorted: sort the new ones, pair list[l] with new[31 - l] (the better of each pair forms a bitonic
// sequence holding the top 32), then one bitonic merge.  Constant cost, no data-dependent loop.
SCB_DEVICE void merge_top32(float& ls, long long& li, float s, long long i, int lane) {
  bitonic_sort32(s, i, lane);
  const float rs = __shfl_sync(0xffffffffu, s, 31 - lane);
  const long long ri = __shfl_sync(0xffffffffu, i, 31 - lane);
  if (ranks_before(rs, ri, ls, li)) {
    ls = rs;
    li = ri;
  }
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) bitonic_step(ls, li, stride, true, lane);
}

// Select the best k of the kWarps sorted lists in shared memory (list w at [w * kMaxK, w * kMaxK + k)) by COUNTING:
// thread (w, j) counts the entries that rank before its own — rows are distinct, so the ranks are a permutation —
// and the entries of rank < k land in out[rank].  All 256 threads work in parallel on <= 256 entries (the serial
// insertion this replaces cost ~100 cycles per entry on one warp).  out_* must hold (-inf, -1) beforehand.
SCB_DEVICE void select_topk_counting(const float* m_s, const long long* m_i, int k, float* out_s, long long* out_i,
                                     int tid) {
  if (tid < kWarps * k) {
    const int w = tid / k, j = tid - w * k;
    const float s = m_s[w * kMaxK + j];
    const long long i = m_i[w * kMaxK + j];
    if (i >= 0) {
      int rank = 0;
      for (int w2 = 0; w2 < kWarps; ++w2)
        for (int j2 = 0; j2 < k; ++j2) {
          const long long i2 = m_i[w2 * kMaxK + j2];
          rank += (i2 >= 0 && ranks_before(m_s[w2 * kMaxK + j2], i2, s, i)) ? 1 : 0;
        }
      if (rank < k) {
        out_s[rank] = s;
        out_i[rank] = i;
      }
    }
  }
}

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int kElems = 4;
  SCB_DEVICE static void fma(const uint4& v, const float* q, float& dot, float& xx) {
    float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z),
          x3 = __uint_as_float(v.w);
    dot = fmaf(x0, q[0], dot); xx = fmaf(x0, x0, xx);
    dot = fmaf(x1, q[1], dot); xx = fmaf(x1, x1, xx);
    dot = fmaf(x2, q[2], dot); xx = fmaf(x2, x2, xx);
    dot = fmaf(x3, q[3], dot); xx = fmaf(x3, x3, xx);
  }
  SCB_DEVICE static float load1(const float* p) { return *p; }
};
template <>
struct Vec<__half> {
  static constexpr int kElems = 8;
  SCB_DEVICE static void fma(const uint4& v, const float* q, float& dot, float& xx) {
    float2 a = h2_to_f2(v.x), b = h2_to_f2(v.y), c = h2_to_f2(v.z), d = h2_to_f2(v.w);
    dot = fmaf(a.x, q[0], dot); xx = fmaf(a.x, a.x, xx);
    dot = fmaf(a.y, q[1], dot); xx = fmaf(a.y, a.y, xx);
    dot = fmaf(b.x, q[2], dot); xx = fmaf(b.x, b.x, xx);
    dot = fmaf(b.y, q[3], dot); xx = fmaf(b.y, b.y, xx);
    dot = fmaf(c.x, q[4], dot); xx = fmaf(c.x, c.x, xx);
    dot = fmaf(c.y, q[5], dot); xx = fmaf(c.y, c.y, xx);
    dot = fmaf(d.x, q[6], dot); xx = fmaf(d.x, d.x, xx);
    dot = fmaf(d.y, q[7], dot); xx = fmaf(d.y, d.y, xx);
  }
  SCB_DEVICE static float load1(const __half* p) { return __half2float(*p); }
};

// kVec: rows are 16-byte aligned and are read with 128-bit loads; otherwise scalar loads.
// kChunks > 0: dim == kChunks * 32 * Vec<T>::kElems exactly, all loads of a row issued up front.
template <typename T, bool kVec, int kChunks>
__global__ void __launch_bounds__(kThreads)
cosine_topk_kernel(const T* __restrict__ index, const void* __restrict__ queries, int query_is_f16,
                   long long n_rows, int dim, long long ld_index, long long ld_query, int k,
                   float* __restrict__ cand_s, long long* __restrict__ cand_i,
                   unsigned* __restrict__ tickets, long long* __restrict__ out_i,
                   float* __restrict__ out_s) {
  extern __shared__ __align__(16) float smem[];
  float* q_s = smem;  // [dim_padded]
  __shared__ float red_s[kWarps];
  __shared__ float m_s[kWarps * kMaxK];
  __shared__ long long m_i[kWarps * kMaxK];
  __shared__ float sel_s[kMaxK];
  __shared__ long long sel_i[kMaxK];
  __shared__ unsigned is_last_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qi = blockIdx.y;
  const int gx = gridDim.x;

  // ---- the first row's loads do not depend on the query: they are in flight while it is staged
  constexpr int E = Vec<T>::kElems;
  constexpr int KC = kChunks > 0 ? kChunks : 1;
  const long long warp_stride = (long long)gx * kWarps;
  long long r0 = (long long)blockIdx.x * kWarps + warp;
  uint4 va[KC];
  if constexpr (kVec && kChunks > 0) {
    if (r0 < n_rows) {
      const T* x0 = index + r0 * ld_index;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) va[c] = ld_stream_u4(x0 + (c * 32 + lane) * E);
    }
  }

  // ---- stage the query in shared memory as fp32 and compute |q|
  float qq = 0.f;
  for (int d = tid; d < dim; d += kThreads) {
    float v = query_is_f16
                  ? __half2float(reinterpret_cast<const __half*>(queries)[qi * ld_query + d])
                  : reinterpret_cast<const float*>(queries)[qi * ld_query + d];
    q_s[d] = v;
    qq = fmaf(v, v, qq);
  }
  qq = warp_sum(qq);
  if (lane == 0) red_s[warp] = qq;
  __syncthreads();
  float qtot = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) qtot += red_s[w];
  const float qn = fmaxf(sqrtf(qtot), 1e-12f);

  TopkList list;
  list.s = -INFINITY;
  list.i = -1;

  if constexpr (kVec && kChunks > 0) {
    // Software pipeline over this warp's rows: the 16-byte loads of row r+1 are in flight while row r is reduced
    // and inserted, so a warp's rows cost one memory latency in total instead of one each (small indices are
    // latency-bound: 10 000 x 1024 fp32 is 6 us of traffic).
    // (two explicit register sets: a single set that is refilled after its last use would only overlap the loads
    // with the shuffle reduction, not with the whole row)
    auto load_row = [&](uint4 (&v)[kChunks], long long r) {
      const T* x = index + r * ld_index;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) v[c] = ld_stream_u4(x + (c * 32 + lane) * E);
    };
    auto consume = [&](const uint4 (&v)[kChunks], long long r) {
      float dot = 0.f, xx = 0.f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) Vec<T>::fma(v[c], q_s + (c * 32 + lane) * E, dot, xx);
      dot = warp_sum(dot);
      xx = warp_sum(xx);
      const float sim = dot / (fmaxf(sqrtf(xx), 1e-12f) * qn);
      topk_insert(list, sim, r, k, lane);
    };
    uint4 vb[kChunks];
    while (r0 < n_rows) {
      const long long r1 = r0 + warp_stride;
      if (r1 < n_rows) load_row(vb, r1);
      consume(va, r0);
      if (r1 >= n_rows) break;
      const long long r2 = r1 + warp_stride;
      if (r2 < n_rows) load_row(va, r2);
      consume(vb, r1);
      r0 = r2;
    }
  } else
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < n_rows; row += warp_stride) {
    const T* x = index + row * ld_index;
    float dot = 0.f, xx = 0.f;
    if constexpr (kVec) {
      const int nvec = dim / E;
      for (int c = lane; c < nvec; c += 32) {
        uint4 v = ld_stream_u4(x + c * E);
        Vec<T>::fma(v, q_s + c * E, dot, xx);
      }
      for (int d = nvec * E + lane; d < dim; d += 32) {
        float xv = Vec<T>::load1(x + d);
        dot = fmaf(xv, q_s[d], dot);
        xx = fmaf(xv, xv, xx);
      }
    } else {
      for (int d = lane; d < dim; d += 32) {
        float xv = Vec<T>::load1(x + d);
        dot = fmaf(xv, q_s[d], dot);
        xx = fmaf(xv, xv, xx);
      }
    }
    dot = warp_sum(dot);
    xx = warp_sum(xx);
    const float sim = dot / (fmaxf(sqrtf(xx), 1e-12f) * qn);
    topk_insert(list, sim, row, k, lane);
  }

  // ---- block merge: every warp publishes its list; the block's best k are selected by counting ranks
  if (lane < k) {
    m_s[warp * kMaxK + lane] = list.s;
    m_i[warp * kMaxK + lane] = list.i;
  }
  if (tid < kMaxK) {
    sel_s[tid] = -INFINITY;
    sel_i[tid] = -1;
  }
  __syncthreads();
  select_topk_counting(m_s, m_i, k, sel_s, sel_i, tid);
  __syncthreads();
  if (warp == 0) {
    if (lane < k) {
      cand_s[((long long)qi * gx + blockIdx.x) * k + lane] = sel_s[lane];
      cand_i[((long long)qi * gx + blockIdx.x) * k + lane] = sel_i[lane];
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      unsigned t = atomicAdd(&tickets[qi], 1u);
      is_last_s = (t == (unsigned)gx - 1u) ? 1u : 0u;
    }
  }
  __syncthreads();
  if (!is_last_s) return;
  __threadfence();

  // ---- final merge by the last block of this query.  Every warp keeps a sorted best-32 list and folds its share
  // of the gx*k candidates into it 32 at a time with bitonic networks (constant cost per 32 candidates, no
  // data-dependent insertion loop); a round of kWarps*32*4 candidates is requested up front (one L2 round trip,
  // and one round covers all candidates at config #3's size).  The 8 lists then meet by counting ranks.
  float ls = -INFINITY;
  long long li = -1;
  {
    constexpr int kPer = 4;
    const int n_cand = gx * k;
    const float* cs = cand_s + (long long)qi * n_cand;
    const long long* ci = cand_i + (long long)qi * n_cand;
    bool first = true;
    for (int base = 0; base < n_cand; base += kWarps * 32 * kPer) {
      float s_l[kPer];
      long long i_l[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int c = base + (j * kWarps + warp) * 32 + lane;
        s_l[j] = -INFINITY;
        i_l[j] = -1;
        if (c < n_cand) {
          s_l[j] = __ldcg(cs + c);
          i_l[j] = __ldcg(ci + c);
        }
      }
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (base + (j * kWarps + warp) * 32 >= n_cand) break;      // warp-uniform: nothing left for this warp
        if (first) {
          ls = s_l[j];
          li = i_l[j];
          bitonic_sort32(ls, li, lane);
          first = false;
        } else {
          merge_top32(ls, li, s_l[j], i_l[j], lane);
        }
      }
    }
  }
  __syncthreads();  // m_s/m_i reuse
  if (lane < k) {
    m_s[warp * kMaxK + lane] = ls;
    m_i[warp * kMaxK + lane] = li;
  }
  if (tid < kMaxK) {
    sel_s[tid] = -INFINITY;
    sel_i[tid] = -1;
  }
  __syncthreads();
  select_topk_counting(m_s, m_i, k, sel_s, sel_i, tid);
  __syncthreads();
  if (warp == 0) {
    if (lane < k) {
      out_s[(long long)qi * k + lane] = sel_s[lane];
      out_i[(long long)qi * k + lane] = sel_i[lane];
    }
    if (lane == 0) tickets[qi] = 0;  // leave the workspace reusable
  }
}

struct Layout {
  size_t tickets_off, cand_s_off, cand_i_off, total;
};
Layout make_layout(int n_queries, int k) {
  Layout l;
  size_t off = 0;
  l.tickets_off = off;
  off = scb_align_up(off + sizeof(unsigned) * (size_t)n_queries, 256);
  l.cand_i_off = off;
  off = scb_align_up(off + sizeof(long long) * (size_t)n_queries * kMaxGx * k, 256);
  l.cand_s_off = off;
  off = scb_align_up(off + sizeof(float) * (size_t)n_queries * kMaxGx * k, 256);
  l.total = off;
  return l;
}

template <typename T>
int launch(const scb_cosine_topk_args* a, cudaStream_t st) {
  Layout lay = make_layout(a->n_queries, a->k);
  char* ws = reinterpret_cast<char*>(a->workspace);
  unsigned* tickets = reinterpret_cast<unsigned*>(ws + lay.tickets_off);
  float* cand_s = reinterpret_cast<float*>(ws + lay.cand_s_off);
  long long* cand_i = reinterpret_cast<long long*>(ws + lay.cand_i_off);
  // tickets: zero before the first call (the caller allocates the workspace zeroed), left zero by the last block —
  // no memset launch on the path (a third of the time of a 10 000-row search)

  const int sms = scb_num_sms();
  // Grid: ONE wave whenever the index is small enough for it.  Two 256-thread blocks are resident per SM (95
  // registers), so 2 x SMs blocks start together; a few blocks more would run as a second wave behind the first.
  // Measured on the 10 000-row index (48 back-to-back searches, us per search): 626 blocks 34.1, 313 blocks 27.4
  // (two waves), 296 blocks 22.1, 157 blocks 19.8 — the per-block merge / ticket / candidate traffic outweighs the
  // shorter per-warp row chain, so the minimum is 8 software-pipelined rows per warp.  Large indices take 4 x SMs
  // blocks of many rows each.  (SCB_TOPK_RPW: tuning aid, minimum rows per warp.)
  static const int rpw = [] { const char* e = getenv("SCB_TOPK_RPW"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
  const long long one_wave = 2ll * sms;
  long long want = (a->n_rows + kWarps * rpw - 1) / (kWarps * rpw);
  if (want > one_wave) want = (a->n_rows <= one_wave * kWarps * 16) ? one_wave : (want > 4ll * sms ? 4ll * sms : want);
  int gx = (int)(want < 1 ? 1 : want);
  if (gx > kMaxGx) gx = kMaxGx;
  dim3 grid(gx, a->n_queries);
  size_t smem = sizeof(float) * (size_t)a->dim;
  constexpr int E = Vec<T>::kElems;
  const bool vec = ((uintptr_t)a->index % 16 == 0) && ((a->ld_index * sizeof(T)) % 16 == 0);
  const T* idx = reinterpret_cast<const T*>(a->index);
  const int qf16 = a->query_dtype == SCB_F16;

#define SCB_TOPK_LAUNCH(VEC, CH)                                                               \
  do {                                                                                         \
    auto kern = cosine_topk_kernel<T, VEC, CH>;                                                \
    if (smem > 48 * 1024)                                                                      \
      SCB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                       (int)smem));                                            \
    kern<<<grid, kThreads, smem, st>>>(idx, a->queries, qf16, a->n_rows, a->dim, a->ld_index,  \
                                       a->ld_query, a->k, cand_s, cand_i, tickets,             \
                                       reinterpret_cast<long long*>(a->out_indices),           \
                                       a->out_scores);                                         \
  } while (0)

  if (vec && a->dim == 8 * 32 * E) {
    SCB_TOPK_LAUNCH(true, 8);
  } else if (vec && a->dim == 4 * 32 * E) {
    SCB_TOPK_LAUNCH(true, 4);
  } else if (vec && a->dim == 3 * 32 * E) {
    SCB_TOPK_LAUNCH(true, 3);
  } else if (vec) {
    SCB_TOPK_LAUNCH(true, 0);
  } else {
    SCB_TOPK_LAUNCH(false, 0);
  }
#undef SCB_TOPK_LAUNCH
  SCB_LAUNCH_OK("cosine_topk_kernel");
  return SCB_OK;
}

}  // namespace

extern "C" {

size_t scb_cosine_topk_workspace_bytes(int64_t n_rows, int32_t n_queries, int32_t k) {
  (void)n_rows;
  if (n_queries < 1 || k < 1 || k > kMaxK) return 0;
  return make_layout(n_queries, k).total;
}

int scb_cosine_topk(const scb_cosine_topk_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_cosine_topk: args is NULL");
  SCB_REQUIRE(a->k >= 1 && a->k <= kMaxK, SCB_ERR_INVALID_ARG, "scb_cosine_topk: k=%d not in 1..%d",
              a->k, kMaxK);
  SCB_REQUIRE(a->n_queries >= 1 && a->n_queries <= 65535, SCB_ERR_INVALID_ARG,
              "scb_cosine_topk: n_queries=%d not in 1..65535", a->n_queries);
  SCB_REQUIRE(a->n_rows >= 0, SCB_ERR_INVALID_ARG, "scb_cosine_topk: n_rows < 0");
  SCB_REQUIRE(a->dim >= 1 && a->dim <= 32768, SCB_ERR_INVALID_ARG,
              "scb_cosine_topk: dim=%d not in 1..32768", a->dim);
  SCB_REQUIRE(a->ld_index >= a->dim && a->ld_query >= a->dim, SCB_ERR_INVALID_ARG,
              "scb_cosine_topk: leading dimension smaller than dim");
  SCB_REQUIRE(a->queries && a->out_indices && a->out_scores && (a->index || a->n_rows == 0),
              SCB_ERR_INVALID_ARG, "scb_cosine_topk: NULL buffer");
  SCB_REQUIRE(a->index_dtype == SCB_F32 || a->index_dtype == SCB_F16, SCB_ERR_UNSUPPORTED,
              "scb_cosine_topk: index dtype %d unsupported", a->index_dtype);
  SCB_REQUIRE(a->query_dtype == SCB_F32 || a->query_dtype == SCB_F16, SCB_ERR_UNSUPPORTED,
              "scb_cosine_topk: query dtype %d unsupported", a->query_dtype);
  size_t need = scb_cosine_topk_workspace_bytes(a->n_rows, a->n_queries, a->k);
  SCB_REQUIRE(a->workspace && a->workspace_bytes >= need, SCB_ERR_WORKSPACE,
              "scb_cosine_topk: workspace %zu B < required %zu B", a->workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->index_dtype == SCB_F32) return launch<float>(a, st);
  return launch<__half>(a, st);
}

}  // extern "C"
