// peer_gather.cu — scb_allgather_embed: the ONE exchange step of the frame-sharded stream as a push
// all-gather over NVLink / NVSwitch peer memory, written straight into the buffer the consumers read (tree merges,
// the answer's visual context, cosine top-k): no staging copy, no repack.
//
// Every rank owns a gather buffer [world][slice_bytes] (+ a flag block) allocated by scb_peer_alloc — plain
// cudaMalloc memory, exportable with CUDA IPC — and maps every peer's buffer with scb_peer_open.  One call per rank:
//   1. copy kernel: the rank's own slice (already in place at offset rank*slice_bytes of its buffer) is stored into
//      the same offset of every peer's buffer with coalesced 16-byte accesses; the grid is cut so that every SM
//      streams an equal share of (world-1) x slice_bytes;
//   2. the last CTA to finish publishes flag[rank] = epoch in every peer's flag block (system-scope release after a
//      __threadfence_system), then one warp polls its OWN flag block until every peer's epoch arrived (acquire):
//      when the kernel ends, the whole gathered buffer is visible to the kernels that follow it on the stream.
// NVSwitch gives every pair full bandwidth, so there is no ring ordering: all peers are written concurrently.
// The reference has no counterpart (two serial replicas).
#include "scb_common.cuh"
#include <string.h>

namespace {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 512;
constexpr int kFlagBytes = 256;   // world uint32 epochs + the CTA ticket at [kMaxWorld]

struct PeerTable {
  char* buf[kMaxWorld];
  unsigned* flags[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads) peer_allgather_kernel(PeerTable t, int rank, int world, long long slice_bytes,
                                                                  unsigned epoch) {
  // epoch == 0: the call number lives on the DEVICE (own flag block, word kMaxWorld + 1) — every CTA reads it before
  // it takes its ticket, the last CTA stores the incremented value at the very end — so the launch can be replayed from
  // a CUDA graph (the per-layer exchanges of a tensor-parallel decode step)
  unsigned* epoch_dev = t.flags[rank] + kMaxWorld + 1;
  const bool dev_epoch = epoch == 0;
  if (dev_epoch) epoch = *reinterpret_cast<volatile unsigned*>(epoch_dev) + 1u;
  const char* src = t.buf[rank] + (long long)rank * slice_bytes;
  const long long n16 = slice_bytes / 16;
  const long long total = n16 * (world - 1);
  const long long stride = (long long)gridDim.x * kThreads;
  // flat index over (peer, 16-byte piece): consecutive threads write consecutive pieces of one peer's slice
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const int pi = (int)(i / n16);
    const long long off = (i - (long long)pi * n16) * 16;
    const int peer = pi < rank ? pi : pi + 1;
    const uint4 v = ld_stream_u4(src + off);
    *reinterpret_cast<uint4*>(t.buf[peer] + (long long)rank * slice_bytes + off) = v;
  }
  // ---- publish: every store of this CTA is ordered before the ticket; the last CTA raises the flags
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned last;
  unsigned* ticket = t.flags[rank] + kMaxWorld;
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < world && threadIdx.x != rank) st_release_sys(t.flags[threadIdx.x] + rank, epoch);
  if (threadIdx.x == 0) *ticket = 0;                       // self-resetting for the next call
  // ---- wait for every peer's slice to have landed in MY buffer
  if (threadIdx.x < world && threadIdx.x != rank) {
    const unsigned* f = t.flags[rank] + threadIdx.x;
    // bounded: a peer that died or never launched must not hang this GPU — after ~20 s the wait gives up and records
    // the fact in the flag block (word kMaxWorld + 2), which the host can read back (the data is then incomplete)
    long long spins = 0;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(200);
      if (++spins > 50000000ll) {
        t.flags[rank][kMaxWorld + 2] = 1u + (unsigned)threadIdx.x;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (dev_epoch && threadIdx.x == 0) *epoch_dev = epoch;
}

}  // namespace

extern "C" {

int scb_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
  SCB_REQUIRE(dev_ptr != nullptr && ipc_handle_out != nullptr && bytes > 0, SCB_ERR_INVALID_ARG,
              "scb_peer_alloc: NULL argument or zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes in the C-ABI");
  *dev_ptr = nullptr;
  void* p = nullptr;
  SCB_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    scb_set_error("scb_peer_alloc: %s", cudaGetErrorString(e));
    return SCB_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return SCB_OK;
}

int scb_peer_free(void* dev_ptr) {
  SCB_REQUIRE(dev_ptr != nullptr, SCB_ERR_INVALID_ARG, "scb_peer_free: NULL pointer");
  SCB_CUDA_OK(cudaFree(dev_ptr));
  return SCB_OK;
}

int scb_peer_open(const void* ipc_handle, void** mapped_ptr) {
  SCB_REQUIRE(ipc_handle != nullptr && mapped_ptr != nullptr, SCB_ERR_INVALID_ARG, "scb_peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  *mapped_ptr = nullptr;
  SCB_CUDA_OK(cudaIpcOpenMemHandle(mapped_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SCB_OK;
}

int scb_peer_close(void* mapped_ptr) {
  SCB_REQUIRE(mapped_ptr != nullptr, SCB_ERR_INVALID_ARG, "scb_peer_close: NULL pointer");
  SCB_CUDA_OK(cudaIpcCloseMemHandle(mapped_ptr));
  return SCB_OK;
}

size_t scb_allgather_embed_buffer_bytes(int64_t slice_bytes, int32_t world) {
  if (slice_bytes <= 0 || slice_bytes % 16 || world < 1 || world > kMaxWorld) return 0;
  return scb_align_up((size_t)slice_bytes * world, 256) + kFlagBytes;
}

int scb_allgather_embed(const scb_allgather_embed_args* a, void* stream) {
  SCB_REQUIRE(a != nullptr, SCB_ERR_INVALID_ARG, "scb_allgather_embed: args is NULL");
  SCB_REQUIRE(a->world >= 1 && a->world <= kMaxWorld && a->rank >= 0 && a->rank < a->world, SCB_ERR_INVALID_ARG,
              "scb_allgather_embed: rank=%d world=%d (world <= %d)", a->rank, a->world, kMaxWorld);
  SCB_REQUIRE(a->slice_bytes > 0 && a->slice_bytes % 16 == 0, SCB_ERR_INVALID_ARG,
              "scb_allgather_embed: slice_bytes=%lld must be a positive multiple of 16", (long long)a->slice_bytes);
  SCB_REQUIRE(a->peer_bufs != nullptr, SCB_ERR_INVALID_ARG, "scb_allgather_embed: peer_bufs is NULL");
  if (a->world == 1) return SCB_OK;
  PeerTable t;
  memset(&t, 0, sizeof(t));
  const size_t flag_off = scb_align_up((size_t)a->slice_bytes * a->world, 256);
  for (int r = 0; r < a->world; ++r) {
    SCB_REQUIRE(a->peer_bufs[r] != nullptr, SCB_ERR_INVALID_ARG, "scb_allgather_embed: peer_bufs[%d] is NULL", r);
    SCB_REQUIRE(((uintptr_t)a->peer_bufs[r]) % 16 == 0, SCB_ERR_INVALID_ARG, "scb_allgather_embed: buffer %d unaligned", r);
    t.buf[r] = reinterpret_cast<char*>(a->peer_bufs[r]);
    t.flags[r] = reinterpret_cast<unsigned*>(t.buf[r] + flag_off);
  }
  // one wave: every CTA must be resident while the last one spins on the peers' flags
  int grid = 2 * scb_num_sms();
  const long long pieces = a->slice_bytes / 16 * (a->world - 1);
  const long long need = (pieces + kThreads - 1) / kThreads;
  if (need < grid) grid = (int)(need < 1 ? 1 : need);
  peer_allgather_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, a->rank, a->world,
                                                                                      a->slice_bytes, a->epoch);
  SCB_LAUNCH_OK("peer_allgather_kernel");
  return SCB_OK;
}

}  // extern "C"
