// scb_api.cu — library-wide C-ABI services: version, thread-local error text, launch counter.
#include "scb_common.cuh"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = {0};
static std::atomic<uint64_t> g_launches{0};

void scb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void scb_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int scb_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool scb_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

extern "C" {

int scb_version(void) { return 100; /* 0.1.0 */ }

const char* scb_last_error(void) { return g_err; }

uint64_t scb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
