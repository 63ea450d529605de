// preprocess.cu — frame ingest (SURVEY §8 a-V1): the CLIPImageProcessor arithmetic of
// `process_images` (utiles.py:71-87) on uint8 frames resident in HBM: Pillow's two-pass 8-bit bicubic
// resample (fixed-point coefficients, 22 fractional bits, each pass rounded to uint8 through clip8),
// centre crop, x 1/255, (x - mean) / std and the HWC -> CHW transpose.  The coefficient tables and
// the 3 x 256 value table come from the host (streamchat_b200/preprocess.py); the kernels are the
// per-pixel integer convolutions — HBM-bound byte work, one thread per output pixel, all three
// channels, contiguous 3-byte pixels across the warp.
#include "scb_common.cuh"
#include "../../include/streamchat_b200.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

SCB_DEVICE unsigned char clip8(int v) {
  v >>= kPrecisionBits;                      // arithmetic shift, then clamp (clip8_lookups)
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: dst[n, r, x, c] = clip8(2^21 + sum_j src[n, r, x0(x)+j, c] * k[x][j])
__global__ void __launch_bounds__(256)
resample_h_kernel(const unsigned char* __restrict__ src, long long img_stride, long long row_stride, int rows,
                  const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize, int out_w,
                  unsigned char* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const long long n = blockIdx.z;
  if (x >= out_w) return;
  const int x0 = bounds[2 * x], cnt = bounds[2 * x + 1];
  const int* k = coeffs + (long long)x * ksize;
  const unsigned char* p = src + n * img_stride + (long long)r * row_stride + (long long)x0 * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < cnt; ++j) {
    const int w = __ldg(k + j);
    a0 += (int)p[3 * j] * w;
    a1 += (int)p[3 * j + 1] * w;
    a2 += (int)p[3 * j + 2] * w;
  }
  unsigned char* o = dst + ((n * rows + r) * (long long)out_w + x) * 3;
  o[0] = clip8(a0);
  o[1] = clip8(a1);
  o[2] = clip8(a2);
}

// Vertical pass (or a plain crop when bounds == nullptr) + value table + CHW store.
template <typename OutT>
__global__ void __launch_bounds__(256)
resample_v_normalize_kernel(const unsigned char* __restrict__ src, long long img_stride, long long row_stride,
                            int x_offset, const int* __restrict__ bounds, const int* __restrict__ coeffs,
                            int ksize, int row_offset, int out_size, const float* __restrict__ lut,
                            int swap_rb, OutT* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const long long n = blockIdx.z;
  if (x >= out_size) return;
  const unsigned char* base = src + n * img_stride + (long long)(x + x_offset) * 3;
  unsigned char v0, v1, v2;
  if (bounds) {
    const int y0 = bounds[2 * y], cnt = bounds[2 * y + 1];
    const int* k = coeffs + (long long)y * ksize;
    const unsigned char* p = base + (long long)y0 * row_stride;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int w = __ldg(k + j);
      a0 += (int)p[0] * w;
      a1 += (int)p[1] * w;
      a2 += (int)p[2] * w;
      p += row_stride;
    }
    v0 = clip8(a0); v1 = clip8(a1); v2 = clip8(a2);
  } else {
    const unsigned char* p = base + (long long)(y + row_offset) * row_stride;
    v0 = p[0]; v1 = p[1]; v2 = p[2];
  }
  if (swap_rb) { const unsigned char t = v0; v0 = v2; v2 = t; }   // cv2 frames are BGR (cvtColor BGR2RGB)
  const long long plane = (long long)out_size * out_size;
  OutT* o = out + n * 3 * plane + (long long)y * out_size + x;
  o[0] = (OutT)__ldg(lut + v0);
  o[plane] = (OutT)__ldg(lut + 256 + v1);
  o[2 * plane] = (OutT)__ldg(lut + 512 + v2);
}

}  // namespace

extern "C" {

int scb_resample_h_u8(const uint8_t* src, int64_t n_img, int64_t img_stride, int64_t row_stride, int32_t rows,
                      const int32_t* bounds, const int32_t* coeffs, int32_t ksize, int32_t out_w, uint8_t* dst,
                      void* stream) {
  SCB_REQUIRE(src && bounds && coeffs && dst, SCB_ERR_INVALID_ARG, "scb_resample_h_u8: NULL buffer");
  SCB_REQUIRE(n_img >= 1 && n_img <= 65535 && rows >= 1 && rows <= 65535 && out_w >= 1 && ksize >= 1,
              SCB_ERR_INVALID_ARG, "scb_resample_h_u8: bad geometry");
  dim3 grid((unsigned)((out_w + 255) / 256), (unsigned)rows, (unsigned)n_img);
  resample_h_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, (long long)img_stride, (long long)row_stride, (int)rows, bounds, coeffs, (int)ksize, (int)out_w, dst);
  SCB_LAUNCH_OK("resample_h_kernel");
  return SCB_OK;
}

int scb_resample_v_normalize(const uint8_t* src, int64_t n_img, int64_t img_stride, int64_t row_stride,
                             int32_t x_offset, const int32_t* bounds, const int32_t* coeffs, int32_t ksize,
                             int32_t row_offset, int32_t out_size, const float* lut, int32_t swap_rb, void* out,
                             int32_t out_dtype, void* stream) {
  SCB_REQUIRE(src && lut && out, SCB_ERR_INVALID_ARG, "scb_resample_v_normalize: NULL buffer");
  SCB_REQUIRE((bounds == nullptr) == (coeffs == nullptr), SCB_ERR_INVALID_ARG,
              "scb_resample_v_normalize: bounds and coeffs go together");
  SCB_REQUIRE(n_img >= 1 && n_img <= 65535 && out_size >= 1 && out_size <= 65535 && x_offset >= 0 &&
                  row_offset >= 0,
              SCB_ERR_INVALID_ARG, "scb_resample_v_normalize: bad geometry");
  SCB_REQUIRE(out_dtype == SCB_F16 || out_dtype == SCB_F32, SCB_ERR_UNSUPPORTED,
              "scb_resample_v_normalize: out dtype %d", out_dtype);
  dim3 grid((unsigned)((out_size + 255) / 256), (unsigned)out_size, (unsigned)n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == SCB_F16)
    resample_v_normalize_kernel<__half><<<grid, 256, 0, st>>>(
        src, (long long)img_stride, (long long)row_stride, (int)x_offset, bounds, coeffs, (int)ksize,
        (int)row_offset, (int)out_size, lut, (int)swap_rb, reinterpret_cast<__half*>(out));
  else
    resample_v_normalize_kernel<float><<<grid, 256, 0, st>>>(
        src, (long long)img_stride, (long long)row_stride, (int)x_offset, bounds, coeffs, (int)ksize,
        (int)row_offset, (int)out_size, lut, (int)swap_rb, reinterpret_cast<float*>(out));
  SCB_LAUNCH_OK("resample_v_normalize_kernel");
  return SCB_OK;
}

}  // extern "C"
