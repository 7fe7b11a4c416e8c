// Worker-side image pre-processing on the device: the resize + ToTensor that the reference's
// inference path does on the CPU with PIL / torchvision before the model sees an image
//   model/CE/testViTModel.py:92-97    transforms.Compose([Resize((224, 224)), ToTensor()]) on a PIL image
//   backend/core/views.py:99-109     the Celery worker fetches the uploaded JPEG
// transforms.Resize on a PIL image is Image.resize(size, BILINEAR): Pillow's two-pass separable convolution
// (libImaging/Resample.c) whose triangle filter is widened by the down-scaling factor (antialiasing), evaluated in
// 22-bit fixed point with a rounded uint8 intermediate between the horizontal and the vertical pass.  The kernels
// below do exactly that integer arithmetic, so for the same decoded pixels the result is BIT-IDENTICAL to Pillow's;
// the coefficient tables (bounds + int32 weights per output coordinate) are built on the host as Resample.c's
// precompute_coeffs / normalize_coeffs_8bpc do (visiontransformer_b200/worker.py).
// Layout: planar uint8 [C, H, W] as nvJPEG / torchvision.io.decode_jpeg deliver it (the arithmetic is per channel).
#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

__device__ __forceinline__ int clip8(int v) {   // clip8_lookups[v >> PRECISION_BITS]
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: one block per (channel, input row); the row is staged in shared memory once
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, long long row_stride, long long plane_stride, int H, int W,
                  const int* __restrict__ bounds, const int* __restrict__ coef, int ksize, int Wout,
                  uint8_t* __restrict__ dst) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t s_row[];
  const int c = blockIdx.x / H, y = blockIdx.x - c * H;
  const uint8_t* in = src + (long long)c * plane_stride + (long long)y * row_stride;
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_row[i] = in[i];
  __syncthreads();
  uint8_t* out = dst + ((long long)c * H + y) * Wout;
  for (int xx = threadIdx.x; xx < Wout; xx += blockDim.x) {
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = coef + (long long)xx * ksize;
    int ss = 1 << (kPrecisionBits - 1);
    for (int i = 0; i < n; ++i) ss += (int)s_row[xmin + i] * k[i];
    out[xx] = (uint8_t)clip8(ss);
  }
}

// vertical pass: thread per output element (x fastest: coalesced reads of the intermediate rows); writes
// float(v) / scale (ToTensor: scale = 255) and/or the uint8 value
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ src, int C, int H, int Wout, const int* __restrict__ bounds,
                  const int* __restrict__ coef, int ksize, int Hout, float* __restrict__ dst_f32, float scale,
                  uint8_t* __restrict__ dst_u8) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)C * Hout * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wout);
    const long long t = i / Wout;
    const int yy = (int)(t % Hout), c = (int)(t / Hout);
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = coef + (long long)yy * ksize;
    const uint8_t* col = src + ((long long)c * H + ymin) * Wout + x;
    int ss = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < n; ++j) ss += (int)col[(long long)j * Wout] * k[j];
    const int v = clip8(ss);
    if (dst_f32 != nullptr) dst_f32[i] = (float)v / scale;   // a true division, as ToTensor's .div(255)
    if (dst_u8 != nullptr) dst_u8[i] = (uint8_t)v;
  }
}

// uint8 [C,H,W] -> fp32 [C,H,W] / scale (ToTensor when no resize is needed in either direction)
__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n, float scale) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i] / scale;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_resample_h_u8(const uint8_t* src, int64_t row_stride, int64_t plane_stride, int32_t C, int32_t H,
                                int32_t W, const int32_t* bounds, const int32_t* coef, int32_t ksize, int32_t Wout,
                                uint8_t* dst, void* stream) {
  VS_CHECK_ARG(src && bounds && coef && dst, "vs_resample_h_u8: null pointer");
  VS_CHECK_ARG(C > 0 && H > 0 && W > 0 && Wout > 0 && ksize > 0, "vs_resample_h_u8: bad shape");
  VS_CHECK_ARG(W <= 200 * 1024, "vs_resample_h_u8: rows wider than 204800 pixels are not supported");
  VS_CHECK_ARG(sm_count() > 0, "vs_resample_h_u8: no CUDA device");
  static int smem_set = 0;
  if (W > 48 * 1024 && W > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W));
    smem_set = W;
  }
  VS_CHECK_CUDA(launch_k(resample_h_kernel, dim3((unsigned)(C * H)), dim3(256), (size_t)W, (cudaStream_t)stream, src,
                         (long long)row_stride, (long long)plane_stride, H, W, bounds, coef, ksize, Wout, dst));
  return 0;
}

extern "C" int vs_resample_v_u8(const uint8_t* src, int32_t C, int32_t H, int32_t Wout, const int32_t* bounds,
                                const int32_t* coef, int32_t ksize, int32_t Hout, float* dst_f32, float scale,
                                uint8_t* dst_u8, void* stream) {
  VS_CHECK_ARG(src && bounds && coef && (dst_f32 || dst_u8), "vs_resample_v_u8: null pointer");
  VS_CHECK_ARG(C > 0 && H > 0 && Wout > 0 && Hout > 0 && ksize > 0, "vs_resample_v_u8: bad shape");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_resample_v_u8: no CUDA device");
  const long long total = (long long)C * Hout * Wout;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
  VS_CHECK_CUDA(launch_k(resample_v_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, (cudaStream_t)stream, src, C, H,
                         Wout, bounds, coef, ksize, Hout, dst_f32, scale, dst_u8));
  return 0;
}

extern "C" int vs_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float scale, void* stream) {
  VS_CHECK_ARG(src && dst && n > 0, "vs_u8_to_f32: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_u8_to_f32: no CUDA device");
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
  VS_CHECK_CUDA(launch_k(u8_to_f32_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, (cudaStream_t)stream, src, dst,
                         (long long)n, scale));
  return 0;
}
