// Exact Euclidean distance transform -> normalised signed-distance targets of the PAED loss on the GPU
// (SURVEY.md §8f rank 4).  Replaces compute_sdf (model/PAED/segmentation.py:6-34), which the reference's dataset runs
// per sample on the CPU through scipy.ndimage.distance_transform_edt:
//     sdf_ext = EDT(~mask) / max,   sdf_int = EDT(mask) / max        (each divided by its own maximum when > 0)
// EDT(a)[p] = distance from p to the nearest zero element of a (0 where a[p] == 0).  Exact integer arithmetic:
//   pass 1 (rows)    d1[y][x]  = distance along row y to the nearest zero of the row           (uint16, 0xFFFF = none)
//   pass 2 (columns) d2[y][x]  = min_y' ( d1[y'][x]^2 + (y - y')^2 )                            (int32, brute force)
//   value = (float) sqrt((double) d2)  — the same double-precision root that SciPy rounds to float32,
//   per-image maximum via atomicMax on the (non-negative) float bits, pass 3 divides.
// An image without any zero element has no nearest zero; SciPy then measures from a virtual zero at (row -1, col 0)
// and so does this kernel (bit-compatibility with the reference's targets for empty / full masks).
#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

constexpr int kSdfNone = 0xFFFF;

// one warp per (variant, image, row): lanes sweep the row in chunks of 32 with ballots
__global__ void __launch_bounds__(256)
sdf_rows_kernel(const float* __restrict__ mask, uint16_t* __restrict__ d1, int* __restrict__ has_zero, int B, int S) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * 8 + warp;
  if (row_id >= 2LL * B * S) return;
  const int v = int(row_id / ((long long)B * S));           // 0: EDT(~mask) (zeros = object pixels), 1: EDT(mask)
  const long long r = row_id - (long long)v * B * S;        // b * S + y
  const float* mrow = mask + r * S;
  uint16_t* out = d1 + row_id * S;
  // nearest zero to the left (inclusive), then to the right, both as positions
  int last = -1;   // position of the last zero seen while sweeping left -> right
  for (int x0 = 0; x0 < S; x0 += 32) {
    const int x = x0 + lane;
    const bool z = x < S && ((mrow[x] > 0.5f) == (v == 0));   // zero element of the transformed array
    const unsigned bal = __ballot_sync(0xffffffffu, z);
    const unsigned upto = bal & (0xffffffffu >> (31 - lane));
    const int left = upto ? x0 + 31 - __clz(upto) : last;
    if (x < S) out[x] = left >= 0 ? (uint16_t)(x - left) : (uint16_t)kSdfNone;
    if (bal) last = x0 + 31 - __clz(bal);
  }
  if (last >= 0 && lane == 0) atomicOr(&has_zero[v * B + int(r / S)], 1);
  int next = -1;   // sweeping right -> left
  for (int x0 = ((S - 1) / 32) * 32; x0 >= 0; x0 -= 32) {
    const int x = x0 + lane;
    const bool z = x < S && ((mrow[x] > 0.5f) == (v == 0));
    const unsigned bal = __ballot_sync(0xffffffffu, z);
    const unsigned from = bal & (0xffffffffu << lane);
    const int right = from ? x0 + __ffs(from) - 1 : next;
    if (x < S && right >= 0) {
      const int d = right - x;
      if (d < (int)out[x]) out[x] = (uint16_t)d;
    }
    if (bal) next = x0 + __ffs(bal) - 1;
  }
}

// block = (32-column strip, image, variant): the strip of d1 in smem, thread = column x row-lane
__global__ void __launch_bounds__(256)
sdf_cols_kernel(const uint16_t* __restrict__ d1, const int* __restrict__ has_zero, float* __restrict__ out_ext,
                float* __restrict__ out_int, unsigned* __restrict__ vmax, int B, int S) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint16_t s_d1[];   // [S][32]
  const int x0 = blockIdx.x * 32, b = blockIdx.y, v = blockIdx.z;
  const uint16_t* src = d1 + ((long long)v * B + b) * S * S;
  for (int i = threadIdx.x; i < S * 32; i += blockDim.x) {
    const int y = i >> 5, xl = i & 31;
    s_d1[i] = x0 + xl < S ? src[(long long)y * S + x0 + xl] : (uint16_t)kSdfNone;
  }
  __syncthreads();
  float* out = (v == 0 ? out_ext : out_int) + (long long)b * S * S;
  const bool any_zero = has_zero[v * B + b] != 0;
  const int xl = threadIdx.x & 31, x = x0 + xl;
  float local_max = 0.0f;
  for (int y = threadIdx.x >> 5; y < S; y += 8) {
    if (x >= S) continue;
    long long best;
    if (!any_zero) {
      best = (long long)(y + 1) * (y + 1) + (long long)x * x;   // SciPy's virtual zero at (-1, 0)
    } else {
      best = 1LL << 40;
      for (int yy = 0; yy < S; ++yy) {
        const int d = s_d1[yy * 32 + xl];
        if (d != kSdfNone) {
          const long long c = (long long)d * d + (long long)(y - yy) * (y - yy);
          best = c < best ? c : best;
        }
      }
    }
    const float val = (float)sqrt((double)best);
    out[(long long)y * S + x] = val;
    local_max = fmaxf(local_max, val);
  }
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0 && local_max > 0.0f) atomicMax(&vmax[v * B + b], __float_as_uint(local_max));
}

__global__ void sdf_normalise_kernel(float* __restrict__ out_ext, float* __restrict__ out_int,
                                     const unsigned* __restrict__ vmax, int B, int S) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i / n);
    const long long o = i - (long long)v * n;
    const float m = __uint_as_float(vmax[v * B + int(o / ((long long)S * S))]);
    float* p = (v == 0 ? out_ext : out_int) + o;
    if (m > 0.0f) *p = *p / m;
  }
}

}  // namespace vs

using namespace vs;

extern "C" int64_t vs_sdf_workspace_bytes(int32_t B, int32_t S) {
  if (B <= 0 || S <= 0) return 0;
  return (int64_t)2 * B * S * S * sizeof(uint16_t) + (int64_t)4 * B * sizeof(int32_t);
}

extern "C" int vs_sdf_targets(const float* mask, float* sdf_ext, float* sdf_int, void* workspace, int32_t B, int32_t S,
                              void* stream) {
  VS_CHECK_ARG(mask && sdf_ext && sdf_int && workspace, "vs_sdf_targets: null pointer");
  VS_CHECK_ARG(B > 0 && S > 0 && S < 0xFFFF && B <= 65535, "vs_sdf_targets: bad shape B=%d S=%d", B, S);
  VS_CHECK_ARG((size_t)S * 32 * sizeof(uint16_t) <= 160 * 1024, "vs_sdf_targets: S=%d too large", S);
  VS_CHECK_ARG((uintptr_t)workspace % 4 == 0, "vs_sdf_targets: workspace must be 4-byte aligned");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_sdf_targets: no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  int* flags = reinterpret_cast<int*>(workspace);                      // has_zero [2][B], vmax [2][B]
  unsigned* vmax = reinterpret_cast<unsigned*>(flags + 2 * B);
  uint16_t* d1 = reinterpret_cast<uint16_t*>(flags + 4 * B);
  VS_CHECK_CUDA(cudaMemsetAsync(flags, 0, (size_t)4 * B * sizeof(int), st));
  const long long rows = 2LL * B * S;
  launch_k(sdf_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), st, mask, d1, flags, B, S);
  const size_t smem = (size_t)S * 32 * sizeof(uint16_t);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(sdf_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  launch_k(sdf_cols_kernel, dim3(dim3((S + 31) / 32, B, 2)), dim3(256), (size_t)(smem), st, d1, flags, sdf_ext, sdf_int, vmax, B, S);
  long long nb = (2LL * B * S * S + 255) / 256;
  if (nb > (long long)nsm * 16) nb = (long long)nsm * 16;
  launch_k(sdf_normalise_kernel, dim3((unsigned)nb), dim3(256), (size_t)(0), st, sdf_ext, sdf_int, vmax, B, S);
  VS_CHECK_LAUNCH();
  return 0;
}
