// Gradient all-reduce of the data-parallel step over NVLink SHARP (NVLS), written for CO-RESIDENCY with the compute kernels.
//
// Problem (profiles/r02: tools/dp_timeline.py, 2 and 8 GPUs): all of the multi-GPU step overhead is the backward GEMMs
// that run while an NCCL all-reduce kernel is resident stretching 1.4x — NCCL's CTAs (hundreds of threads, ~96
// registers each) cannot share an SM with a persistent tcgen05 GEMM CTA (640 threads x 96 registers, 225 KB of shared
// memory), so they DISPLACE GEMM CTAs, whose tiles then run as a second wave.  Nothing else slows down.
//
// This kernel is sized to fit in what a GEMM CTA leaves free on its SM: 128 threads x <= 32 registers, no shared
// memory — one CTA per SM, resident NEXT TO the GEMM / attention CTAs instead of in their place.  The reduction itself
// happens in the NVSwitch: the gradient arena lives in symmetric memory mapped through a multicast address;
//   multimem.ld_reduce.add.v4.f32  reads one 16-byte vector summed over every GPU's copy (in-switch reduction),
//   multimem.st.v4.f32             writes the result to every GPU's copy,
// and each rank does this for its 1/world slice of a bucket (two-shot all-reduce, both shots in the switch).  Per GPU
// the NVLink traffic is 2/world of the bucket; the SMs only move addresses and 16-byte registers.
// Cross-GPU ordering (all ranks finished producing the bucket / all ranks finished writing it back) is the caller's
// job: visiontransformer_b200/dp.py brackets the launch with symmetric-memory barriers on its communication stream.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// mc: multicast address of the first element of this rank's slice; n4: 16-byte vectors in the slice
__global__ void __launch_bounds__(128, 16)
multimem_allreduce_kernel(float* __restrict__ mc, long long n4, float scale) {
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent vectors in flight per thread: the NVLink round trip is ~3 us, so the kernel's lifetime — which is
  // what delays GEMM CTAs that want to START on its SMs — is set by the bytes outstanding (148 CTAs x 128 threads x 4 x
  // 16 B = 1.2 MB)
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
      multimem_st(mc + 4 * (i + u * stride), v[u]);
    }
  }
  for (; i < n4; i += stride) {
    float4 a = multimem_ld_reduce_add(mc + 4 * i);
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    multimem_st(mc + 4 * i, a);
  }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_multimem_allreduce_f32(void* multicast_ptr, int64_t n, int32_t rank, int32_t world, float scale,
                                         void* stream) {
  VS_CHECK_ARG(multicast_ptr != nullptr && n > 0 && world > 0 && rank >= 0 && rank < world,
               "vs_multimem_allreduce_f32: bad arguments");
  VS_CHECK_ARG((uintptr_t)multicast_ptr % 16 == 0 && n % 4 == 0, "vs_multimem_allreduce_f32: 16-byte alignment required");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_multimem_allreduce_f32: no CUDA device");
  // this rank's slice: vectors [rank * per, min((rank + 1) * per, n4))
  const long long n4 = n / 4;
  const long long per = (n4 + world - 1) / world;
  const long long lo = (long long)rank * per;
  const long long hi = lo + per < n4 ? lo + per : n4;
  if (hi <= lo) return 0;
  // Tuning knobs (read once): VS_MM_CTAS = CTAs of the reduction kernel (default: one per SM), VS_MM_CARVEOUT=1 requests
  // the maximum shared-memory carve-out so that a tcgen05 GEMM CTA can JOIN an SM that hosts only this kernel (without
  // it, such a GEMM CTA waits for the SM to drain).  Measured at 2 GPUs (r02): joining is not the better deal — the two
  // kernels then contend for the memory system (GEMM 69 -> 113 us, reduction 79 -> 172 us) instead of the GEMM starting
  // 26 us late — so the default leaves the carve-out alone.
  static int n_ctas = 0, carve = -1;
  if (carve < 0) {
    const char* e = getenv("VS_MM_CARVEOUT");
    carve = (e && e[0] == '1') ? 1 : 0;
    const char* c = getenv("VS_MM_CTAS");
    n_ctas = c ? atoi(c) : 0;
    if (n_ctas <= 0 || n_ctas > nsm) n_ctas = nsm;
    if (carve)
      VS_CHECK_CUDA(cudaFuncSetAttribute(multimem_allreduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)cudaSharedmemCarveoutMaxShared));
  }
  float* base = reinterpret_cast<float*>(multicast_ptr) + 4 * lo;
  VS_CHECK_CUDA(launch_k(multimem_allreduce_kernel, dim3((unsigned)n_ctas), dim3(128), (size_t)0, (cudaStream_t)stream, base,
                         hi - lo, scale));
  return 0;
}
