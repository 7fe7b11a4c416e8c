// Warp-shuffle LayerNorm forward / backward over the fp32 residual stream (HBM-bound).
// Replaces nn.LayerNorm(D, eps=1e-12) at TF:325-326,333,340,416,455 and its autograd.
// One warp per token row; every lane keeps its D/32 elements in registers (128-bit loads), statistics in fp32
// with a two-pass (mean, then centred variance) formulation — never E[x^2]-E[x]^2.
#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

template <int NV>  // NV = D / 128 float4 chunks per lane
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              int M, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ mean_out,
              float* __restrict__ rstd_out) {
  pdl_wait();
  pdl_trigger();
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[NV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bb.x;
    o.y = (v[i].y - mean) * rstd * g.y + bb.y;
    o.z = (v[i].z - mean) * rstd * g.z + bb.z;
    o.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (y_bf16) {
      uint2 pk = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      reinterpret_cast<uint2*>(y_bf16 + (size_t)row * D)[lane + 32 * i] = pk;
    }
    if (y_f32) reinterpret_cast<float4*>(y_f32 + (size_t)row * D)[lane + 32 * i] = o;
  }
}

// dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma;  dgamma += sum dy*xhat; dbeta += sum dy
//
// HBM-bound (16 B per element) and, as a plain load->reduce->load->store loop, latency-bound: with the row held in
// registers only ~50 KB per SM were ever in flight (ncu r01: 15 warps stalled on long-scoreboard per issue, 3.0 TB/s
// cold).  Here the operands land in shared memory instead: one producer thread streams groups of 8 consecutive rows
// of dy, x and dx_in (contiguous in memory) with cp.async.bulk into a 2-3 stage mbarrier ring (up to ~180 KB in
// flight per SM); 8 consumer warps take one row each from smem, so their only global accesses are the stores.
// dgamma / dbeta — and optionally the column sums of the bf16 output, i.e. the bias gradient of the Linear layer that
// consumes it (saves the separate vs_colsum_bf16 pass) — stay in registers across rows, are combined through smem
// slabs (no smem atomics) and leave as one global atomic per column per block.
constexpr int kLnbRows = 8;                    // rows per stage = consumer warps
constexpr int kLnbThreads = (kLnbRows + 1) * 32;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NV, bool DY_F32>
__global__ void __launch_bounds__(kLnbThreads, 1)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dx_in, int M,
              float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dbias, const DropCfg drop, const int stages) {
  pdl_wait();
  pdl_trigger();
  constexpr int D = NV * 128;
  constexpr int kDyRow = D * (DY_F32 ? 4 : 2);
  constexpr int kStageBytes = kLnbRows * (2 * D * 4 + kDyRow);   // x | dx_in | dy
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [stages]
  uint64_t* empty = full + 4;                                    // [stages]
  float* s_gamma = reinterpret_cast<float*>(smem + 128);
  uint8_t* ring = smem + 128 + D * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (M + kLnbRows - 1) / kLnbRows;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kLnbRows); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_gamma[i] = gamma[i];
  __syncthreads();

  float4 acc_g[NV], acc_b[NV], acc_c[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc_g[i] = make_float4(0, 0, 0, 0);
    acc_b[i] = make_float4(0, 0, 0, 0);
    acc_c[i] = make_float4(0, 0, 0, 0);
  }

  if (warp == kLnbRows) {
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      int k = 0;
      for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++k) {
        const int s = k % stages;
        if (k >= stages) mbar_wait(&empty[s], ((k / stages) - 1) & 1);
        const int row0 = g * kLnbRows;
        const uint32_t nrows = (uint32_t)min(kLnbRows, M - row0);
        uint8_t* st = ring + (size_t)s * kStageBytes;
        mbar_expect_tx(&full[s], nrows * (uint32_t)((dx_in ? 2 : 1) * D * 4 + kDyRow));
        bulk_load(st, x + (size_t)row0 * D, nrows * D * 4, &full[s]);
        if (dx_in) bulk_load(st + kLnbRows * D * 4, dx_in + (size_t)row0 * D, nrows * D * 4, &full[s]);
        bulk_load(st + 2 * kLnbRows * D * 4, reinterpret_cast<const uint8_t*>(dy_) + (size_t)row0 * kDyRow,
                  nrows * kDyRow, &full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------ consumers: warp w owns row w of every group
    const float4* g4 = reinterpret_cast<const float4*>(s_gamma) + lane;   // g4[32 * i]
    const uint32_t dseed = drop.thresh != 0u ? drop_seed(drop) : 0u;
    int k = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++k) {
      const int s = k % stages;
      const int row = g * kLnbRows + warp;
      const bool row_ok = row < M;
      float mu = 0.0f, rs = 0.0f;
      if (row_ok) { mu = mean[row]; rs = rstd[row]; }
      mbar_wait(&full[s], (k / stages) & 1);
      if (row_ok) {
        const uint8_t* st = ring + (size_t)s * kStageBytes;
        const float4* sx = reinterpret_cast<const float4*>(st + warp * D * 4) + lane;
        const float4* sr = reinterpret_cast<const float4*>(st + (kLnbRows + warp) * D * 4) + lane;
        const uint8_t* sdy = st + 2 * kLnbRows * D * 4 + warp * kDyRow;
        float4 dy[NV], xh[NV];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (DY_F32) {
            dy[i] = reinterpret_cast<const float4*>(sdy)[lane + 32 * i];
          } else {
            const uint2 pk = reinterpret_cast<const uint2*>(sdy)[lane + 32 * i];
            const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
            dy[i] = make_float4(a.x, a.y, b.x, b.y);
          }
          const float4 xv = sx[32 * i];
          xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
          const float4 gm = g4[32 * i];
          const float gx = dy[i].x * gm.x, gy = dy[i].y * gm.y, gz = dy[i].z * gm.z, gw = dy[i].w * gm.w;
          s1 += (gx + gy) + (gz + gw);
          s2 += (gx * xh[i].x + gy * xh[i].y) + (gz * xh[i].z + gw * xh[i].w);
          acc_g[i].x += dy[i].x * xh[i].x; acc_g[i].y += dy[i].y * xh[i].y;
          acc_g[i].z += dy[i].z * xh[i].z; acc_g[i].w += dy[i].w * xh[i].w;
          acc_b[i].x += dy[i].x; acc_b[i].y += dy[i].y; acc_b[i].z += dy[i].z; acc_b[i].w += dy[i].w;
        }
        const float m1 = warp_sum(s1) * (1.0f / D);
        const float m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 gm = g4[32 * i];
          float4 o;
          o.x = rs * (dy[i].x * gm.x - m1 - xh[i].x * m2);
          o.y = rs * (dy[i].y * gm.y - m1 - xh[i].y * m2);
          o.z = rs * (dy[i].z * gm.z - m1 - xh[i].z * m2);
          o.w = rs * (dy[i].w * gm.w - m1 - xh[i].w * m2);
          if (dx_in) {
            const float4 r = sr[32 * i];
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
          reinterpret_cast<float4*>(dx_out + (size_t)row * D)[lane + 32 * i] = o;
          if (dx_bf16) {
            if (drop.thresh != 0u) {
              // the bf16 copy feeds the dgrad/wgrad of the GEMM whose output was dropped out in forward:
              // d(acc) = m * dy / (1-p)
              const uint32_t e = (uint32_t)row * (uint32_t)D + (uint32_t)(lane + 32 * i) * 4u;
              bool k0, k1, k2, k3;
              drop_keep2(e, dseed, drop.thresh, k0, k1);
              drop_keep2(e + 2, dseed, drop.thresh, k2, k3);
              o.x = k0 ? o.x * drop.scale : 0.0f; o.y = k1 ? o.y * drop.scale : 0.0f;
              o.z = k2 ? o.z * drop.scale : 0.0f; o.w = k3 ? o.w * drop.scale : 0.0f;
            }
            reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[lane + 32 * i] =
                make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            acc_c[i].x += o.x; acc_c[i].y += o.y; acc_c[i].z += o.z; acc_c[i].w += o.w;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  // ---- block reduction of the per-warp column accumulators through slabs laid over the (now idle) ring
  __syncthreads();
  const int nvec = dbias ? 3 : 2;
  float* slab = reinterpret_cast<float*>(ring);   // [warp][vec][D]
  if (warp < kLnbRows) {
    float4* sl = reinterpret_cast<float4*>(slab + (size_t)warp * nvec * D) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sl[32 * i] = acc_g[i];
      sl[D / 4 + 32 * i] = acc_b[i];
      if (dbias) sl[2 * (D / 4) + 32 * i] = acc_c[i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nvec * D; idx += blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kLnbRows; ++w) t += slab[(size_t)w * nvec * D + idx];
    const int vec = idx / D, c = idx - vec * D;
    atomicAdd((vec == 0 ? dgamma : (vec == 1 ? dbeta : dbias)) + c, t);
  }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t M, int32_t D,
                                void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream) {
  VS_CHECK_ARG(x && gamma && beta && (y_bf16 || y_f32), "vs_layernorm_fwd: null pointer");
  VS_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "vs_layernorm_fwd: D=%d must be a multiple of 128, <= 1024", D);
  VS_CHECK_ARG(sm_count() > 0, "vs_layernorm_fwd: no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (M + 7) / 8;
  __nv_bfloat16* yb = (__nv_bfloat16*)y_bf16;
  switch (D / 128) {
#define VS_LN_CASE(NV) case NV: launch_k(ln_fwd_kernel<NV>, dim3(grid), dim3(256), (size_t)(0), st, x, gamma, beta, eps, M, yb, y_f32, mean, rstd); break;
    VS_LN_CASE(1) VS_LN_CASE(2) VS_LN_CASE(3) VS_LN_CASE(4) VS_LN_CASE(5) VS_LN_CASE(6) VS_LN_CASE(7) VS_LN_CASE(8)
#undef VS_LN_CASE
  }
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, const float* gamma,
                                const float* mean, const float* rstd, const float* dx_in, int32_t M, int32_t D,
                                float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, float* dbias_colsum,
                                float dropout_p, const uint32_t* dropout_seed, uint32_t dropout_site, void* stream) {
  VS_CHECK_ARG(dy && x && gamma && mean && rstd && dx_out && dgamma && dbeta, "vs_layernorm_bwd: null pointer");
  VS_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "vs_layernorm_bwd: D=%d must be a multiple of 128, <= 1024", D);
  VS_CHECK_ARG(dbias_colsum == nullptr || dx_bf16 != nullptr, "vs_layernorm_bwd: dbias_colsum needs the bf16 output");
  VS_CHECK_ARG(((uintptr_t)dy % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dx_in % 16 == 0) &&
                   ((uintptr_t)dx_out % 16 == 0) && ((uintptr_t)dx_bf16 % 8 == 0),
               "vs_layernorm_bwd: operands must be 16-byte aligned");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_layernorm_bwd: no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  const int stage_bytes = kLnbRows * (2 * D * 4 + D * (dy_is_f32 ? 4 : 2));
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 3) stages = 3;
  const int smem_bytes = 128 + 128 + D * 4 + stages * stage_bytes;   // alignment slack | barriers | gamma | ring
  const int ngroups = (M + kLnbRows - 1) / kLnbRows;
  const int grid = ngroups < nsm ? ngroups : nsm;   // one persistent block per SM
  __nv_bfloat16* db = (__nv_bfloat16*)dx_bf16;
  DropCfg dc{0u, 1.0f, nullptr, 0u};
  if (dropout_p > 0.0f) {
    VS_CHECK_ARG(dropout_p < 1.0f && dropout_seed != nullptr, "vs_layernorm_bwd: bad dropout arguments");
    dc.thresh = (uint32_t)(dropout_p * 65536.0f + 0.5f);
    dc.scale = 1.0f / (1.0f - (float)dc.thresh / 65536.0f);
    dc.seed = dropout_seed;
    dc.site = dropout_site;
  }
  switch (D / 128) {
#define VS_LN_LAUNCH(NV, F32)                                                                                        \
  {                                                                                                                  \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      VS_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                         128 + 128 + 1024 * 4 + 200 * 1024));                                        \
      attr = true;                                                                                                   \
    }                                                                                                                \
    launch_k(ln_bwd_kernel<NV, F32>, dim3(grid), dim3(kLnbThreads), (size_t)(smem_bytes), st, dy, x, gamma, mean, rstd, dx_in, M, dx_out, db, dgamma, dbeta, dbias_colsum, dc, stages);         \
  }
#define VS_LN_CASE(NV)                                                                                               \
  case NV:                                                                                                           \
    if (dy_is_f32) VS_LN_LAUNCH(NV, true) else VS_LN_LAUNCH(NV, false)                                               \
    break;
    VS_LN_CASE(1) VS_LN_CASE(2) VS_LN_CASE(3) VS_LN_CASE(4) VS_LN_CASE(5) VS_LN_CASE(6) VS_LN_CASE(7) VS_LN_CASE(8)
#undef VS_LN_CASE
#undef VS_LN_LAUNCH
  }
  VS_CHECK_LAUNCH();
  return 0;
}
