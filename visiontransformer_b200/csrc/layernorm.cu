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
// Occupancy matters more than anything else here: the kernel is a chain of dependent global loads, two warp
// reductions and stores per row.  gamma lives in shared memory (not 4*NV registers) and the register budget is capped
// so that two 8-warp blocks fit per SM (first version: 254 registers, one block per SM, ~2 TB/s).
template <int NV, bool DY_F32>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dx_in, int M,
              float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, const DropCfg drop) {
  constexpr int D = NV * 128;
  __shared__ float s_dg[D];
  __shared__ float s_db[D];
  __shared__ __align__(16) float s_gamma[D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_dg[i] = 0.0f; s_db[i] = 0.0f; s_gamma[i] = gamma[i]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float4* g = reinterpret_cast<const float4*>(s_gamma) + lane;   // g[32 * i]
  float4 acc_g[NV], acc_b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc_g[i] = make_float4(0, 0, 0, 0);
    acc_b[i] = make_float4(0, 0, 0, 0);
  }
  for (int row = blockIdx.x * wpb + warp; row < M; row += gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    float4 dy[NV], xh[NV];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (DY_F32) {
        dy[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * D)[lane + 32 * i];
      } else {
        const uint2 pk =
            reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + (size_t)row * D)[lane + 32 * i];
        const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
        dy[i] = make_float4(a.x, a.y, b.x, b.y);
      }
      const float4 xv = reinterpret_cast<const float4*>(x + (size_t)row * D)[lane + 32 * i];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      const float4 gm = g[32 * i];
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
      const float4 gm = g[32 * i];
      float4 o;
      o.x = rs * (dy[i].x * gm.x - m1 - xh[i].x * m2);
      o.y = rs * (dy[i].y * gm.y - m1 - xh[i].y * m2);
      o.z = rs * (dy[i].z * gm.z - m1 - xh[i].z * m2);
      o.w = rs * (dy[i].w * gm.w - m1 - xh[i].w * m2);
      if (dx_in) {
        const float4 r = reinterpret_cast<const float4*>(dx_in + (size_t)row * D)[lane + 32 * i];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx_out + (size_t)row * D)[lane + 32 * i] = o;
      if (dx_bf16) {
        if (drop.thresh != 0u) {
          // the bf16 copy feeds the dgrad/wgrad of the GEMM whose output was dropped out in forward: d(acc) = m*dy/(1-p)
          const uint32_t e = (uint32_t)row * (uint32_t)D + (uint32_t)(lane + 32 * i) * 4u;
          const uint32_t sd = drop_seed(drop);
          bool k0, k1, k2, k3;
          drop_keep2(e, sd, drop.thresh, k0, k1);
          drop_keep2(e + 2, sd, drop.thresh, k2, k3);
          o.x = k0 ? o.x * drop.scale : 0.0f; o.y = k1 ? o.y * drop.scale : 0.0f;
          o.z = k2 ? o.z * drop.scale : 0.0f; o.w = k3 ? o.w * drop.scale : 0.0f;
        }
        reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[lane + 32 * i] =
            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    atomicAdd(&s_dg[c + 0], acc_g[i].x); atomicAdd(&s_dg[c + 1], acc_g[i].y);
    atomicAdd(&s_dg[c + 2], acc_g[i].z); atomicAdd(&s_dg[c + 3], acc_g[i].w);
    atomicAdd(&s_db[c + 0], acc_b[i].x); atomicAdd(&s_db[c + 1], acc_b[i].y);
    atomicAdd(&s_db[c + 2], acc_b[i].z); atomicAdd(&s_db[c + 3], acc_b[i].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_dg[i]);
    atomicAdd(&dbeta[i], s_db[i]);
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
#define VS_LN_CASE(NV) case NV: ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, yb, y_f32, mean, rstd); break;
    VS_LN_CASE(1) VS_LN_CASE(2) VS_LN_CASE(3) VS_LN_CASE(4) VS_LN_CASE(5) VS_LN_CASE(6) VS_LN_CASE(7) VS_LN_CASE(8)
#undef VS_LN_CASE
  }
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, const float* gamma,
                                const float* mean, const float* rstd, const float* dx_in, int32_t M, int32_t D,
                                float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, float dropout_p,
                                const uint32_t* dropout_seed, uint32_t dropout_site, void* stream) {
  VS_CHECK_ARG(dy && x && gamma && mean && rstd && dx_out && dgamma && dbeta, "vs_layernorm_bwd: null pointer");
  VS_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "vs_layernorm_bwd: D=%d must be a multiple of 128, <= 1024", D);
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_layernorm_bwd: no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = nsm * 2;   // two resident blocks per SM, each a persistent loop over rows
  if (grid > (M + 7) / 8) grid = (M + 7) / 8;
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
#define VS_LN_CASE(NV)                                                                                              \
  case NV:                                                                                                          \
    if (dy_is_f32) ln_bwd_kernel<NV, true><<<grid, 256, 0, st>>>(dy, x, gamma, mean, rstd, dx_in, M, dx_out, db, dgamma, dbeta, dc); \
    else ln_bwd_kernel<NV, false><<<grid, 256, 0, st>>>(dy, x, gamma, mean, rstd, dx_in, M, dx_out, db, dgamma, dbeta, dc);          \
    break;
    VS_LN_CASE(1) VS_LN_CASE(2) VS_LN_CASE(3) VS_LN_CASE(4) VS_LN_CASE(5) VS_LN_CASE(6) VS_LN_CASE(7) VS_LN_CASE(8)
#undef VS_LN_CASE
  }
  VS_CHECK_LAUNCH();
  return 0;
}
