// Glue kernels around the GEMMs (all HBM-bound, 128-bit vectorised where layouts allow):
//   bias-gradient column sums, patch extraction, CLS/position rows, embedding backward, segmentation-head
//   im2col / col2im, the 1x1 classifier conv (forward + backward with fused ReLU'), weight casts / packing.
// Reference anchors: TF:100-128,153-167 (embeddings); model/CE/classes.py:240-244,250-257 (seg head).
#include <stdlib.h>

#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] (+)= sum_m x[m, n]
// block = 256 threads = 32 column groups (8 bf16 = 16 B each) x 8 row lanes; grid.y splits the rows
// ------------------------------------------------------------------------------------------------
// FUSED_CAST: the first `nf` columns are still fp32 (xf: the dQ accumulator of the attention backward); they are rounded
// to bf16, written into x (so the QKV weight / data gradient GEMMs can read one bf16 matrix) and summed from the rounded
// values — replaces the separate cast pass over dQ (r01: cast_rows_kernel, 12 launches, 1.3 % of the step).
template <bool FUSED_CAST>
__global__ void __launch_bounds__(256)
colsum_kernel(__nv_bfloat16* __restrict__ x, long long ldx, int M, int N, float* __restrict__ out,
              const float* __restrict__ xf, long long ldf, int nf) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  auto add8 = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_bf16(w[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  };
  if (col < N) {
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    if (FUSED_CAST && col < nf) {
      int r = r0 + rl;
      for (; r + 8 < r1; r += 16) {   // two rows (4 x 16-byte loads) in flight per thread
        float4 a[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float* src = xf + (long long)(r + 8 * u) * ldf + col;
          a[u][0] = *reinterpret_cast<const float4*>(src);
          a[u][1] = *reinterpret_cast<const float4*>(src + 4);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint4 v = make_uint4(pack_bf16(a[u][0].x, a[u][0].y), pack_bf16(a[u][0].z, a[u][0].w),
                                     pack_bf16(a[u][1].x, a[u][1].y), pack_bf16(a[u][1].z, a[u][1].w));
          *reinterpret_cast<uint4*>(x + (long long)(r + 8 * u) * ldx + col) = v;
          add8(v);
        }
      }
      for (; r < r1; r += 8) {
        const float* src = xf + (long long)r * ldf + col;
        const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
        const uint4 v = make_uint4(pack_bf16(a0.x, a0.y), pack_bf16(a0.z, a0.w), pack_bf16(a1.x, a1.y), pack_bf16(a1.z, a1.w));
        *reinterpret_cast<uint4*>(x + (long long)r * ldx + col) = v;
        add8(v);
      }
    } else {
      // four independent 16-byte loads in flight per thread: the inputs are L2-resident (just written by the producing
      // GEMM), so the loop is bound by load latency, not bandwidth
      int r = r0 + rl;
      for (; r + 24 < r1; r += 32) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(x + (long long)(r + 8 * u) * ldx + col);
#pragma unroll
        for (int u = 0; u < 4; ++u) add8(v[u]);
      }
      for (; r < r1; r += 8) add8(*reinterpret_cast<const uint4*>(x + (long long)r * ldx + col));
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][cg * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns of this block
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][c];
  const int gc = blockIdx.x * 256 + c;
  if (gc < N) atomicAdd(&out[gc], s);
}

// ------------------------------------------------------------------------------------------------
// patchify: NCHW fp32 image -> bf16 [B*T, 3*P*P] with K = (c, ph, pw); 4 pixels per thread
// ------------------------------------------------------------------------------------------------
// One block per (image, patch row): consecutive threads walk a full image line (coalesced 16-byte reads), and the
// index arithmetic is 32-bit with one decode per element (the flat 64-bit grid-stride form spent 68 % of its issue slots
// on five 64-bit divisions per vector: 25 us for 58 MB, ncu r02 step table).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int S, int P) {
  pdl_wait();
  pdl_trigger();
  const int gp = S / P;
  const int K = 3 * P * P;
  const int S4 = S / 4;
  const int b = blockIdx.x / gp, ty = blockIdx.x - b * gp;
  const int n = 3 * P * S4;   // (channel, line within the patch row, 4-pixel group)
  const float* src = img + ((long long)b * 3 * S + (long long)ty * P) * S;
  __nv_bfloat16* dst = out + ((long long)b * (gp * gp + 1) + 1 + (long long)ty * gp) * K;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int line = i / S4, x = (i - line * S4) * 4;
    const int c = line / P, ph = line - c * P;
    const int tx = x / P, pw = x - tx * P;
    const float4 v = *reinterpret_cast<const float4*>(src + ((long long)c * S + ph) * S + x);
    *reinterpret_cast<uint2*>(dst + (long long)tx * K + c * P * P + ph * P + pw) =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                int B, int T1, int D) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  x[(long long)b * T1 * D + d] = cls[d] + pos[d];
}

// dpos[t, d] += sum_b dx[b, t, d]; dcls[d] += sum_b dx[b, 0, d]
__global__ void embed_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dcls, float* __restrict__ dpos,
                                 float* __restrict__ dbias, int B, int T1, int D) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over T1 * D/4
  const int D4 = D / 4;
  if (idx >= T1 * D4) return;
  float4 s = make_float4(0, 0, 0, 0);
  const float4* p = reinterpret_cast<const float4*>(dx) + idx;
  // eight independent loads in flight per thread (the grid is only T1 * D / 4 threads, about one block per SM: with a
  // plain dependent loop the kernel ran at 1.7 TB/s, r02 step table)
  int b = 0;
  for (; b + 8 <= B; b += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[(long long)(b + u) * T1 * D4];
#pragma unroll
    for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
  }
  for (; b < B; ++b) {
    const float4 v = p[(long long)b * T1 * D4];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4* dp = reinterpret_cast<float4*>(dpos) + idx;
  float4 o = *dp;
  o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
  *dp = o;
  if (idx < D4) {
    float4* dc = reinterpret_cast<float4*>(dcls) + idx;
    float4 c = *dc;
    c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
    *dc = c;
  } else {
    float* db = dbias + (idx % D4) * 4;
    atomicAdd(db + 0, s.x); atomicAdd(db + 1, s.y); atomicAdd(db + 2, s.z); atomicAdd(db + 3, s.w);
  }
}

// ------------------------------------------------------------------------------------------------
// seg head: im2col of the token grid (CLS dropped), 3x3 window, zero padding; K = (ky, kx, c)
// ------------------------------------------------------------------------------------------------
// One warp per (pixel, tap): the 2*D-byte source row and destination slice are both contiguous, so a lane moves 16-byte
// vectors at a fixed stride and the only index arithmetic is one pixel decode per block iteration.  (The first version
// was a flat grid-stride loop with five 64-bit divisions per 16-byte vector: 79 us = 2.2 TB/s for 173 MB written,
// 73 % issue-bound — ncu r02 step table.)
__global__ void __launch_bounds__(288)
head_im2col_kernel(const __nv_bfloat16* __restrict__ tok, __nv_bfloat16* __restrict__ col, int B, int g, int D) {
  pdl_wait();
  pdl_trigger();
  const int D8 = D / 8;
  const int kk = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 9 warps = the 9 taps
  const int dy = kk / 3 - 1, dx = kk % 3 - 1;
  const int npix = B * g * g;
  for (int pix = blockIdx.x; pix < npix; pix += gridDim.x) {
    const int b = pix / (g * g), t = pix - b * g * g;
    const int y = t / g, x = t - y * g;
    const int yy = y + dy, xx = x + dx;
    const bool in = yy >= 0 && yy < g && xx >= 0 && xx < g;
    const uint4* src = reinterpret_cast<const uint4*>(tok + ((long long)b * (g * g + 1) + 1 + yy * g + xx) * D);
    uint4* dst = reinterpret_cast<uint4*>(col + (long long)pix * (9LL * D) + (long long)kk * D);
    if (in) {
      for (int c8 = lane; c8 < D8; c8 += 32) dst[c8] = src[c8];
    } else {
      for (int c8 = lane; c8 < D8; c8 += 32) dst[c8] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// adjoint of im2col: dtok[b, 1 + y*g + x, c] = sum_{ky,kx} dcol[(b, y-ky+1, x-kx+1), (ky,kx,c)]; CLS row = 0
__global__ void head_col2im_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dtok, int B, int g, int D) {
  pdl_wait();
  pdl_trigger();
  const int D8 = D / 8;
  const int T1 = g * g + 1;
  const long long total = (long long)B * T1 * D8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(idx % D8);
    const long long r = idx / D8;
    const int t = int(r % T1), b = int(r / T1);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (t > 0) {
      const int y = (t - 1) / g, x = (t - 1) % g;
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) {
        const int yy = y - (kk / 3) + 1, xx = x - (kk % 3) + 1;
        if (yy >= 0 && yy < g && xx >= 0 && xx < g) {
          const uint4 v = *reinterpret_cast<const uint4*>(dcol + (((long long)b * g + yy) * g + xx) * (9LL * D) +
                                                          (long long)kk * D + c8 * 8);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = unpack_bf16(w[i]);
            acc[2 * i] += f.x;
            acc[2 * i + 1] += f.y;
          }
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(dtok + r * D + c8 * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1 classifier conv: one warp per pixel, weights in shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv1x1_fwd_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ logits, int B, int T, int F, int C) {
  pdl_wait();
  pdl_trigger();
  // weights as [C][F/256][8][32]: lane l holds features 8l .. 8l+7 of a 256-chunk (one 16-byte load), and reads
  // weight (i, l) at word i*32 + l — conflict-free (the plain [C][F] layout put 8 lanes on every bank: 60 us, ncu r01)
  extern __shared__ float s_w[];
  const int Fp = (F + 255) & ~255;   // class stride: whole 256-feature chunks
  for (int i = threadIdx.x; i < C * F; i += blockDim.x) {
    const int c = i / F, f = i - c * F;
    const int chunk = f >> 8, l = (f & 255) >> 3, k = f & 7;
    s_w[c * Fp + chunk * 256 + k * 32 + l] = w[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long npix = (long long)B * T;
  for (long long pix = (long long)blockIdx.x * 8 + warp; pix < npix; pix += (long long)gridDim.x * 8) {
    const int b = int(pix / T), t = int(pix - (long long)b * T);
    for (int f0 = 0; f0 < F; f0 += 256) {
      // each lane holds 8 consecutive features of this 256-chunk
      float x[8];
      const int f = f0 + lane * 8;
      if (f < F) {
        const uint4 v = *reinterpret_cast<const uint4*>(feat + pix * F + f);
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 q = unpack_bf16(ww[i]);
          x[2 * i] = q.x;
          x[2 * i + 1] = q.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.0f;
      }
      for (int c = 0; c < C; ++c) {
        float s = 0.0f;
        if (f < F) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s += x[i] * s_w[c * Fp + f0 + i * 32 + lane];
        }
        s = warp_sum(s);
        if (lane == 0) {
          float* o = logits + ((long long)b * C + c) * T + t;
          if (f0 == 0) *o = s + bias[c];
          else *o += s;
        }
      }
    }
  }
}

// thread f owns feature f: dfeat[pix,f] = relu'(feat) * sum_c dl[pix,c] w[c,f];  dw[c,f] += sum_pix dl[pix,c] feat[pix,f]
constexpr int kMaxClasses = 32;
__global__ void __launch_bounds__(256)
conv1x1_bwd_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ feat, const float* __restrict__ w,
                   __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ db, int B, int T,
                   int F, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float s_dl[kMaxClasses][64];   // class-major: one LDS.128 = the gradients of 4 pixels
  __shared__ __align__(16) __nv_bfloat16 s_x[64][256];   // feature rows of the current pixel group (F <= 256)
  const int f = threadIdx.x;
  const long long npix = (long long)B * T;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  float wc[kMaxClasses], acc[kMaxClasses];
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c) {
    wc[c] = (c < C && f < F) ? w[c * F + f] : 0.0f;
    acc[c] = 0.0f;
  }
  float dbacc = 0.0f;  // thread c (< C) accumulates db[c]
  for (long long base = p0; base < p1; base += 64) {
    const int n = int(min((long long)64, p1 - base));
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * C; i += blockDim.x) {
      const int c = i / 64, pi = i - c * 64;
      float v = 0.0f;   // pixels beyond n contribute zero (the loop below runs in groups of 4)
      if (pi < n) {
        const long long pix = base + pi;
        const int b = int(pix / T), t = int(pix - (long long)b * T);
        v = dlogits[((long long)b * C + c) * T + t];
      }
      s_dl[c][pi] = v;
    }
    // all feature rows of the group in flight at once (16-byte loads) instead of one dependent 2-byte load per pixel
    // and thread: the kernel was latency-bound (100 us for 6.4 MB, ncu r01)
    if ((F & 7) == 0) {
      const int f8 = F >> 3;
      for (int i = threadIdx.x; i < 64 * f8; i += blockDim.x) {
        const int pi = i / f8, q = i - pi * f8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (pi < n) v = *reinterpret_cast<const uint4*>(feat + (base + pi) * F + q * 8);
        *reinterpret_cast<uint4*>(&s_x[pi][q * 8]) = v;
      }
    } else {
      for (int i = threadIdx.x; i < 64 * F; i += blockDim.x) {
        const int pi = i / F, ff = i % F;
        s_x[pi][ff] = pi < n ? feat[(base + pi) * F + ff] : __float2bfloat16(0.0f);
      }
    }
    __syncthreads();
    if (f < C)
      for (int pi = 0; pi < n; ++pi) dbacc += s_dl[f][pi];
    if (f < F) {
      // four pixels per iteration: independent FMA chains, and the four gradients of a class arrive in one LDS.128
      for (int pi = 0; pi < n; pi += 4) {
        float x[4], d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = __bfloat162float(s_x[pi + u][f]);
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
          if (c < C) {
            const float4 g = *reinterpret_cast<const float4*>(&s_dl[c][pi]);
            d[0] = fmaf(g.x, wc[c], d[0]); d[1] = fmaf(g.y, wc[c], d[1]);
            d[2] = fmaf(g.z, wc[c], d[2]); d[3] = fmaf(g.w, wc[c], d[3]);
            acc[c] = fmaf(g.x, x[0], fmaf(g.y, x[1], fmaf(g.z, x[2], fmaf(g.w, x[3], acc[c]))));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (pi + u < n) dfeat[(base + pi + u) * F + f] = __float2bfloat16(x[u] > 0.0f ? d[u] : 0.0f);
      }
    }
  }
  // dw: per-block partials leave as 16-byte vector reductions (red.global.add.v4.f32): the [C][F] tile is transposed
  // through shared memory so a thread owns 4 consecutive features of one class.  (r01: 17 scalar atomics per thread
  // from 296 blocks = 1.3 M same-address atomics, 80 us; now C*F/4 vector reductions from one block per SM.)
  __syncthreads();
  float* s_out = reinterpret_cast<float*>(&s_x[0][0]);   // 64*256 bf16 = 32 KB >= C*F floats (C <= 32, F <= 256)
  if (f < F) {
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) s_out[c * F + f] = acc[c];
  }
  __syncthreads();
  if ((F & 3) == 0) {
    for (int i = threadIdx.x; i < (C * F) >> 2; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(s_out)[i];
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < C * F; i += blockDim.x) atomicAdd(&dw[i], s_out[i]);
  }
  if (f < C) atomicAdd(&db[f], dbacc);
}

// ------------------------------------------------------------------------------------------------
// casts / packing
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}

__global__ void cast_rows_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                 long long ldd, int M, int D) {
  pdl_wait();
  pdl_trigger();
  const int D4 = D / 4;
  const long long total = (long long)M * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D4;
    const int c = int(i - r * D4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// OIHW fp32 [O, I, 3, 3] -> bf16 [O, (ky, kx, I)]
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int I) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)O * I * 9;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = int(idx % I);
    const long long r = idx / I;
    const int kk = int(r % 9), o = int(r / 9);
    out[idx] = __float2bfloat16(w[((long long)o * I + i) * 9 + kk]);
  }
}
// grad fp32 [O, (ky,kx,I)] -> dw OIHW +=
__global__ void unpack_conv3x3_grad_kernel(const float* __restrict__ g, float* __restrict__ dw, int O, int I) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)O * I * 9;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = int(idx % 9);
    const long long r = idx / 9;
    const int i = int(r % I), o = int(r / I);
    dw[idx] += g[((long long)o * 9 + kk) * I + i];
  }
}

// in-place hidden-state dropout on an fp32 [M, D] matrix (+ optional bf16 copy); pair-hash scheme of common.cuh.
// Used for the embedding dropout (TF:126) forward and — same mask — on the gradient in backward.
__global__ void dropout_rows_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ x16, long long n4,
                                    const DropCfg drop) {
  pdl_wait();
  pdl_trigger();
  const uint32_t sd = drop_seed(drop);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    if (drop.thresh != 0u) {
      bool k0, k1, k2, k3;
      drop_keep2((uint32_t)(i * 4), sd, drop.thresh, k0, k1);
      drop_keep2((uint32_t)(i * 4 + 2), sd, drop.thresh, k2, k3);
      v.x = k0 ? v.x * drop.scale : 0.0f; v.y = k1 ? v.y * drop.scale : 0.0f;
      v.z = k2 ? v.z * drop.scale : 0.0f; v.w = k3 ? v.w * drop.scale : 0.0f;
      reinterpret_cast<float4*>(x)[i] = v;
    }
    if (x16) reinterpret_cast<uint2*>(x16)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// writes the keep mask (1/0 bytes) of a dropout site: scheme 0 = pair hash over a dense [n] index space (hidden
// states), scheme 1 = attention probabilities viewed as [n / row_len, row_len].  Test / debugging aid.
__global__ void dropout_mask_kernel(uint8_t* __restrict__ out, long long n, int scheme, int row_len,
                                    const DropCfg drop) {
  pdl_wait();
  pdl_trigger();
  const uint32_t sd = drop_seed(drop);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    bool keep;
    if (scheme == 0) {
      bool k0, k1;
      drop_keep2((uint32_t)(i & ~1LL), sd, drop.thresh, k0, k1);
      keep = (i & 1) ? k1 : k0;
    } else {
      // attention probabilities [B*H*N rows, row_len = N]: keys 4k .. 4k+3 of query q share hash(q * ceil(N/4) + k)
      const long long row = i / row_len;
      const int kv = int(i - row * row_len);
      bool k4[4];
      // rows are (batch*head, query) with row_len queries per (batch, head): per-(batch, head) seed, as the kernels
      const uint32_t bh = (uint32_t)(row / row_len), q = (uint32_t)(row - (long long)bh * row_len);
      drop_keep4(q * (uint32_t)((row_len + 3) >> 2) + (uint32_t)(kv >> 2), drop_hash(bh, sd), drop.thresh, k4);
      keep = k4[kv & 3];
    }
    out[i] = keep ? 1 : 0;
  }
}

static inline int grid_for(long long total, int block, int nsm) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)nsm * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vs

using namespace vs;

static int colsum_launch(void* x, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate, const float* xf,
                         int64_t ldf, int32_t nf, cudaStream_t st) {
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_colsum_bf16: no CUDA device");
  if (!accumulate) VS_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  const int gx = (N + 255) / 256;
  int gy = (nsm * 4 + gx - 1) / gx;
  if (gy > (M + 63) / 64) gy = (M + 63) / 64;
  if (gy < 1) gy = 1;
  if (xf != nullptr)
    launch_k(colsum_kernel<true>, dim3(gx, gy), dim3(256), (size_t)0, st, (__nv_bfloat16*)x, (long long)ldx, M, N, out, xf,
             (long long)ldf, nf);
  else
    launch_k(colsum_kernel<false>, dim3(gx, gy), dim3(256), (size_t)0, st, (__nv_bfloat16*)x, (long long)ldx, M, N, out,
             (const float*)nullptr, 0LL, 0);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_colsum_bf16(const void* x, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate,
                              void* stream) {
  VS_CHECK_ARG(x && out && M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0, "vs_colsum_bf16: bad arguments");
  return colsum_launch(const_cast<void*>(x), ldx, M, N, out, accumulate, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int vs_colsum_cast_bf16(void* x, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate,
                                   const float* x_f32, int64_t ldf, int32_t nf, void* stream) {
  VS_CHECK_ARG(x && out && x_f32 && M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0, "vs_colsum_cast_bf16: bad arguments");
  VS_CHECK_ARG(nf > 0 && nf <= N && nf % 8 == 0 && ldf % 4 == 0 && (uintptr_t)x_f32 % 16 == 0,
               "vs_colsum_cast_bf16: the fp32 column block must be a multiple of 8 columns, 16-byte aligned");
  return colsum_launch(x, ldx, M, N, out, accumulate, x_f32, ldf, nf, (cudaStream_t)stream);
}

extern "C" int vs_patchify(const float* img, void* out, int32_t B, int32_t S, int32_t P, void* stream) {
  VS_CHECK_ARG(img && out && B > 0 && S > 0 && P >= 4 && P % 4 == 0 && S % P == 0, "vs_patchify: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_patchify: no CUDA device");
  VS_CHECK_ARG((long long)B * (S / P) < (1LL << 31), "vs_patchify: too many patch rows");
  launch_k(patchify_kernel, dim3((unsigned)(B * (S / P))), dim3(256), (size_t)(0), (cudaStream_t)stream, img, (__nv_bfloat16*)out, B, S, P);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_cls_rows(const float* cls, const float* pos, float* x, int32_t B, int32_t T1, int32_t D, void* stream) {
  VS_CHECK_ARG(cls && pos && x && B > 0 && T1 > 0 && D > 0, "vs_cls_rows: bad arguments");
  VS_CHECK_ARG(sm_count() > 0, "vs_cls_rows: no CUDA device");
  launch_k(cls_rows_kernel, dim3((B * D + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, cls, pos, x, B, T1, D);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_embed_bwd(const float* dx, float* dcls, float* dpos, float* dbias, int32_t B, int32_t T1, int32_t D,
                            void* stream) {
  VS_CHECK_ARG(dx && dcls && dpos && dbias && B > 0 && T1 > 0 && D % 4 == 0, "vs_embed_bwd: bad arguments");
  VS_CHECK_ARG(sm_count() > 0, "vs_embed_bwd: no CUDA device");
  const int total = T1 * (D / 4);
  launch_k(embed_bwd_kernel, dim3((total + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, dx, dcls, dpos, dbias, B, T1, D);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_head_im2col(const void* tokens, void* col, int32_t B, int32_t g, int32_t D, void* stream) {
  VS_CHECK_ARG(tokens && col && B > 0 && g > 0 && D % 8 == 0, "vs_head_im2col: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_head_im2col: no CUDA device");
  const long long npix = (long long)B * g * g;
  VS_CHECK_ARG(npix < (1LL << 31), "vs_head_im2col: too many pixels");
  const long long grid = npix < (long long)nsm * 7 ? npix : (long long)nsm * 7;   // 7 blocks of 288 threads per SM
  launch_k(head_im2col_kernel, dim3((unsigned)grid), dim3(288), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)tokens, (__nv_bfloat16*)col, B, g, D);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_head_col2im(const void* dcol, float* dtokens, int32_t B, int32_t g, int32_t D, void* stream) {
  VS_CHECK_ARG(dcol && dtokens && B > 0 && g > 0 && D % 8 == 0, "vs_head_col2im: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_head_col2im: no CUDA device");
  const long long total = (long long)B * (g * g + 1) * (D / 8);
  launch_k(head_col2im_kernel, dim3(grid_for(total, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)dcol, dtokens, B, g, D);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_conv1x1_fwd(const void* feat, const float* w, const float* b, float* logits, int32_t B, int32_t g,
                              int32_t F, int32_t C, void* stream) {
  VS_CHECK_ARG(feat && w && b && logits && B > 0 && g > 0 && F % 8 == 0 && C > 0, "vs_conv1x1_fwd: bad arguments");
  const int Fp = (F + 255) & ~255;
  VS_CHECK_ARG((size_t)C * Fp * 4 <= 96 * 1024, "vs_conv1x1_fwd: C*F too large for shared memory");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_conv1x1_fwd: no CUDA device");
  const size_t smem = (size_t)C * Fp * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(conv1x1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const long long npix = (long long)B * g * g;
  // 8 blocks per SM: the warp-per-pixel loop is latency-bound (a 5-step shuffle reduction per class), so occupancy wins
  // over amortising the per-block weight staging (r02 s25: 2 blocks per SM measured 39.8 us against 32.5 us)
  int grid = (int)((npix + 7) / 8);
  if (grid > nsm * 8) grid = nsm * 8;
  launch_k(conv1x1_fwd_kernel, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, (const __nv_bfloat16*)feat, w, b, logits, B, g * g, F, C);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_conv1x1_bwd(const float* dlogits, const void* feat, const float* w, void* dfeat, float* dw, float* db,
                              int32_t B, int32_t g, int32_t F, int32_t C, void* stream) {
  VS_CHECK_ARG(dlogits && feat && w && dfeat && dw && db, "vs_conv1x1_bwd: null pointer");
  VS_CHECK_ARG(B > 0 && g > 0 && F > 0 && F <= 256 && C > 0 && C <= kMaxClasses,
               "vs_conv1x1_bwd: unsupported shape (F<=256, C<=%d)", kMaxClasses);
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_conv1x1_bwd: no CUDA device");
  const long long npix = (long long)B * g * g;
  VS_CHECK_ARG((uintptr_t)dw % 16 == 0, "vs_conv1x1_bwd: dw must be 16-byte aligned");
  // two blocks per SM (r02 s25, cold L2: 50.5 us with one, 41.2 us with two or three, 62 us with four — more blocks
  // shorten each block's serial pixel loop but multiply the same-address dW reductions); VS_C1B_BLOCKS_PER_SM overrides
  static int bwd_mult = -1;
  if (bwd_mult < 0) {
    const char* e = getenv("VS_C1B_BLOCKS_PER_SM");
    bwd_mult = e ? atoi(e) : 2;
    if (bwd_mult < 1 || bwd_mult > 4) bwd_mult = 2;
  }
  int grid = nsm * bwd_mult;
  if (grid > (npix + 63) / 64) grid = (int)((npix + 63) / 64);
  launch_k(conv1x1_bwd_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, dlogits, (const __nv_bfloat16*)feat, w, (__nv_bfloat16*)dfeat, dw, db, B, g * g, F, C);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VS_CHECK_ARG(src && dst && n > 0, "vs_cast_f32_bf16: bad arguments");
  VS_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), "vs_cast_f32_bf16: misaligned");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_cast_f32_bf16: no CUDA device");
  launch_k(cast_f32_bf16_kernel, dim3(grid_for(n / 4 + 1, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_cast_bf16_rows(const float* src, int64_t lds, void* dst, int64_t ldd, int32_t M, int32_t D,
                                 void* stream) {
  VS_CHECK_ARG(src && dst && M > 0 && D > 0 && D % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "vs_cast_bf16_rows: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_cast_bf16_rows: no CUDA device");
  launch_k(cast_rows_kernel, dim3(grid_for((long long)M * D / 4, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, src, lds, (__nv_bfloat16*)dst, ldd, M, D);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_pack_conv3x3(const float* w, void* out, int32_t O, int32_t I, void* stream) {
  VS_CHECK_ARG(w && out && O > 0 && I > 0, "vs_pack_conv3x3: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_pack_conv3x3: no CUDA device");
  launch_k(pack_conv3x3_kernel, dim3(grid_for((long long)O * I * 9, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, w, (__nv_bfloat16*)out, O, I);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_unpack_conv3x3_grad(const float* g, float* dw, int32_t O, int32_t I, void* stream) {
  VS_CHECK_ARG(g && dw && O > 0 && I > 0, "vs_unpack_conv3x3_grad: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_unpack_conv3x3_grad: no CUDA device");
  launch_k(unpack_conv3x3_grad_kernel, dim3(grid_for((long long)O * I * 9, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, g, dw, O, I);
  VS_CHECK_LAUNCH();
  return 0;
}

static int make_drop_cfg(DropCfg* dc, float p, const uint32_t* seed, uint32_t site, const char* who) {
  dc->thresh = 0u; dc->scale = 1.0f; dc->seed = nullptr; dc->site = 0u;
  if (p > 0.0f) {
    if (!(p < 1.0f) || seed == nullptr) { set_error("%s: need 0 < p < 1 and a device seed pointer", who); return -1; }
    dc->thresh = (uint32_t)(p * 65536.0f + 0.5f);
    dc->scale = 1.0f / (1.0f - (float)dc->thresh / 65536.0f);
    dc->seed = seed;
    dc->site = site;
  }
  return 0;
}

extern "C" int vs_dropout_rows(float* x, void* x_bf16, int64_t n, float dropout_p, const uint32_t* dropout_seed,
                               uint32_t dropout_site, void* stream) {
  VS_CHECK_ARG(x && n > 0 && n % 4 == 0 && n < (1LL << 32), "vs_dropout_rows: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_dropout_rows: no CUDA device");
  DropCfg dc;
  if (int rc = make_drop_cfg(&dc, dropout_p, dropout_seed, dropout_site, "vs_dropout_rows")) return rc;
  launch_k(dropout_rows_kernel, dim3(grid_for(n / 4, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (__nv_bfloat16*)x_bf16, n / 4, dc);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_dropout_mask(uint8_t* out, int64_t n, int32_t scheme, int32_t row_len, float dropout_p,
                               const uint32_t* dropout_seed, uint32_t dropout_site, void* stream) {
  VS_CHECK_ARG(out && n > 0 && n < (1LL << 32) && (scheme == 0 || (scheme == 1 && row_len > 0 && n % row_len == 0)),
               "vs_dropout_mask: bad arguments");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_dropout_mask: no CUDA device");
  DropCfg dc;
  if (int rc = make_drop_cfg(&dc, dropout_p, dropout_seed, dropout_site, "vs_dropout_mask")) return rc;
  launch_k(dropout_mask_kernel, dim3(grid_for(n, 256, nsm)), dim3(256), (size_t)(0), (cudaStream_t)stream, out, n, scheme, row_len, dc);
  VS_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Fused Adam / AdamW over the flat parameter arena (model/CE/classes.py:296-297 Adam lr 1e-5;
// model/PAED/classes.py:486-487 Adam, :536-548 AdamW).  One pass: reads p, g, m, v; writes p, m, v, the bf16 weight
// shadow used by the tensor cores, and (optionally) zeroes g — replacing torch's multi-tensor Adam + the separate
// cast pass + the gradient memset.  lr / step live in device memory so a captured CUDA graph sees schedulers.
// ------------------------------------------------------------------------------------------------
namespace vs {
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ shadow, long long n4, const float* __restrict__ lr_ptr,
            const int* __restrict__ step_ptr, float beta1, float beta2, float eps, float wd, int decoupled,
            float grad_scale, int zero_grad, long long skip_begin4, long long skip_end4) {
  pdl_wait();
  pdl_trigger();
  const float lr = lr_ptr[0];
  const float t = (float)step_ptr[0];
  const float bc1 = 1.0f - exp2f(t * log2f(beta1));
  const float bc2 = 1.0f - exp2f(t * log2f(beta2));
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = decoupled ? 1.0f - lr * wd : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    if (i < skip_begin4 || i >= skip_end4) {
      float4 gv = reinterpret_cast<float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gg = gp[k] * grad_scale;
        if (!decoupled) gg = fmaf(wd, pp[k], gg);
        const float mm = fmaf(beta1, mp[k], (1.0f - beta1) * gg);
        const float v2 = fmaf(beta2, vp[k], (1.0f - beta2) * gg * gg);
        mp[k] = mm;
        vp[k] = v2;
        const float denom = sqrtf(v2) * inv_sqrt_bc2 + eps;
        pp[k] = pp[k] * decay - step_size * (mm / denom);
      }
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pv.x, pv.y), pack_bf16(pv.z, pv.w));
  }
}
}  // namespace vs

extern "C" int vs_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                            const float* lr_dev, const int32_t* step_dev, float beta1, float beta2, float eps,
                            float weight_decay, int32_t decoupled, float grad_scale, int32_t zero_grad,
                            int64_t skip_begin, int64_t skip_end, void* stream) {
  VS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && shadow_bf16 && lr_dev && step_dev, "vs_adam_step: null pointer");
  VS_CHECK_ARG(n > 0 && n % 4 == 0 && skip_begin % 4 == 0 && skip_end % 4 == 0 && skip_begin <= skip_end,
               "vs_adam_step: n and the skip range must be multiples of 4");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_adam_step: no CUDA device");
  launch_k(adam_kernel, dim3(nsm * 8), dim3(256), (size_t)(0), (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n / 4, lr_dev, step_dev, beta1, beta2, eps, weight_decay, decoupled, grad_scale, zero_grad, skip_begin / 4, skip_end / 4);
  VS_CHECK_LAUNCH();
  return 0;
}
