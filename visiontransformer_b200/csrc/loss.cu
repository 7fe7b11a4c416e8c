// Segmentation-head output stage: bilinear upsample (align_corners=False) fused with the losses.
//   model/CE/classes.py:260        F.interpolate(out, size, mode='bilinear', align_corners=False)
//   model/CE/classes.py:276-285    nn.CrossEntropyLoss on the upsampled logits
//   model/PAED/classes.py:608-701  PAEDTrainer: sigmoid, BCE, Dice, Sobel edge x SDF terms
//   model/PAED/classes.py:336-369  paed_loss_multiclass_soft on softmax probabilities
//
// The full-resolution logits [B,C,S,S] are only materialised for the inference contract; training kernels
// interpolate on the fly from the low-resolution grid and reduce gradients back onto it.
//
// "Region" decomposition used by every adjoint: with P = S/g, output rows [ry*P - P/2, ry*P + P/2) (clipped) for
// ry = 0..g all interpolate between the same two grid rows, likewise for columns.  One warp (or block) owns one
// region, keeps per-lane partial sums and issues 4 atomics per class at the end.
#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

// PyTorch's area_pixel_compute_source_index (align_corners=False) + the lambda computation of upsample_bilinear2d
__device__ __forceinline__ void bil_coord(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  src = src < 0.0f ? 0.0f : src;
  i0 = (int)src;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

__device__ __forceinline__ float bil_sample(const float* plane, int g, int y0, int y1, float ly0, float ly1, int x0,
                                            int x1, float lx0, float lx1) {
  return ly0 * (lx0 * plane[y0 * g + x0] + lx1 * plane[y0 * g + x1]) +
         ly1 * (lx0 * plane[y1 * g + x0] + lx1 * plane[y1 * g + x1]);
}

struct Region {
  int b, y_lo, y_hi, x_lo, x_hi;
};
__device__ __forceinline__ Region region_of(long long rid, int g, int S) {
  const int P = S / g, g1 = g + 1;
  Region r;
  const int rx = int(rid % g1);
  const long long t = rid / g1;
  const int ry = int(t % g1);
  r.b = int(t / g1);
  r.y_lo = max(0, ry * P - P / 2);
  r.y_hi = min(S, ry * P + P / 2);
  r.x_lo = max(0, rx * P - P / 2);
  r.x_hi = min(S, rx * P + P / 2);
  return r;
}

// ================================================================================================
// plain upsample (inference contract), its adjoint, and the fused argmax
// ================================================================================================
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float* __restrict__ low, float* __restrict__ full, int g, int S, int chunks) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_plane[];
  const long long plane = blockIdx.x / chunks;
  const int chunk = blockIdx.x % chunks;
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) s_plane[i] = low[plane * g * g + i];
  __syncthreads();
  const float scale = (float)g / (float)S;
  const int S4 = S / 4;
  const int rows_per = (S + chunks - 1) / chunks;
  const int y_begin = chunk * rows_per, y_end = min(S, y_begin + rows_per);
  float* out = full + plane * S * S;
  // A thread keeps ONE group of 4 columns and walks down the rows: the column coordinates / weights are computed once,
  // and the horizontal interpolation of the two source rows (top, bot) is refreshed only when the source row pair
  // changes (every P rows) — a pixel then costs one FMUL + one FFMA.  The first version recomputed four column
  // coordinates and four 4-tap samples per vector: ~25 instructions per pixel, 83 us for 218 MB (2.6 TB/s,
  // instruction-bound; tools/head_bench.py r02 s23).  Same association order as bil_sample.
  const int xl = S4 < (int)blockDim.x ? S4 : (int)blockDim.x;   // threads along a row
  const int ny = blockDim.x / xl;                                // rows in flight per block
  const int tx = threadIdx.x % xl, tyl = threadIdx.x / xl;
  if (tyl >= ny) return;
  for (int xg = tx; xg < S4; xg += xl) {
    const int x4 = xg * 4;
    int x0[4], x1[4];
    float lx0[4], lx1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bil_coord(x4 + k, scale, g, x0[k], x1[k], lx0[k], lx1[k]);
    int cy0 = -1, cy1 = -1;
    float top[4], bot[4];
    for (int y = y_begin + tyl; y < y_end; y += ny) {
      int y0, y1;
      float ly0, ly1;
      bil_coord(y, scale, g, y0, y1, ly0, ly1);
      if (y0 != cy0 || y1 != cy1) {
        cy0 = y0; cy1 = y1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          top[k] = lx0[k] * s_plane[y0 * g + x0[k]] + lx1[k] * s_plane[y0 * g + x1[k]];
          bot[k] = lx0[k] * s_plane[y1 * g + x0[k]] + lx1[k] * s_plane[y1 * g + x1[k]];
        }
      }
      __stcs(reinterpret_cast<float4*>(out + (long long)y * S + x4),
             make_float4(ly0 * top[0] + ly1 * bot[0], ly0 * top[1] + ly1 * bot[1], ly0 * top[2] + ly1 * bot[2],
                         ly0 * top[3] + ly1 * bot[3]));
    }
  }
}

// one warp per (plane, region): dlow[plane, cells] += sum over the region's pixels
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ dfull, float* __restrict__ dlow, long long planes, int g, int S) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nreg = planes * (g + 1) * (g + 1);
  const long long rid = (long long)blockIdx.x * 8 + warp;
  if (rid >= nreg) return;
  const Region r = region_of(rid, g, S);  // r.b is the plane index here
  const float scale = (float)g / (float)S;
  const int w = r.x_hi - r.x_lo, h = r.y_hi - r.y_lo;
  int y0, y1, x0, x1;
  float t0, t1;
  bil_coord(r.y_lo, scale, g, y0, y1, t0, t1);
  bil_coord(r.x_lo, scale, g, x0, x1, t0, t1);
  float a00 = 0, a01 = 0, a10 = 0, a11 = 0;
  const float* src = dfull + (long long)r.b * S * S;
  for (int i = lane; i < w * h; i += 32) {
    const int yy = r.y_lo + i / w, xx = r.x_lo + i % w;
    int q0, q1;
    float ly0, ly1, lx0, lx1;
    bil_coord(yy, scale, g, q0, q1, ly0, ly1);
    bil_coord(xx, scale, g, q0, q1, lx0, lx1);
    const float v = src[(long long)yy * S + xx];
    a00 += v * ly0 * lx0; a01 += v * ly0 * lx1; a10 += v * ly1 * lx0; a11 += v * ly1 * lx1;
  }
  a00 = warp_sum(a00); a01 = warp_sum(a01); a10 = warp_sum(a10); a11 = warp_sum(a11);
  if (lane == 0) {
    float* d = dlow + (long long)r.b * g * g;
    atomicAdd(&d[y0 * g + x0], a00);
    atomicAdd(&d[y0 * g + x1], a01);
    atomicAdd(&d[y1 * g + x0], a10);
    atomicAdd(&d[y1 * g + x1], a11);
  }
}

// The "image in shared memory" kernels stage only the low-resolution rows their output-row chunk interpolates from
// (all of them when one block handles the whole image): C * g * g floats no longer bound the grid size, so patch-4
// models (g = 56, C = 17: 213 KB for whole planes) run.  Returns the plane stride (rows * g) and sets `base` so that
// base + c * stride indexes like a full plane: base[c * stride + y * g + x] for staged rows y.
__device__ __forceinline__ int low_rows_needed(int y_begin, int y_end, float scale, int g, int& r0) {
  int a0, a1, b0, b1;
  float t0, t1;
  bil_coord(y_begin, scale, g, a0, a1, t0, t1);
  bil_coord(y_end - 1, scale, g, b0, b1, t0, t1);
  r0 = a0;
  return b1 - a0 + 1;
}
__device__ __forceinline__ const float* stage_low_rows(const float* __restrict__ low_b, float* s_low, int C, int g,
                                                       int r0, int nrows) {
  const int per = nrows * g;
  for (int i = threadIdx.x; i < C * per; i += blockDim.x) {
    const int c = i / per, o = i - c * per;
    s_low[i] = low_b[(long long)c * g * g + r0 * g + o];
  }
  return s_low - r0 * g;
}
static int low_rows_bound(int rows_per, int g, int S) {   // host: upper bound of low_rows_needed for a chunk
  const int n = (rows_per * g + S - 1) / S + 2;
  return n < g ? n : g;
}

__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ low, uint8_t* __restrict__ mask, int C, int g, int S, int chunks) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_low_raw[];  // [C][staged rows][g]
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const float scale = (float)g / (float)S;
  const int S4 = S / 4;
  const int rows_per = (S + chunks - 1) / chunks;
  const int y_begin = chunk * rows_per, y_end = min(S, y_begin + rows_per);
  if (y_begin >= y_end) return;
  int r0;
  const int pstride = low_rows_needed(y_begin, y_end, scale, g, r0) * g;
  const float* s_low = stage_low_rows(low + (long long)b * C * g * g, s_low_raw, C, g, r0, pstride / g);
  __syncthreads();
  for (int idx = y_begin * S4 + threadIdx.x; idx < y_end * S4; idx += blockDim.x) {
    const int y = idx / S4, x4 = (idx - y * S4) * 4;
    int y0, y1;
    float ly0, ly1;
    bil_coord(y, scale, g, y0, y1, ly0, ly1);
    uint8_t res[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int x0, x1;
      float lx0, lx1;
      bil_coord(x4 + k, scale, g, x0, x1, lx0, lx1);
      if (C == 1) {
        res[k] = bil_sample(s_low, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1) > 0.0f ? 1 : 0;
      } else {
        float best = -INFINITY;
        int bi = 0;
        for (int c = 0; c < C; ++c) {
          const float v = bil_sample(s_low + c * pstride, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1);
          if (v > best) { best = v; bi = c; }
        }
        res[k] = (uint8_t)bi;
      }
    }
    *reinterpret_cast<uchar4*>(mask + ((long long)b * S + y) * S + x4) = make_uchar4(res[0], res[1], res[2], res[3]);
  }
}

// Region form of the fused upsample + argmax (C <= 32): one warp per (image, region), lane = column.  As in the fused
// cross-entropy, the column-constant part of the bilinear interpolation is hoisted: z_c(y, x) = ly0 * a_c + ly1 * b_c with
// a_c, b_c in registers, so a pixel costs 2 FMAs + a compare/select pair per class instead of 4 shared-memory reads and
// ~12 instructions (r01: the fused mask path was SLOWER than writing 218 MB of logits and taking their argmax).
template <int CMAX, bool kExact>
__global__ void __launch_bounds__(256)
upsample_argmax_region_kernel(const float* __restrict__ low, uint8_t* __restrict__ mask, int B, int C, int g, int S) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_cell[8][4][CMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nreg = (long long)B * (g + 1) * (g + 1);
  const long long rid = (long long)blockIdx.x * 8 + warp;
  if (rid >= nreg) return;
  const Region r = region_of(rid, g, S);
  const float scale = (float)g / (float)S;
  const int P = S / g;
  int y0, y1, x0, x1;
  float t0, t1;
  bil_coord(r.y_lo, scale, g, y0, y1, t0, t1);
  bil_coord(r.x_lo, scale, g, x0, x1, t0, t1);
  const float* lb = low + (long long)r.b * C * g * g;
  for (int c = lane; c < C; c += 32) {
    s_cell[warp][0][c] = lb[c * g * g + y0 * g + x0];
    s_cell[warp][1][c] = lb[c * g * g + y0 * g + x1];
    s_cell[warp][2][c] = lb[c * g * g + y1 * g + x0];
    s_cell[warp][3][c] = lb[c * g * g + y1 * g + x1];
  }
  __syncwarp();
  const int xi = lane % P, roff = lane / P, rstep = (32 / P) > 0 ? (32 / P) : 1;
  const int x = r.x_lo + xi;
  if (x >= r.x_hi || lane >= P * rstep) return;
  int q0, q1;
  float lx0, lx1;
  bil_coord(x, scale, g, q0, q1, lx0, lx1);
  float a[CMAX], bq[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    a[c] = 0.0f; bq[c] = 0.0f;
    if (kExact || c < C) {
      a[c] = lx0 * s_cell[warp][0][c] + lx1 * s_cell[warp][1][c];
      bq[c] = lx0 * s_cell[warp][2][c] + lx1 * s_cell[warp][3][c];
    }
  }
  uint8_t* out = mask + (long long)r.b * S * S + x;
  for (int y = r.y_lo + roff; y < r.y_hi; y += rstep) {
    float ly0, ly1;
    bil_coord(y, scale, g, q0, q1, ly0, ly1);
    int bi = 0;
    if (CMAX == 1) {
      bi = (ly0 * a[0] + ly1 * bq[0]) > 0.0f ? 1 : 0;
    } else {
      float best = -INFINITY;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (kExact || c < C) {
          const float v = ly0 * a[c] + ly1 * bq[c];
          if (v > best) { best = v; bi = c; }
        }
      }
    }
    out[(long long)y * S] = (uint8_t)bi;
  }
}

// class map -> colour image through the class palette: colored_pred = index_to_color[pred_labels]
// (model/CE/testViTModel.py:139-143; the mask_image the worker posts back, backend/core/views.py:116-149).
// uint8 [n] class ids -> uint8 [n, 3] RGB; 4 pixels (12 output bytes = three 32-bit words) per thread.
__global__ void __launch_bounds__(256)
colorize_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ palette, uint8_t* __restrict__ rgb,
                long long n, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint8_t s_pal[256 * 3];
  for (int i = threadIdx.x; i < 256 * 3; i += blockDim.x) s_pal[i] = i < C * 3 ? palette[i] : 0;
  __syncthreads();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uchar4 m = reinterpret_cast<const uchar4*>(mask)[i];
    const uint8_t* a = s_pal + 3 * m.x;
    const uint8_t* b = s_pal + 3 * m.y;
    const uint8_t* c = s_pal + 3 * m.z;
    const uint8_t* d = s_pal + 3 * m.w;
    uint32_t* o = reinterpret_cast<uint32_t*>(rgb) + 3 * i;
    o[0] = a[0] | (a[1] << 8) | (a[2] << 16) | ((uint32_t)b[0] << 24);
    o[1] = b[1] | (b[2] << 8) | (c[0] << 16) | ((uint32_t)c[1] << 24);
    o[2] = c[2] | (d[0] << 8) | (d[1] << 16) | ((uint32_t)d[2] << 24);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    for (int k = 0; k < 3; ++k) rgb[3 * i + k] = s_pal[3 * mask[i] + k];
  }
}

// ------------------------------------------------------------------------------------------------
// fused upsample + argmax + segmentation statistics: the predicted class map AND, per image and
// class, the exact pixel counts  counts[b][c] = {intersection, predicted, target}  from which pixel accuracy, IoU, Dice,
// precision and recall follow (model/PAED/classes.py:430-447,684-689; model/PAED/segmentation.py:38-86;
// model/CE/datasetTestViTmodel.py:188-217) — the reference recomputes them with ~100 eager passes per step.
// Same prediction rule as upsample_argmax_kernel (C == 1: class 1 iff logit > 0, two classes counted).
// Counting: per-warp shared histograms; the lanes of a warp that hit the same counter are merged with
// __match_any_sync so one lane adds the population count.
// ------------------------------------------------------------------------------------------------
constexpr int kStatMaxClasses = 32;
__device__ __forceinline__ void warp_hist_add(int* hist, int key, bool valid) {
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned peers = __match_any_sync(act, key);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) hist[key] += __popc(peers);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(256)
upsample_argmax_stats_kernel(const float* __restrict__ low, const long long* __restrict__ labels,
                             uint8_t* __restrict__ mask, int* __restrict__ counts, int C, int g, int S, int chunks) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_low_raw[];  // [C][staged rows][g]
  __shared__ int s_hist[8][3][kStatMaxClasses];   // per warp: intersection / predicted / target
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int warp = threadIdx.x >> 5;
  const int NC = C == 1 ? 2 : C;
  const float scale = (float)g / (float)S;
  const int rows_per = (S + chunks - 1) / chunks;
  const int y_begin = chunk * rows_per, y_end = min(S, y_begin + rows_per);
  if (y_begin >= y_end) return;
  int r0;
  const int pstride = low_rows_needed(y_begin, y_end, scale, g, r0) * g;
  const float* s_low = stage_low_rows(low + (long long)b * C * g * g, s_low_raw, C, g, r0, pstride / g);
  for (int i = threadIdx.x; i < 8 * 3 * kStatMaxClasses; i += blockDim.x) (&s_hist[0][0][0])[i] = 0;
  __syncthreads();
  const int npix = (y_end - y_begin) * S;
  for (int base = 0; base < npix; base += blockDim.x) {   // whole warps stay converged for the match / ballot
    const int idx = base + threadIdx.x;
    const bool valid = idx < npix;
    int pred = 0, tgt = -1;
    if (valid) {
      const int y = y_begin + idx / S, x = idx % S;
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bil_coord(y, scale, g, y0, y1, ly0, ly1);
      bil_coord(x, scale, g, x0, x1, lx0, lx1);
      if (C == 1) {
        pred = bil_sample(s_low, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1) > 0.0f ? 1 : 0;
      } else {
        float best = -INFINITY;
        for (int c = 0; c < C; ++c) {
          const float v = bil_sample(s_low + c * pstride, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1);
          if (v > best) { best = v; pred = c; }
        }
      }
      const long long o = ((long long)b * S + y) * S + x;
      if (mask != nullptr) mask[o] = (uint8_t)pred;
      const long long t = labels[o];
      tgt = (t >= 0 && t < NC) ? (int)t : -1;   // ignore_index / out-of-range targets are not counted
    }
    warp_hist_add(s_hist[warp][1], pred, valid);
    warp_hist_add(s_hist[warp][2], tgt, valid && tgt >= 0);
    warp_hist_add(s_hist[warp][0], pred, valid && tgt == pred);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * NC; i += blockDim.x) {
    const int k = i / NC, c = i - k * NC;
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_hist[w][k][c];
    if (t != 0) atomicAdd(&counts[((long long)b * NC + c) * 3 + k], t);
  }
}

// ================================================================================================
// fused upsample + cross-entropy (+ gradient onto the low-res grid), labels gathered at their own resolution
//
// One warp per (image, region); lane = column within the region (x fixed per lane), rows strided.  Everything that is
// constant along a column is hoisted out of the pixel loop:
//   z_c(y, x) = ly0 * (lx0 v00_c + lx1 v01_c) + ly1 * (lx0 v10_c + lx1 v11_c) = ly0 * a_c + ly1 * b_c
// with a_c, b_c in registers (pre-multiplied by log2 e, so the softmax needs one FADD + one MUFU.EX2 per class), i.e.
// 9 instructions per class and pixel instead of the 4 shared-memory reads + ~25 instructions of the first version
// (ncu r01: 180 us, 61 % issue-bound, 152 GB/s).  The per-lane gradient partials (2 * C values: rows y0 / y1) are
// combined across the warp through a padded shared-memory tile (each lane then owns ~2 of the 4 * C (class, cell) sums and
// does 32 FMAs per sum) instead of 4 * C five-step shuffle reductions.
// The legacy-'nearest' resize of the label map (model/CE/classes.py:273-274: F.interpolate(y.float(), size, 'nearest'),
// src = min(floor(dst * in/out), in - 1)) is folded into the label read: labels may come at their stored resolution
// [B, LH, LW] (256 x 256 from the dataset, model/CE/classes.py:77) as int64 or uint8.
// ================================================================================================
template <int CMAX>
struct CeCfg {
  static constexpr int kWarps = CMAX <= 17 ? 8 : 4;
  static constexpr int kPad = 36;   // floats per row of the reduction tile: LDS.128 of a quarter-warp stays conflict-free
};

// kExact: C == CMAX, so the per-class `c < C` tests vanish at compile time.  With a runtime C the compiler cannot keep 17
// class predicates alive across the pixel loop and re-derives them in every phase of every pixel (SASS r02: 17 ISETP +
// 17 FSEL + 33 LOP3 in the max phase alone; 340 instructions per pixel, of which the mathematics needs ~190).
template <int CMAX, typename LabelT, bool kExact>
__global__ void __launch_bounds__(CeCfg<CMAX>::kWarps * 32)
upsample_ce_kernel(const float* __restrict__ low, const LabelT* __restrict__ labels, int LH, int LW, float lab_sy,
                   float lab_sx, float* __restrict__ loss_sum, float* __restrict__ dlow, int B, int C, int g, int S) {
  pdl_wait();
  pdl_trigger();
  constexpr int kWarps = CeCfg<CMAX>::kWarps;
  constexpr int kPad = CeCfg<CMAX>::kPad;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  __shared__ float s_cell[kWarps][4][CMAX];
  __shared__ __align__(16) float s_acc[kWarps][2 * CMAX][kPad];
  __shared__ __align__(16) float s_lx[kWarps][2][32];
  __shared__ float s_red[kWarps][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nreg = (long long)B * (g + 1) * (g + 1);
  const long long rid = (long long)blockIdx.x * kWarps + warp;
  float loss = 0.0f, cnt = 0.0f;
  if (rid < nreg) {
    const Region r = region_of(rid, g, S);
    const float scale = (float)g / (float)S;
    const int P = S / g;
    int y0, y1, x0, x1;
    float t0, t1;
    bil_coord(r.y_lo, scale, g, y0, y1, t0, t1);
    bil_coord(r.x_lo, scale, g, x0, x1, t0, t1);
    const float* lb = low + (long long)r.b * C * g * g;
    for (int c = lane; c < (kExact ? CMAX : C); c += 32) {
      s_cell[warp][0][c] = lb[c * g * g + y0 * g + x0];
      s_cell[warp][1][c] = lb[c * g * g + y0 * g + x1];
      s_cell[warp][2][c] = lb[c * g * g + y1 * g + x0];
      s_cell[warp][3][c] = lb[c * g * g + y1 * g + x1];
    }
    const int xi = lane % P, roff = lane / P, rstep = (32 / P) > 0 ? (32 / P) : 1;
    const int x = r.x_lo + xi;
    const bool x_ok = (x < r.x_hi) && (lane < P * rstep);
    int q0, q1;
    float lx0 = 0.0f, lx1 = 0.0f;
    if (x_ok) bil_coord(x, scale, g, q0, q1, lx0, lx1);
    // the label terms of the gradient (-1 at the label class) go straight into the reduction tile (column = lane)
#pragma unroll
    for (int k = 0; k < 2 * CMAX; ++k) s_acc[warp][k][lane] = 0.0f;
    s_lx[warp][0][lane] = lx0;
    s_lx[warp][1][lane] = lx1;
    __syncwarp();
    float a[CMAX], bq[CMAX], acc0[CMAX], acc1[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      acc0[c] = 0.0f; acc1[c] = 0.0f;
      a[c] = 0.0f; bq[c] = 0.0f;
      if (kExact || c < C) {
        // PyTorch's association order: ly0*(lx0*v00 + lx1*v01) + ly1*(lx0*v10 + lx1*v11); log2(e) folded in afterwards
        a[c] = (lx0 * s_cell[warp][0][c] + lx1 * s_cell[warp][1][c]) * kLog2e;
        bq[c] = (lx0 * s_cell[warp][2][c] + lx1 * s_cell[warp][3][c]) * kLog2e;
      }
    }
    if (x_ok) {
      const int lx_src = (LW == S) ? x : min((int)floorf((float)x * lab_sx), LW - 1);
      const LabelT* lab_b = labels + (long long)r.b * LH * LW + lx_src;
      auto label_at = [&](int y) -> int {
        const int ys = (LH == S) ? y : min((int)floorf((float)y * lab_sy), LH - 1);
        return (int)lab_b[(long long)ys * LW];
      };
      int y = r.y_lo + roff;
      int next_label = y < r.y_hi ? label_at(y) : -100;
      for (; y < r.y_hi; y += rstep) {
        const int label = next_label;
        if (y + rstep < r.y_hi) next_label = label_at(y + rstep);
        float ly0, ly1;
        bil_coord(y, scale, g, q0, q1, ly0, ly1);
        float z[CMAX];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          if (kExact || c < C) {
            z[c] = ly0 * a[c] + ly1 * bq[c];     // log2(e) * logit
            m = fmaxf(m, z[c]);
          }
        }
        float ssum = 0.0f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          if (kExact || c < C) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z[c] - m));
            z[c] = e;
            ssum += e;
          }
        }
        if (label != -100) {
          // logit of the label class, interpolated directly (no per-class select chain in the loop above)
          const int lc = min(max(label, 0), C - 1);
          const float zl = ly0 * (lx0 * s_cell[warp][0][lc] + lx1 * s_cell[warp][1][lc]) +
                           ly1 * (lx0 * s_cell[warp][2][lc] + lx1 * s_cell[warp][3][lc]);
          float lg;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(ssum));
          loss += (m + lg) * kLn2 - zl;
          cnt += 1.0f;
          float inv;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ssum));
          const float w0 = ly0 * inv, w1 = ly1 * inv;
#pragma unroll
          for (int c = 0; c < CMAX; ++c) {
            if (kExact || c < C) {
              acc0[c] = fmaf(z[c], w0, acc0[c]);
              acc1[c] = fmaf(z[c], w1, acc1[c]);
            }
          }
          s_acc[warp][2 * lc][lane] -= ly0;
          s_acc[warp][2 * lc + 1][lane] -= ly1;
        }
      }
    }
    if (dlow != nullptr) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (kExact || c < C) {
          s_acc[warp][2 * c][lane] += acc0[c];
          s_acc[warp][2 * c + 1][lane] += acc1[c];
        }
      }
      __syncwarp();
      // output o = (class c, row ky, column kx): sum over lanes of s_acc[2c + ky][lane] * lx_kx[lane]
      float* d = dlow + (long long)r.b * C * g * g;
      for (int o = lane; o < 4 * C; o += 32) {
        const int row = o >> 1, kx = o & 1;
        const float4* av = reinterpret_cast<const float4*>(&s_acc[warp][row][0]);
        const float4* wv = reinterpret_cast<const float4*>(&s_lx[warp][kx][0]);
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 p4 = av[i], w4 = wv[i];
          t = fmaf(p4.x, w4.x, t); t = fmaf(p4.y, w4.y, t); t = fmaf(p4.z, w4.z, t); t = fmaf(p4.w, w4.w, t);
        }
        const int c = row >> 1, ky = row & 1;
        atomicAdd(&d[c * g * g + (ky ? y1 : y0) * g + (kx ? x1 : x0)], t);
      }
    }
  }
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  if (lane == 0) { s_red[warp][0] = loss; s_red[warp][1] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.0f, n = 0.0f;
    for (int i = 0; i < kWarps; ++i) { l += s_red[i][0]; n += s_red[i][1]; }
    atomicAdd(&loss_sum[0], l);
    atomicAdd(&loss_sum[1], n);
  }
}

// ================================================================================================
// PAED binary (C = 1): one block per (image, region)
// stats[b*8 + {0..6}] = {sum bce, sum p*t, sum p, sum t, sum sdf_int*p, sum sdf_ext*edge, -}; keys[b] = packed
// (edge bits << 32 | ~pixel index) maximum, i.e. the per-image max edge and its FIRST arg-max (torch.max semantics).
// ================================================================================================
constexpr int kMaxP = 32;   // largest patch size (pixels per grid cell)
constexpr int kMaxG = 64;   // largest grid side the PAED-binary kernels stage (patch 4 at 224: 56)

__device__ __forceinline__ float sigmoid_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

// fills s_p[(hh+2*halo) x (ww+2*halo)] with sigmoid(upsampled logit) (0 outside the image)
__device__ __forceinline__ void fill_prob_tile(float* s_p, const float* s_low, const Region& r, int halo, int g, int S) {
  const float scale = (float)g / (float)S;
  const int tw = (r.x_hi - r.x_lo) + 2 * halo, th = (r.y_hi - r.y_lo) + 2 * halo;
  for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
    const int y = r.y_lo - halo + i / tw, x = r.x_lo - halo + i % tw;
    float p = 0.0f;
    if (y >= 0 && y < S && x >= 0 && x < S) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bil_coord(y, scale, g, y0, y1, ly0, ly1);
      bil_coord(x, scale, g, x0, x1, lx0, lx1);
      p = sigmoid_acc(bil_sample(s_low, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1));
    }
    s_p[i] = p;
  }
}

__device__ __forceinline__ void sobel_at(const float* s_p, int tw, int ty, int tx, float& gx, float& gy) {
  const float a = s_p[(ty - 1) * tw + tx - 1], b = s_p[(ty - 1) * tw + tx], c = s_p[(ty - 1) * tw + tx + 1];
  const float d = s_p[ty * tw + tx - 1], f = s_p[ty * tw + tx + 1];
  const float gq = s_p[(ty + 1) * tw + tx - 1], hq = s_p[(ty + 1) * tw + tx], iq = s_p[(ty + 1) * tw + tx + 1];
  gx = (a - c) + 2.0f * (d - f) + (gq - iq);
  gy = (a + 2.0f * b + c) - (gq + 2.0f * hq + iq);
}

__global__ void __launch_bounds__(256)
paed_binary_stats_kernel(const float* __restrict__ low, const float* __restrict__ mask, const float* __restrict__ sdf_ext,
                         const float* __restrict__ sdf_int, float* __restrict__ stats,
                         unsigned long long* __restrict__ keys, int B, int g, int S) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_low[kMaxG * kMaxG];
  __shared__ float s_p[(kMaxP + 2) * (kMaxP + 2)];
  __shared__ float s_red[8][6];
  __shared__ unsigned long long s_key[8];
  const Region r = region_of(blockIdx.x, g, S);
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) s_low[i] = low[(long long)r.b * g * g + i];
  __syncthreads();
  fill_prob_tile(s_p, s_low, r, 1, g, S);
  __syncthreads();
  const int w = r.x_hi - r.x_lo, h = r.y_hi - r.y_lo, tw = w + 2;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long key = 0ull;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int ly = i / w, lx = i % w;
    const int y = r.y_lo + ly, x = r.x_lo + lx;
    const long long pix = ((long long)r.b * S + y) * S + x;
    const float p = s_p[(ly + 1) * tw + lx + 1];
    const float t = mask[pix];
    float gx, gy;
    sobel_at(s_p, tw, ly + 1, lx + 1, gx, gy);
    const float edge = sqrtf(gx * gx + gy * gy + 1e-6f);
    const float lp = fmaxf(logf(p), -100.0f), lq = fmaxf(logf(1.0f - p), -100.0f);
    acc[0] += -(t * lp + (1.0f - t) * lq);
    acc[1] += p * t;
    acc[2] += p;
    acc[3] += t;
    acc[4] += sdf_int[pix] * p;
    acc[5] += sdf_ext[pix] * edge;
    const unsigned long long k =
        ((unsigned long long)__float_as_uint(edge) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(y * S + x));
    key = k > key ? k : key;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 6; ++i) acc[i] = warp_sum(acc[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) s_red[warp][i] = acc[i];
    s_key[warp] = key;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += s_red[i][threadIdx.x];
    atomicAdd(&stats[r.b * 8 + threadIdx.x], s);
  }
  if (threadIdx.x == 32) {
    unsigned long long k = 0ull;
    for (int i = 0; i < 8; ++i) k = s_key[i] > k ? s_key[i] : k;
    atomicMax(&keys[r.b], k);
  }
}

// coef[b*8 + {0..6}] = dL/d{bce_sum, inter, psum, tsum, int_sum, ext_sum, max_edge}
__global__ void __launch_bounds__(256)
paed_binary_bwd_kernel(const float* __restrict__ low, const float* __restrict__ mask, const float* __restrict__ sdf_ext,
                       const float* __restrict__ sdf_int, const float* __restrict__ coef,
                       const unsigned long long* __restrict__ keys, float* __restrict__ dlow, int B, int g, int S) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_low[kMaxG * kMaxG];
  __shared__ float s_p[(kMaxP + 4) * (kMaxP + 4)];
  __shared__ float s_gx[(kMaxP + 2) * (kMaxP + 2)];
  __shared__ float s_gy[(kMaxP + 2) * (kMaxP + 2)];
  __shared__ float s_red[8][4];
  const Region r = region_of(blockIdx.x, g, S);
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) s_low[i] = low[(long long)r.b * g * g + i];
  __syncthreads();
  fill_prob_tile(s_p, s_low, r, 2, g, S);
  __syncthreads();
  const float* cf = coef + r.b * 8;
  const float c_bce = cf[0], c_inter = cf[1], c_psum = cf[2], c_int = cf[4], c_ext = cf[5], c_max = cf[6];
  const int arg = (int)(0xFFFFFFFFu - (uint32_t)(keys[r.b] & 0xFFFFFFFFull));
  const int w = r.x_hi - r.x_lo, h = r.y_hi - r.y_lo;
  const int tw2 = w + 4, tw1 = w + 2, th1 = h + 2;
  // dL/dgx, dL/dgy on the region + 1 halo
  for (int i = threadIdx.x; i < tw1 * th1; i += blockDim.x) {
    const int ly = i / tw1, lx = i % tw1;
    const int y = r.y_lo - 1 + ly, x = r.x_lo - 1 + lx;
    float vx = 0.0f, vy = 0.0f;
    if (y >= 0 && y < S && x >= 0 && x < S) {
      float gx, gy;
      sobel_at(s_p, tw2, ly + 1, lx + 1, gx, gy);
      const float edge = sqrtf(gx * gx + gy * gy + 1e-6f);
      float de = c_ext * sdf_ext[((long long)r.b * S + y) * S + x];
      if (y * S + x == arg) de += c_max;
      vx = de * gx / edge;
      vy = de * gy / edge;
    }
    s_gx[i] = vx;
    s_gy[i] = vy;
  }
  __syncthreads();
  const float scale = (float)g / (float)S;
  int y0, y1, x0, x1;
  float t0, t1;
  bil_coord(r.y_lo, scale, g, y0, y1, t0, t1);
  bil_coord(r.x_lo, scale, g, x0, x1, t0, t1);
  float a00 = 0, a01 = 0, a10 = 0, a11 = 0;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int ly = i / w, lx = i % w;
    const int y = r.y_lo + ly, x = r.x_lo + lx;
    const long long pix = ((long long)r.b * S + y) * S + x;
    const float p = s_p[(ly + 2) * tw2 + lx + 2];
    const float t = mask[pix];
    float dp = c_bce * (p - t) / fmaxf(p * (1.0f - p), 1e-12f) + c_inter * t + c_psum + c_int * sdf_int[pix];
    const int cy = ly + 1, cx = lx + 1;  // position in the halo-1 tiles
#define GX(dy, dx) s_gx[(cy + (dy)) * tw1 + cx + (dx)]
#define GY(dy, dx) s_gy[(cy + (dy)) * tw1 + cx + (dx)]
    dp += (GX(1, 1) - GX(1, -1)) + 2.0f * (GX(0, 1) - GX(0, -1)) + (GX(-1, 1) - GX(-1, -1));
    dp += (GY(1, 1) + 2.0f * GY(1, 0) + GY(1, -1)) - (GY(-1, 1) + 2.0f * GY(-1, 0) + GY(-1, -1));
#undef GX
#undef GY
    const float dz = dp * p * (1.0f - p);
    int q0, q1;
    float ly0, ly1, lx0, lx1;
    bil_coord(y, scale, g, q0, q1, ly0, ly1);
    bil_coord(x, scale, g, q0, q1, lx0, lx1);
    a00 += dz * ly0 * lx0; a01 += dz * ly0 * lx1; a10 += dz * ly1 * lx0; a11 += dz * ly1 * lx1;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  a00 = warp_sum(a00); a01 = warp_sum(a01); a10 = warp_sum(a10); a11 = warp_sum(a11);
  if (lane == 0) { s_red[warp][0] = a00; s_red[warp][1] = a01; s_red[warp][2] = a10; s_red[warp][3] = a11; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += s_red[i][threadIdx.x];
    const int yy = (threadIdx.x & 2) ? y1 : y0, xx = (threadIdx.x & 1) ? x1 : x0;
    atomicAdd(&dlow[(long long)r.b * g * g + yy * g + xx], s);
  }
}

// ================================================================================================
// PAED multi-class soft loss
// ================================================================================================
// per pixel softmax of the upsampled logits (low-res logits of the image in smem), three passes of the loss:
//   mode 0: out[c] = onehot_c - p_c                                   (dense; its blur is t)
//   mode 1: loss += 2 (1 - p_l) |t_l|;  out[c] = c == l ? 2 (1 - p_l) sign(t_l) : 0   (l = label; dense u)
//   mode 2: dp_c = -blur(u)_c - [c == l] 2 |t_l|;  out[c] = dz_c = p_c (dp_c - sum_k p_k dp_k)   (dense; the adjoint
//           of the bilinear upsample then carries it to the low-res logits)
// Only the label channel of t is ever read (the class-mismatch penalty is zero elsewhere): a 4-byte gather per pixel.
template <int CMAX>
__global__ void __launch_bounds__(256)
pm_pixel_kernel(const float* __restrict__ low, const long long* __restrict__ labels, const float* __restrict__ tin,
                const float* __restrict__ bu, float* __restrict__ out, float* __restrict__ loss_sum, int mode, int C,
                int g, int S, int chunks) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_low_raw[];  // [C][staged rows][g]
  __shared__ float s_red[8];
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const float scale = (float)g / (float)S;
  const int rows_per = (S + chunks - 1) / chunks;
  const int y_begin = chunk * rows_per, y_end = min(S, y_begin + rows_per);
  if (y_begin >= y_end) return;
  int r0;
  const int pstride = low_rows_needed(y_begin, y_end, scale, g, r0) * g;
  const float* s_low = stage_low_rows(low + (long long)b * C * g * g, s_low_raw, C, g, r0, pstride / g);
  __syncthreads();
  const long long plane = (long long)S * S;
  float loss = 0.0f;
  for (int idx = y_begin * S + threadIdx.x; idx < y_end * S; idx += blockDim.x) {
    const int y = idx / S, x = idx - y * S;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bil_coord(y, scale, g, y0, y1, ly0, ly1);
    bil_coord(x, scale, g, x0, x1, lx0, lx1);
    const int label = (int)labels[((long long)b * S + y) * S + x];
    const long long base = (long long)b * C * plane + (long long)y * S + x;
    float tl = 0.0f;
    if (mode != 0 && label >= 0 && label < C) tl = tin[base + label * plane];
    float z[CMAX];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        z[c] = bil_sample(s_low + c * pstride, g, y0, y1, ly0, ly1, x0, x1, lx0, lx1);
        m = fmaxf(m, z[c]);
      }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { z[c] = expf(z[c] - m); s += z[c]; }
    const float inv = 1.0f / s;
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) out[base + c * plane] = ((c == label) ? 1.0f : 0.0f) - z[c] * inv;
    } else if (mode == 1) {
      float pen = 0.0f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C && c == label) pen = (1.0f - z[c] * inv) * 2.0f;
      loss += pen * fabsf(tl);
      const float sg = pen * (tl > 0.0f ? 1.0f : (tl < 0.0f ? -1.0f : 0.0f));
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) out[base + c * plane] = (c == label) ? sg : 0.0f;
    } else {
      float dp[CMAX];
      float dot = 0.0f;
      const float pen = 2.0f * fabsf(tl);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          z[c] *= inv;  // p_c
          float d = -bu[base + c * plane];
          if (c == label) d -= pen;
          dp[c] = d;
          dot += z[c] * d;
        }
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) out[base + c * plane] = z[c] * (dp[c] - dot);
    }
  }
  if (mode == 1) {
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float l = 0.0f;
      for (int i = 0; i < 8; ++i) l += s_red[i];
      atomicAdd(loss_sum, l);
    }
  }
}

// 19x19 normalised Gaussian (sigma 3, zero padding) = two 19-tap passes, both inside one block: rows
// [y0-9, y0+TR+9) of a plane are staged in smem (x-padded with zeros), blurred along x into a second smem tile and then
// along y straight to the output — one HBM read and one write per element instead of two round trips, and every tap
// is an smem / register access (the first version, two global 1-D passes with 19 predicated loads per output, ran at
// 0.7 TB/s: ncu r01, 545 + 685 us per blur at B=64, C=17, 224x224).  Each thread produces 4 outputs per step from 22
// inputs (4 x 19 FMAs per 7 LDS.128 / 22 LDS.32).
__constant__ float c_gauss[19];
constexpr int kBlurTR = 32;     // output rows per block
constexpr int kBlurPad = 12;    // zero columns on both sides of a staged row (>= 9, keeps float4 alignment)

__global__ void __launch_bounds__(256)
blur2d_kernel(const float* __restrict__ in, float* __restrict__ out, int S, int tiles_y) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float s_blur[];
  const int SP = S + 2 * kBlurPad;
  constexpr int R = kBlurTR + 18;
  float* s_in = s_blur;              // [R][SP]
  float* s_x = s_blur + R * SP;      // [R][S]
  const long long pl = blockIdx.x / tiles_y;
  const int y0 = (blockIdx.x % tiles_y) * kBlurTR;
  const float* src = in + pl * S * S;
  float* dst = out + pl * S * S;
  float gk[19];
#pragma unroll
  for (int k = 0; k < 19; ++k) gk[k] = c_gauss[k];

  // stage (zero rows outside the plane, zero pad columns)
  const int S4 = S >> 2, SP4 = SP >> 2;
  for (int i = threadIdx.x; i < R * SP4; i += blockDim.x) {
    const int r = i / SP4, c4 = i - r * SP4;
    const int y = y0 - 9 + r, x = c4 * 4 - kBlurPad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < S && x >= 0 && x < S) v = *reinterpret_cast<const float4*>(src + (long long)y * S + x);
    reinterpret_cast<float4*>(s_in)[i] = v;
  }
  __syncthreads();
  // x pass: 4 consecutive outputs of one row per step
  for (int i = threadIdx.x; i < R * S4; i += blockDim.x) {
    const int r = i / S4, x0 = (i - r * S4) * 4;
    const float4* row = reinterpret_cast<const float4*>(s_in + r * SP + x0);   // element x0 - 12 of the padded row
    float v[28];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float4 t = row[q];
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 19; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaf(gk[k], v[3 + j + k], o[j]);   // input x0 + j + k - 9
    }
    *reinterpret_cast<float4*>(s_x + r * S + x0) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // y pass: 4 consecutive output rows of one column per step (lanes = consecutive columns: conflict-free, coalesced)
  for (int i = threadIdx.x; i < (kBlurTR / 4) * S; i += blockDim.x) {
    const int yb = (i / S) * 4, x = i - (i / S) * S;
    float v[22];
#pragma unroll
    for (int q = 0; q < 22; ++q) v[q] = s_x[(yb + q) * S + x];
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 19; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaf(gk[k], v[j + k], o[j]);       // input row y0 + yb + j + k - 9
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int y = y0 + yb + j;
      if (y < S) dst[(long long)y * S + x] = o[j];
    }
  }
}

static int launch_blur2d(const float* in, float* out, long long planes, int S, cudaStream_t st) {
  VS_CHECK_ARG(S % 4 == 0 && S <= 512, "PAED blur: S=%d must be a multiple of 4 and <= 512", S);
  const int tiles_y = (S + kBlurTR - 1) / kBlurTR;
  const size_t smem = (size_t)(kBlurTR + 18) * (2 * S + 2 * kBlurPad) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(blur2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  VS_CHECK_ARG(planes * tiles_y < (1LL << 31), "PAED blur: too many planes");
  launch_k(blur2d_kernel, dim3((unsigned)(planes * tiles_y)), dim3(256), (size_t)(smem), st, in, out, S, tiles_y);
  return 0;
}

// dense-tensor form (free function paed_loss_multiclass_soft on [B,C,S,S] mask / probability tensors)
// mode 0: out = m - p ; mode 1: loss += pen*|t| (class penalty) or |t|, out = u ; mode 2: out = dp = a - bu
__global__ void __launch_bounds__(256)
pmd_elem_kernel(const float* __restrict__ m, const float* __restrict__ p, const float* __restrict__ t,
                const float* __restrict__ bu, float* __restrict__ out, float* __restrict__ loss_sum, long long n,
                int mode, int class_penalty) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_red[8];
  float loss = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      out[i] = m[i] - p[i];
    } else if (mode == 1) {
      const float tv = t[i];
      const float pen = class_penalty ? m[i] * (1.0f - p[i]) * 2.0f : 1.0f;
      loss += pen * fabsf(tv);
      out[i] = pen * (tv > 0.0f ? 1.0f : (tv < 0.0f ? -1.0f : 0.0f));
    } else {
      const float a = class_penalty ? -2.0f * m[i] * fabsf(t[i]) : 0.0f;
      out[i] = a - bu[i];
    }
  }
  if (mode == 1) {
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float l = 0.0f;
      for (int i = 0; i < 8; ++i) l += s_red[i];
      atomicAdd(loss_sum, l);
    }
  }
}

static int check_grid(const char* fn, int B, int C, int g, int S) {
  if (B <= 0 || C <= 0 || g <= 0 || S <= 0 || S % g != 0) { set_error("%s: bad shape B=%d C=%d g=%d S=%d", fn, B, C, g, S); return -1; }
  const int P = S / g;
  if (P < 2 || P > kMaxP || (P & (P - 1)) != 0) { set_error("%s: S/g=%d must be a power of two in [2,%d]", fn, P, kMaxP); return -1; }
  if (S % 4 != 0) { set_error("%s: S must be a multiple of 4", fn); return -1; }
  if (sm_count() <= 0) return -1;
  return 0;
}

static int ensure_gauss() {
  static bool done = false;
  if (done) return 0;
  // model/PAED/classes.py:341-345 builds the 19x19 kernel in fp32 as g g^T / sum(g g^T) == (g/sum g)(g/sum g)^T
  float gk[19];
  float sum = 0.0f;
  for (int k = 0; k < 19; ++k) { const float x = float(k - 9); gk[k] = expf(-(x * x) / 18.0f); sum += gk[k]; }
  for (int k = 0; k < 19; ++k) gk[k] /= sum;
  VS_CHECK_CUDA(cudaMemcpyToSymbol(c_gauss, gk, sizeof(gk)));
  done = true;
  return 0;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_upsample_bilinear_fwd(const float* low, float* full, int32_t B, int32_t C, int32_t g, int32_t S,
                                        void* stream) {
  VS_CHECK_ARG(low && full, "vs_upsample_bilinear_fwd: null pointer");
  if (int rc = check_grid("vs_upsample_bilinear_fwd", B, C, g, S)) return rc;
  const long long planes = (long long)B * C;
  int chunks = 1;
  while (planes * chunks < (long long)sm_count() * 4 && chunks < S / 8) chunks *= 2;
  launch_k(upsample_fwd_kernel, dim3((unsigned)(planes * chunks)), dim3(256), (size_t)(g * g * sizeof(float)), (cudaStream_t)stream, low, full, g, S, chunks);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_upsample_bilinear_bwd(const float* dfull, float* dlow, int32_t B, int32_t C, int32_t g, int32_t S,
                                        void* stream) {
  VS_CHECK_ARG(dfull && dlow, "vs_upsample_bilinear_bwd: null pointer");
  if (int rc = check_grid("vs_upsample_bilinear_bwd", B, C, g, S)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long planes = (long long)B * C;
  VS_CHECK_CUDA(cudaMemsetAsync(dlow, 0, (size_t)planes * g * g * sizeof(float), st));
  const long long nreg = planes * (g + 1) * (g + 1);
  launch_k(upsample_bwd_kernel, dim3((unsigned)((nreg + 7) / 8)), dim3(256), (size_t)(0), st, dfull, dlow, planes, g, S);
  VS_CHECK_LAUNCH();
  return 0;
}

// chunks (output-row groups per image) and dynamic smem of the image-in-smem kernels: enough blocks to fill the GPU,
// and more chunks when the staged low-resolution rows would not fit otherwise
static int plan_image_chunks(const char* who, int B, int C, int g, int S, int max_chunks, int* chunks_out,
                             size_t* smem_out) {
  const size_t limit = 160 * 1024;
  int chunks = 1;
  while ((long long)B * chunks < (long long)sm_count() * 4 && chunks < max_chunks) chunks *= 2;
  for (;;) {
    const int rows_per = (S + chunks - 1) / chunks;
    const size_t smem = (size_t)C * low_rows_bound(rows_per, g, S) * g * sizeof(float);
    if (smem <= limit || chunks >= S / 2) {
      if (smem > limit) { set_error("%s: C=%d, g=%d does not fit shared memory", who, C, g); return -1; }
      *chunks_out = chunks;
      *smem_out = smem;
      return 0;
    }
    chunks *= 2;
  }
}

extern "C" int vs_upsample_argmax(const float* low, uint8_t* mask, int32_t B, int32_t C, int32_t g, int32_t S,
                                  void* stream) {
  VS_CHECK_ARG(low && mask, "vs_upsample_argmax: null pointer");
  if (int rc = check_grid("vs_upsample_argmax", B, C, g, S)) return rc;
  VS_CHECK_ARG(C <= 255, "vs_upsample_argmax: C must be <= 255");
  if (C <= 32) {
    const long long nreg = (long long)B * (g + 1) * (g + 1);
    const unsigned grid = (unsigned)((nreg + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
#define VS_AM_LAUNCH(CM)                                                                                              \
  do {                                                                                                                \
    if (C == CM) launch_k(upsample_argmax_region_kernel<CM, true>, dim3(grid), dim3(256), (size_t)0, st, low, mask, B, C, g, S); \
    else launch_k(upsample_argmax_region_kernel<CM, false>, dim3(grid), dim3(256), (size_t)0, st, low, mask, B, C, g, S);        \
  } while (0)
    if (C == 1) VS_AM_LAUNCH(1);
    else if (C <= 8) VS_AM_LAUNCH(8);
    else if (C <= 17) VS_AM_LAUNCH(17);
    else VS_AM_LAUNCH(32);
#undef VS_AM_LAUNCH
    VS_CHECK_LAUNCH();
    return 0;
  }
  int chunks;
  size_t smem;
  if (int rc = plan_image_chunks("vs_upsample_argmax", B, C, g, S, S / 8, &chunks, &smem)) return rc;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(upsample_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  launch_k(upsample_argmax_kernel, dim3(B * chunks), dim3(256), (size_t)(smem), (cudaStream_t)stream, low, mask, C, g, S, chunks);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_colorize_mask(const uint8_t* mask, const uint8_t* palette, uint8_t* rgb, int64_t n, int32_t C,
                                void* stream) {
  VS_CHECK_ARG(mask && palette && rgb && n > 0 && C > 0 && C <= 256, "vs_colorize_mask: bad arguments");
  VS_CHECK_ARG(((uintptr_t)mask % 4 == 0) && ((uintptr_t)rgb % 4 == 0), "vs_colorize_mask: mask / rgb must be 4-byte aligned");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_colorize_mask: no CUDA device");
  long long nb = (n / 4 + 255) / 256;
  if (nb > (long long)nsm * 16) nb = (long long)nsm * 16;
  if (nb < 1) nb = 1;
  launch_k(colorize_kernel, dim3((unsigned)nb), dim3(256), (size_t)(0), (cudaStream_t)stream, mask, palette, rgb, n, C);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_upsample_argmax_stats(const float* low, const int64_t* labels, uint8_t* mask, int32_t* counts,
                                        int32_t B, int32_t C, int32_t g, int32_t S, void* stream) {
  VS_CHECK_ARG(low && labels && counts, "vs_upsample_argmax_stats: null pointer");
  if (int rc = check_grid("vs_upsample_argmax_stats", B, C, g, S)) return rc;
  VS_CHECK_ARG(C <= kStatMaxClasses, "vs_upsample_argmax_stats: C must be <= %d", kStatMaxClasses);
  int chunks;
  size_t smem;
  if (int rc = plan_image_chunks("vs_upsample_argmax_stats", B, C, g, S, S / 8, &chunks, &smem)) return rc;
  static size_t smem_set = 0;
  if (smem > 40 * 1024 && smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(upsample_argmax_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    smem_set = smem;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int NC = C == 1 ? 2 : C;
  VS_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * NC * 3 * sizeof(int32_t), st));
  launch_k(upsample_argmax_stats_kernel, dim3(B * chunks), dim3(256), (size_t)(smem), st, low, (const long long*)labels, mask, counts, C, g, S, chunks);
  VS_CHECK_LAUNCH();
  return 0;
}

template <int CMAX, typename LabelT>
static void launch_upsample_ce(const float* low, const void* labels, int LH, int LW, float* loss_sum, float* dlow, int B,
                               int C, int g, int S, cudaStream_t st) {
  const long long nreg = (long long)B * (g + 1) * (g + 1);
  constexpr int kWarps = CeCfg<CMAX>::kWarps;
  const unsigned grid = (unsigned)((nreg + kWarps - 1) / kWarps);
  // the scale PyTorch's nearest kernel uses: (float)input_size / output_size
  const float sy = (float)LH / (float)S, sx = (float)LW / (float)S;
  if (C == CMAX)
    launch_k(upsample_ce_kernel<CMAX, LabelT, true>, dim3(grid), dim3(kWarps * 32), (size_t)(0), st, low, (const LabelT*)labels, LH, LW, sy, sx, loss_sum, dlow, B, C, g, S);
  else
    launch_k(upsample_ce_kernel<CMAX, LabelT, false>, dim3(grid), dim3(kWarps * 32), (size_t)(0), st, low, (const LabelT*)labels, LH, LW, sy, sx, loss_sum, dlow, B, C, g, S);
}

extern "C" int vs_upsample_ce(const float* low, const void* labels, int32_t label_dtype, int32_t LH, int32_t LW,
                              float* loss_sum, float* dlow, int32_t B, int32_t C, int32_t g, int32_t S, void* stream) {
  VS_CHECK_ARG(low && labels && loss_sum, "vs_upsample_ce: null pointer");
  if (int rc = check_grid("vs_upsample_ce", B, C, g, S)) return rc;
  VS_CHECK_ARG(C <= 32, "vs_upsample_ce: C must be <= 32");
  VS_CHECK_ARG(LH > 0 && LW > 0, "vs_upsample_ce: bad label size %d x %d", LH, LW);
  VS_CHECK_ARG(label_dtype == 0 || label_dtype == 1, "vs_upsample_ce: label_dtype must be 0 (int64) or 1 (uint8)");
  cudaStream_t st = (cudaStream_t)stream;
#define VS_CE_DISPATCH(CM)                                                                                       \
  do {                                                                                                           \
    if (label_dtype == 0) launch_upsample_ce<CM, long long>(low, labels, LH, LW, loss_sum, dlow, B, C, g, S, st); \
    else launch_upsample_ce<CM, uint8_t>(low, labels, LH, LW, loss_sum, dlow, B, C, g, S, st);                   \
  } while (0)
  if (C == 1) VS_CE_DISPATCH(1);
  else if (C <= 4) VS_CE_DISPATCH(4);
  else if (C <= 8) VS_CE_DISPATCH(8);
  else if (C <= 17) VS_CE_DISPATCH(17);
  else VS_CE_DISPATCH(32);
#undef VS_CE_DISPATCH
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_paed_binary_stats(const float* low, const float* mask, const float* sdf_ext, const float* sdf_int,
                                    float* stats, uint64_t* keys, int32_t B, int32_t g, int32_t S, void* stream) {
  VS_CHECK_ARG(low && mask && sdf_ext && sdf_int && stats && keys, "vs_paed_binary_stats: null pointer");
  if (int rc = check_grid("vs_paed_binary_stats", B, 1, g, S)) return rc;
  VS_CHECK_ARG(g <= kMaxG, "vs_paed_binary_stats: g must be <= %d", kMaxG);
  const long long nreg = (long long)B * (g + 1) * (g + 1);
  launch_k(paed_binary_stats_kernel, dim3((unsigned)nreg), dim3(256), (size_t)(0), (cudaStream_t)stream, low, mask, sdf_ext, sdf_int, stats, (unsigned long long*)keys, B, g, S);
  VS_CHECK_LAUNCH();
  return 0;
}

extern "C" int vs_paed_binary_bwd(const float* low, const float* mask, const float* sdf_ext, const float* sdf_int,
                                  const float* coef, const uint64_t* keys, float* dlow, int32_t B, int32_t g, int32_t S,
                                  void* stream) {
  VS_CHECK_ARG(low && mask && sdf_ext && sdf_int && coef && keys && dlow, "vs_paed_binary_bwd: null pointer");
  if (int rc = check_grid("vs_paed_binary_bwd", B, 1, g, S)) return rc;
  VS_CHECK_ARG(g <= kMaxG, "vs_paed_binary_bwd: g must be <= %d", kMaxG);
  const long long nreg = (long long)B * (g + 1) * (g + 1);
  launch_k(paed_binary_bwd_kernel, dim3((unsigned)nreg), dim3(256), (size_t)(0), (cudaStream_t)stream, low, mask, sdf_ext, sdf_int, coef, (const unsigned long long*)keys, dlow, B, g, S);
  VS_CHECK_LAUNCH();
  return 0;
}

template <int CMAX>
static int paed_multiclass_impl(const float* low, const long long* labels, float* t1, float* t2, float* t3,
                                float* loss_sum, float* dlow, int B, int C, int g, int S, cudaStream_t st) {
  int chunks;
  size_t smem;
  if (int rc = plan_image_chunks("vs_paed_multiclass", B, C, g, S, S / 4, &chunks, &smem)) return rc;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(pm_pixel_kernel<CMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const long long planes = (long long)B * C;
  // t1 = onehot - p ; t2 = blur(t1) = t
  launch_k(pm_pixel_kernel<CMAX>, dim3(B * chunks), dim3(256), (size_t)(smem), st, low, labels, nullptr, nullptr, t1, nullptr, 0, C, g, S, chunks);
  if (int rc = launch_blur2d(t1, t2, planes, S, st)) return rc;
  // loss and u -> t1 (reads t only at the label channel)
  launch_k(pm_pixel_kernel<CMAX>, dim3(B * chunks), dim3(256), (size_t)(smem), st, low, labels, t2, nullptr, t1, loss_sum, 1, C, g, S, chunks);
  VS_CHECK_LAUNCH();
  if (dlow != nullptr) {
    // t3 = blur(u) ; t1 = d(loss)/d(upsampled logits) ; dlow = upsample^T(t1)
    if (int rc = launch_blur2d(t1, t3, planes, S, st)) return rc;
    launch_k(pm_pixel_kernel<CMAX>, dim3(B * chunks), dim3(256), (size_t)(smem), st, low, labels, t2, t3, t1, nullptr, 2, C, g, S, chunks);
    VS_CHECK_LAUNCH();
    return vs_upsample_bilinear_bwd(t1, dlow, B, C, g, S, (void*)st);
  }
  return 0;
}

extern "C" int vs_paed_multiclass(const float* low, const int64_t* labels, float* t1, float* t2, float* t3,
                                  float* loss_sum, float* dlow, int32_t B, int32_t C, int32_t g, int32_t S,
                                  void* stream) {
  VS_CHECK_ARG(low && labels && t1 && t2 && loss_sum, "vs_paed_multiclass: null pointer");
  VS_CHECK_ARG(dlow == nullptr || t3 != nullptr, "vs_paed_multiclass: t3 scratch required for the backward pass");
  if (int rc = check_grid("vs_paed_multiclass", B, C, g, S)) return rc;
  VS_CHECK_ARG(C <= 32, "vs_paed_multiclass: C must be <= 32");
  if (int rc = ensure_gauss()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long* lab = (const long long*)labels;
  if (C <= 4) return paed_multiclass_impl<4>(low, lab, t1, t2, t3, loss_sum, dlow, B, C, g, S, st);
  if (C <= 17) return paed_multiclass_impl<17>(low, lab, t1, t2, t3, loss_sum, dlow, B, C, g, S, st);
  return paed_multiclass_impl<32>(low, lab, t1, t2, t3, loss_sum, dlow, B, C, g, S, st);
}

extern "C" int vs_paed_multiclass_dense(const float* msk, const float* prob, float* t1, float* t2, float* t3,
                                        float* loss_sum, float* dprob, int32_t B, int32_t C, int32_t S,
                                        int32_t class_penalty, void* stream) {
  VS_CHECK_ARG(msk && prob && t1 && t2 && loss_sum, "vs_paed_multiclass_dense: null pointer");
  VS_CHECK_ARG(dprob == nullptr || t3 != nullptr, "vs_paed_multiclass_dense: t3 scratch required for the backward pass");
  VS_CHECK_ARG(B > 0 && C > 0 && S > 0, "vs_paed_multiclass_dense: bad shape");
  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_paed_multiclass_dense: no CUDA device");
  if (int rc = ensure_gauss()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long planes = (long long)B * C;
  const long long n = planes * S * S;
  long long bg = (n + 255) / 256;
  if (bg > (long long)nsm * 32) bg = (long long)nsm * 32;
  const unsigned grid = (unsigned)bg;
  // t1 = m - p ; t2 = blur(t1) = t ; t1 = u ; t3 = blur(u) ; dprob
  launch_k(pmd_elem_kernel, dim3(grid), dim3(256), (size_t)(0), st, msk, prob, nullptr, nullptr, t1, nullptr, n, 0, class_penalty);
  if (int rc = launch_blur2d(t1, t2, planes, S, st)) return rc;
  launch_k(pmd_elem_kernel, dim3(grid), dim3(256), (size_t)(0), st, msk, prob, t2, nullptr, t1, loss_sum, n, 1, class_penalty);
  if (dprob != nullptr) {
    if (int rc = launch_blur2d(t1, t3, planes, S, st)) return rc;
    launch_k(pmd_elem_kernel, dim3(grid), dim3(256), (size_t)(0), st, msk, prob, t2, t3, dprob, nullptr, n, 2, class_penalty);
  }
  VS_CHECK_LAUNCH();
  return 0;
}
