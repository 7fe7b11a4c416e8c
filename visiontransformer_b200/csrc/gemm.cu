// Persistent, warp-specialised bf16 GEMM on tcgen05 tensor cores for sm_100a.
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Two flavours of one kernel template:
//   CTA2 = 1 : CTA pairs (cluster 2x1) run tcgen05.mma.cta_group::2 on 256 x BN tiles.  Each CTA stages its own
//              128 rows of A and HALF of the B tile; the pair's tensor cores share the B halves, which halves both
//              the L2->SM operand traffic and the smem read bandwidth per MMA.  (ncu r01: the single-CTA M=128
//              form tops out near 50 % tensor-pipe activity with operands always ready.)
//   CTA2 = 0 : single-CTA 128 x BN tiles (cta_group::1) for problems too small to fill 74 pairs.
// Roles per CTA (384 threads):
//   warp 0      : TMA producer  — 128B-swizzled smem ring (6 x 32 KB stages for CTA2/BN=256), mbarrier completion
//   warp 1      : MMA issuer    — leader CTA only; whole warp converged, one elected lane issues 4 UMMAs per stage
//   warp 2      : TMEM allocator (512 columns = two fp32 accumulator stages)
//   warps 4..11 : epilogue      — tcgen05.ld (thread = output row), fused bias / GELU / ReLU / GELU' / ReLU-mask /
//                                 residual / bf16-or-fp32 / atomic-add stores; overlaps the next tile's MMAs
// Operands can be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) so the dgrad (dY·W) and wgrad
// (dYᵀ·X) GEMMs of the backward pass read forward tensors in place (replaces autograd's mm_backward for
// TF:216-218,262,290,305).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

constexpr int BM = 128;  // rows of A per CTA
constexpr int BK = 64;
constexpr int kEpiWarps = 16;  // 4 per TMEM lane quadrant = 4 per SM sub-partition: the epilogue is latency-bound
constexpr int kThreads = 128 + kEpiWarps * 32;

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kblocks;  // tiles_m counts (CTA2 ? 256 : 128)-row tiles; kblocks = ceil(K / BK)
  void* out;
  long long ldo;
  int out_f32, accumulate;
  const float* bias;
  int act;
  __nv_bfloat16* out2;
  long long ldo2;
  const __nv_bfloat16* aux;
  long long ldaux;
  int aux_mode;
  const float* residual;
  long long ldr;
  int row_tokens;
  DropCfg drop;  // hidden-state dropout applied to (acc + bias) before the residual add (fp32 outputs only)
  int dynamic;   // 1: tiles are handed out by cluster launch control (grid = one cluster per work item)
  float* colsum; // EPI 1: out_colsum[n] += sum over rows of the bf16-rounded output (bias gradient), or nullptr
  int raster;    // 1: column-persistent tile order (see decode_work): a unit keeps ONE column tile for all its tiles
  int per;       //    units per column tile in that order
  int tail_from; // work items >= tail_from are column SUB-tiles (BN / tail_split wide) of the tiles >= tail_from:
  int tail_split;//    the partial last wave of the static schedule is cut into 2 or 4 narrower items (1: off)
};

// EPI = 0: register-direct / smem-transposed epilogues (fp32 outputs, accumulation, BN = 192).
// EPI = 1: bf16 outputs leave through shared memory and cp.async.bulk.tensor stores: each epilogue warp owns one
//          swizzled [32 rows x BN/4 columns] bf16 staging tile (4 KB for BN = 256), which also receives the GELU' / ReLU
//          mask operand (aux) by TMA load ahead of the accumulator.
// EPI = 2: fp32 outputs (+ fp32 residual) of the 128-column tiles: two [32 x 32] fp32 tiles (128-byte rows) per warp; the
//          residual tile of the NEXT work item is fetched by TMA while the current one is processed, results are added in
//          place and leave by TMA store.
template <int BN, int CTA2, int EPI>
struct Cfg {
  static constexpr int kBNH = CTA2 ? BN / 2 : BN;                // B rows staged per CTA
  static constexpr int kBBoxes = (kBNH + 63) / 64;               // 64-wide MN-major boxes per CTA
  static constexpr int kABytes = BM * BK * 2;                    // 16 KB
  static constexpr int kBAlloc = kBBoxes * 64 * BK * 2;          // smem reserved for B per stage
  static constexpr int kStageBytes = kABytes + kBAlloc;
  static constexpr int kEpiWarpBytes = EPI == 1 ? (BN / 4) * 32 * 2 : (EPI == 2 ? 2 * 4096 : 2048);   // per-warp staging
  static constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - kBarBytes - 1024 - kEpiBytes;   // 1024: manual alignment slack
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kEpiOffset = kStages * kStageBytes;
  static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
  static constexpr int kTotal = kBarOffset + kBarBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
  static_assert(kStageBytes % 1024 == 0, "staging tiles must stay 1024-byte aligned");
  static constexpr int kAccStride = 256;                         // TMEM columns between accumulator stages
  static_assert(kStages >= 3, "pipeline too shallow");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same-offset mbarrier of CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  // relaxed: the accumulator hand-off is ordered by tcgen05.fence::before_thread_sync; a release here would make the
  // warp wait for all its outstanding global stores (ncu r01b: MEMBAR.ALL.CTA + ERRBAR per tile per warp)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// pair-wide variants of the tcgen05 helpers
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t tmem_addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs of the pair once the issued MMAs retire
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// Dynamic tile scheduling by cluster launch control (sm_100 CLC).  The kernel is launched with ONE cluster per work
// item; the clusters that get an SM pair keep running and "cancel" clusters that have not started yet, taking over
// their work item.  Unlike a persistent grid with a static tile assignment, this adapts to how many SMs the kernel
// actually received: when NCCL's all-reduce CTAs of the data-parallel step hold some SMs, a static 148-CTA grid runs
// its displaced CTAs as a second wave (SCALE_r01: GEMM roofline fraction 0.61 -> 0.53 at 8 GPUs); here the remaining
// SMs simply take more tiles.
//   scheduler (warp 3 of the leader CTA): try_cancel -> 16-byte response multicast into the same smem slot of every CTA of
//   the cluster, signalled on each CTA's sched_full[slot]; consumers (producer / MMA / epilogue warps of both CTAs) decode
//   it and release the slot on the leader's sched_empty[slot].
// ------------------------------------------------------------------------------------------------
constexpr int kSched = 4;
struct SchedSmem {
  uint4 resp[kSched];
  uint64_t full[kSched];
  uint64_t empty[kSched];
};
__device__ __forceinline__ void mbar_expect_tx_remote(uint64_t* bar, uint32_t rank, uint32_t bytes) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(bytes)
               : "memory");
}
template <int CTA2>
__device__ __forceinline__ void clc_try_cancel(uint4* resp, uint64_t* bar) {
  if (CTA2)
    asm volatile(
        "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 "
        "[%0], [%1];" ::"r"(smem_u32(resp)), "r"(smem_u32(bar))
        : "memory");
  else
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                 ::"r"(smem_u32(resp)), "r"(smem_u32(bar))
                 : "memory");
}
// -> first CTA id (x) of the cancelled cluster, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_decode(const uint4* resp) {
  uint32_t x = 0, valid = 0;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%2];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %1, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %0, r;\n\t}\n"
      : "+r"(x), "=r"(valid)
      : "r"(smem_u32(resp))
      : "memory");
  return valid ? (int)x : -1;
}
// work item -> (k split, row tile, column tile).
//   default order : w = split * tiles + tm * tiles_n + tn, handed out round-robin (unit u takes w = u, u + U, ...): the
//                   units of one "wave" share A row tiles and sweep the column tiles.
//   raster = 1    : column-persistent order for GEMMs that also form the column sums of their output: unit u keeps the
//                   column tile tn = u % tiles_n for ALL of its tiles (row tiles j, j + per, ... with j = u / tiles_n), so
//                   an epilogue warp owns the same columns throughout and accumulates their sums in registers — one
//                   reduction per warp and kernel instead of one per warp and tile (per * 8 same-address reductions per
//                   column instead of tiles_m * 8).  The grid is per * tiles_n units; splits = 1.
//   tail split    : with T tiles on U units the static schedule ends in a partial wave of T % U tiles that costs a whole
//                   tile time (M = 12608: 150 tiles of 256 x 256 on 74 pairs = 2 + 2/74 waves -> 3 tile times).  The items
//                   of that last wave are cut into tail_split column sub-tiles each (sub = 0 .. tail_split - 1, BN /
//                   tail_split columns: the MMA runs with a narrower N, the B box is loaded at the sub-tile's column
//                   offset and only the epilogue warps of the first 4 / tail_split column slices have work), so the
//                   tail costs 1/2 or 1/4 of a tile time.  sub = -1: full tile.
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int tiles, int nunits, int& split, int& tm,
                                            int& tn, int& sub) {
  sub = -1;
  if (p.raster) {
    const int u = w % nunits, i = w / nunits;
    split = 0;
    tn = u % p.tiles_n;
    tm = u / p.tiles_n + i * p.per;
  } else {
    int t;
    if (w >= p.tail_from && p.tail_split > 1) {   // (splits == 1 whenever the tail is split)
      const int j = w - p.tail_from;
      t = p.tail_from + j / p.tail_split;
      sub = j % p.tail_split;
      split = 0;
    } else {
      split = w / tiles;
      t = w - split * tiles;
    }
    tm = t / p.tiles_n;
    tn = t - tm * p.tiles_n;
  }
}

// iteration state of one consumer warp over the work items of its cluster
template <int CTA2>
struct TileIter {
  SchedSmem* ss;
  int it, step, total;
  uint32_t rank;
  bool dynamic;
  int raster_per, tiles_m, tiles_n;   // raster_per > 0: column-persistent order (decode_work)
  __device__ __forceinline__ int next(int w, int lane) {
    if (!dynamic) {
      const int wn = w + step;
      if (raster_per > 0) return ((wn % step) / tiles_n + (wn / step) * raster_per < tiles_m) ? wn : -1;
      return (wn < total) ? wn : -1;
    }
    const int slot = it & (kSched - 1);
    mbar_wait(&ss->full[slot], (uint32_t)(it / kSched) & 1u);
    const int x = clc_decode(&ss->resp[slot]);
    fence_proxy_async_smem();   // the slot is re-written by the async proxy after the release below
    __syncwarp();
    if (lane == 0) {
      if (CTA2 && rank != 0) mbar_arrive_remote(&ss->empty[slot], 0);
      else mbar_arrive_relaxed(&ss->empty[slot]);
    }
    ++it;
    return x < 0 ? -1 : x / (CTA2 ? 2 : 1);
  }
};

// ------------------------------------------------------------------------------------------------
// Epilogue math + I/O for 4 consecutive columns of one output row.  The epilogue warps transpose each 32x32
// accumulator chunk through a swizzled smem tile so that here a warp covers 4 rows x 32 contiguous columns: every
// global access (bias, residual, aux, out, out2) is a full-sector, row-contiguous request.  (The first version
// stored one row per thread: 32 sectors per request; ncu r01 l1tex st sectors/request = 32, epilogue-bound.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fn(int act, float v) {
  return act == 1 ? gelu_erf(v) : (act == 2 ? fmaxf(v, 0.0f) : v);
}
__device__ __forceinline__ float aux_fn(int mode, float v, float x) {
  return mode == 1 ? v * gelu_erf_grad(x) : (x > 0.0f ? v : 0.0f);
}

// One warp's epilogue for one tile: kChunks chunks of 32 rows x 16 columns (fp32 output).
//   * accumulators: tcgen05.ld.x16 (thread = row) -> swizzled smem transpose -> lane <-> 4 columns, 8 rows per request,
//     so every global access (bias, residual, out) is a full-sector row-contiguous request
//   * software pipeline: bias for all chunks and the first chunk's residual tile are requested BEFORE waiting for the
//     accumulator; while chunk c is processed, the tcgen05.ld and the residual loads of chunk c+1 are in flight.
// 16 epilogue warps (4 per sub-partition): ncu r01b showed the 8-warp epilogue issue-starved (IPC 0.28, stalls =
// long scoreboard + fixed-latency waits) and, at K = 768, longer than the MMA main loop.
template <int kChunks>
__device__ __forceinline__ void epilogue_tile_f32(const GemmParams& p, uint8_t* stg, uint32_t tmem_addr, uint64_t* tfull,
                                                  uint32_t tfull_phase, int row0, int col0, bool first_split, int lane) {
  const int cq = lane & 3;            // column quad within the 16-column chunk
  const int rsub = lane >> 2;         // row within each group of 8 rows
  const int r0 = row0 + rsub;
  const bool use_bias = p.bias != nullptr && first_split;
  const bool use_res = p.residual != nullptr && first_split;

  float4 bias_r[kChunks];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    bias_r[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = col0 + c * 16 + cq * 4;
    if (use_bias && n < p.N) bias_r[c] = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  }
  float4 res_r[2][4];
  auto prefetch = [&](int c, int buf) {
    const int n = col0 + c * 16 + cq * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 8 * i;
      res_r[buf][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (use_res && r < p.M && n < p.N) {
        const long long rrow = (p.row_tokens > 0) ? (r % p.row_tokens) : r;
        res_r[buf][i] = *reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + n);
      }
    }
  };
  prefetch(0, 0);

  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem_addr, v);
  const uint32_t dseed = p.drop.thresh != 0u ? drop_seed(p.drop) : 0u;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int nb = col0 + c * 16;
    tmem_ld_wait();
    // rows are 64 B apart in the staging tile; 16-byte slots XOR-swizzled by (row >> 1): conflict-free both ways
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    if (c + 1 < kChunks) tmem_ld16(tmem_addr + (c + 1) * 16, v);   // in flight while this chunk is processed
    __syncwarp();
    if (c + 1 < kChunks) prefetch(c + 1, (c + 1) & 1);
    if (nb < p.N) {
      const int n = nb + cq * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = rsub + 8 * i;
        const int r = r0 + 8 * i;
        float4 a = *reinterpret_cast<const float4*>(stg + rl * 64 + ((cq ^ ((rl >> 1) & 3)) << 4));
        if (r < p.M) {
          a.x += bias_r[c].x; a.y += bias_r[c].y; a.z += bias_r[c].z; a.w += bias_r[c].w;
          if (p.act == 1) {
            a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
          } else if (p.act == 2) {
            a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
          }
          if (p.drop.thresh != 0u) {   // TF:267 / TF:310: dropout(dense(x)) then + residual
            const uint32_t e = (uint32_t)r * (uint32_t)p.N + (uint32_t)n;
            bool k0, k1, k2, k3;
            drop_keep2(e, dseed, p.drop.thresh, k0, k1);
            drop_keep2(e + 2, dseed, p.drop.thresh, k2, k3);
            a.x = k0 ? a.x * p.drop.scale : 0.0f; a.y = k1 ? a.y * p.drop.scale : 0.0f;
            a.z = k2 ? a.z * p.drop.scale : 0.0f; a.w = k3 ? a.w * p.drop.scale : 0.0f;
          }
          const float4 rr = res_r[c & 1][i];
          a.x += rr.x; a.y += rr.y; a.z += rr.z; a.w += rr.w;
          float* o = reinterpret_cast<float*>(p.out) + (long long)r * p.ldo + n;
          if (p.accumulate) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                         : "memory");
          } else {
            *reinterpret_cast<float4*>(o) = a;
          }
        }
      }
    }
    __syncwarp();
  }
}

// bf16-output epilogue, one output row per thread straight from the tcgen05.ld registers (no smem staging: for
// 2-byte outputs the transposed form was measured slower — its shared-memory traffic competes with the UMMA operand
// reads and TMA writes of the main loop).  16-column chunks: each thread writes one full 32-byte sector per chunk.
template <int kChunks>
__device__ __forceinline__ void epilogue_tile_bf16(const GemmParams& p, uint32_t tmem_addr, uint64_t* tfull,
                                                   uint32_t tfull_phase, int row0, int col0, bool first_split, int lane) {
  const int r = row0 + lane;
  const bool row_ok = r < p.M;
  const bool add_bias = p.bias != nullptr && first_split;
  // bias and aux of a chunk do not depend on the accumulator: chunk 0 is requested before waiting for it, chunk c+1
  // while chunk c is processed.  All row-scattered accesses are 256-bit (one full sector per thread per instruction).
  u32x8 ax[2];
  u32x8 bs[2][2];
  auto prefetch = [&](int c, int buf) {
    const int n = col0 + c * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ax[buf].v[j] = 0u; bs[buf][0].v[j] = 0u; bs[buf][1].v[j] = 0u; }
    if (n < p.N) {
      if (add_bias) {
        bs[buf][0] = ld_global_nc_256(p.bias + n);
        bs[buf][1] = ld_global_nc_256(p.bias + n + 8);
      }
      if (p.aux_mode != 0 && row_ok) ax[buf] = ld_global_nc_256(p.aux + (long long)r * p.ldaux + n);
    }
  };
  prefetch(0, 0);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem_addr, v);
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int n = col0 + c * 16;
    tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    if (c + 1 < kChunks) {
      tmem_ld16(tmem_addr + (c + 1) * 16, v);
      prefetch(c + 1, (c + 1) & 1);
    }
    if (n < p.N) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] += __uint_as_float(bs[c & 1][0].v[j]);
        f[8 + j] += __uint_as_float(bs[c & 1][1].v[j]);
      }
      if (row_ok) {
        if (p.out2 != nullptr) {
          u32x8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
          st_global_256(p.out2 + (long long)r * p.ldo2 + n, o);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
        }
        if (p.aux_mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 x = unpack_bf16(ax[c & 1].v[j]);
            f[2 * j] *= gelu_erf_grad(x.x);
            f[2 * j + 1] *= gelu_erf_grad(x.y);
          }
        } else if (p.aux_mode == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 x = unpack_bf16(ax[c & 1].v[j]);
            f[2 * j] = x.x > 0.0f ? f[2 * j] : 0.0f;
            f[2 * j + 1] = x.y > 0.0f ? f[2 * j + 1] : 0.0f;
          }
        }
        u32x8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
        st_global_256(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)r * p.ldo + n, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 epilogue through shared memory + TMA (EPI = 1).
// Thread = accumulator row (tcgen05.ld 32x32b); every 16-column chunk is packed to 32 bytes and written into the warp's
// [32 x W] bf16 tile with the TMA swizzle (W = 64: 128-byte rows, SWIZZLE_128B; W = 32: 64-byte rows, SWIZZLE_64B), which
// makes the 16-byte st.shared of a quarter-warp hit 32 different banks.  One elected lane then issues a single
// cp.async.bulk.tensor store for the tile: the LSU sees no row-scattered global access at all (ncu r01: the
// 32-sectors-per-request st.global / ld.global of the register-direct form bounded the K = 768 GEMMs with two bf16
// streams).  The aux operand (pre-activation for GELU', ReLU mask) arrives in the SAME tile by TMA load, issued as soon
// as the previous tile's store has drained the buffer, i.e. a whole MMA main loop ahead of its use; each thread reads
// its own row chunk and overwrites it in place with the result.  Two outputs (fc1 forward: pre-activation copy +
// GELU) take two passes over the TMEM accumulator through the one buffer.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int W>
__device__ __forceinline__ uint32_t stg_off(int lane, int unit16) {
  if constexpr (W == 64) return (uint32_t)lane * 128u + (uint32_t)((unit16 ^ (lane & 7)) << 4);
  else return (uint32_t)lane * 64u + (uint32_t)((unit16 ^ ((lane >> 1) & 3)) << 4);
}

template <int kChunks>
__device__ __forceinline__ void epilogue_tile_bf16_tma(const GemmParams& p, const CUtensorMap* tm_out,
                                                       const CUtensorMap* tm_out2, uint8_t* stg, uint64_t* aux_bar,
                                                       uint32_t aux_phase, uint32_t tmem_addr, uint64_t* tfull,
                                                       uint32_t tfull_phase, int row0, int col0, bool first_split,
                                                       int lane, float& cs0, float& cs1) {
  constexpr int W = kChunks * 16;
  const bool add_bias = p.bias != nullptr && first_split;
  const int npass = p.out2 != nullptr ? 2 : 1;
  u32x8 bs[2][2];
  auto prefetch = [&](int c, int buf) {
    const int n = col0 + c * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { bs[buf][0].v[j] = 0u; bs[buf][1].v[j] = 0u; }
    if (add_bias && n < p.N) {
      bs[buf][0] = ld_global_nc_256(p.bias + n);
      bs[buf][1] = ld_global_nc_256(p.bias + n + 8);
    }
  };
  prefetch(0, 0);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (p.aux_mode != 0) mbar_wait(aux_bar, aux_phase);
  uint32_t v[16];
  int g = 0;   // running chunk counter over the passes (bias double buffer)
  for (int pass = 0; pass < npass; ++pass) {
    const bool final_pass = pass == npass - 1;
    tmem_ld16(tmem_addr, v);
#pragma unroll
    for (int c = 0; c < kChunks; ++c, ++g) {
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
      if (c + 1 < kChunks) tmem_ld16(tmem_addr + (c + 1) * 16, v);
      if (c + 1 < kChunks || !final_pass) prefetch((c + 1) % kChunks, (g + 1) & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] += __uint_as_float(bs[g & 1][0].v[j]);
        f[8 + j] += __uint_as_float(bs[g & 1][1].v[j]);
      }
      uint8_t* s0 = stg + stg_off<W>(lane, 2 * c);
      uint8_t* s1 = stg + stg_off<W>(lane, 2 * c + 1);
      if (final_pass) {
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
        }
        if (p.aux_mode != 0) {
          const uint4 a0 = *reinterpret_cast<const uint4*>(s0);
          const uint4 a1 = *reinterpret_cast<const uint4*>(s1);
          const uint32_t ax[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          if (p.aux_mode == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 x = unpack_bf16(ax[j]);
              f[2 * j] *= gelu_erf_grad(x.x);
              f[2 * j + 1] *= gelu_erf_grad(x.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 x = unpack_bf16(ax[j]);
              f[2 * j] = x.x > 0.0f ? f[2 * j] : 0.0f;
              f[2 * j + 1] = x.y > 0.0f ? f[2 * j + 1] : 0.0f;
            }
          }
        }
      }
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
      if (c == 0 && p.aux_mode == 0) {
        // the previous store out of this buffer must have finished READING it (with aux, the lane that re-filled the
        // buffer already waited for that)
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
      *reinterpret_cast<uint4*>(s0) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(s1) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0 && row0 < p.M && col0 < p.N) {
      tma_store_2d(final_pass ? tm_out : tm_out2, stg, col0, row0);
      bulk_commit();
    }
    if (final_pass && p.colsum != nullptr && col0 < p.N) {
      // bias gradient of the producing Linear: column sums of this warp's [32 x W] tile, read back from the staged
      // shared-memory copy (i.e. of the bf16-rounded values, exactly what a separate pass over the output would
      // sum).  Lane l owns the 32-bit word l of every 128-byte row (2 columns) — 32 conflict-free LDS (the swizzle
      // permutes 16-byte units inside a row, a row still covers all banks); rows past M hold padding and are skipped.
      // The TMA store of the tile has already been issued (both only read it).  One coalesced red per warp and tile;
      // replaces a 77 MB re-read of the fc1 hidden gradient per layer (r01 / r02 step tables: 12 colsum launches of 19 us).
      const int nrows = min(32, p.M - row0);
      constexpr int kWords = W / 2;            // 32-bit words per row
      if (lane < kWords) {
        float sa[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sb[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // 4 independent chains per column
#pragma unroll
        for (int r4 = 0; r4 < 32; r4 += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = r4 + u;
            if (rr < nrows) {
              const uint32_t wv = *reinterpret_cast<const uint32_t*>(stg + stg_off<W>(rr, lane >> 2) + ((lane & 3) << 2));
              const float2 f2 = unpack_bf16(wv);
              sa[u] += f2.x;
              sb[u] += f2.y;
            }
          }
        }
        const float s0 = (sa[0] + sa[1]) + (sa[2] + sa[3]), s1 = (sb[0] + sb[1]) + (sb[2] + sb[3]);
        if (p.raster) {   // same columns for every tile of this warp: keep the sums, one reduction at the end of the kernel
          cs0 += s0;
          cs1 += s1;
        } else {
          const int n = col0 + 2 * lane;
          if (n < p.N) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.colsum + n), "f"(s0) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.colsum + n + 1), "f"(s1) : "memory");
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// fp32 epilogue through shared memory + TMA (EPI = 2, 128-column tiles: a warp's 32 columns are one 128-byte row).
// out = dropout(act(acc + bias)) + residual.  The register-direct / transposed form reads the fp32 residual with LSU
// loads issued one 16-column chunk ahead of their use: with K = 768 the residual stream, not the MMA, bounded the
// out-projection GEMM (r02 breakdown: 40 us for 15 GFLOP = 369 TFLOP/s).  Here the residual tile of the NEXT work item is
// already in shared memory when its accumulator completes (TMA load issued one tile ahead into the other of two
// buffers); thread = row adds its 32 values in place (swizzled 16-byte accesses, conflict-free) and one TMA store writes
// the tile.
// ------------------------------------------------------------------------------------------------
template <int kChunks>
__device__ __forceinline__ void epilogue_tile_f32_tma(const GemmParams& p, const CUtensorMap* tm_out, uint8_t* buf,
                                                      uint64_t* res_bar, uint32_t res_phase, uint32_t tmem_addr,
                                                      uint64_t* tfull, uint32_t tfull_phase, int row0, int col0,
                                                      int lane) {
  static_assert(kChunks == 2, "EPI = 2 is defined for 128-column tiles (32 fp32 columns per warp)");
  const bool add_bias = p.bias != nullptr;
  const bool use_res = p.residual != nullptr;
  u32x8 bs[2][2];
  auto prefetch = [&](int c, int b) {
    const int n = col0 + c * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { bs[b][0].v[j] = 0u; bs[b][1].v[j] = 0u; }
    if (add_bias && n < p.N) {
      bs[b][0] = ld_global_nc_256(p.bias + n);
      bs[b][1] = ld_global_nc_256(p.bias + n + 8);
    }
  };
  prefetch(0, 0);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (use_res) {
    mbar_wait(res_bar, res_phase);
  } else {
    if (lane == 0) bulk_wait_read1();   // this buffer's previous store (two tiles ago) has drained
    __syncwarp();
  }
  uint32_t v[16];
  tmem_ld16(tmem_addr, v);
  const uint32_t dseed = p.drop.thresh != 0u ? drop_seed(p.drop) : 0u;
  const int r = row0 + lane;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int n = col0 + c * 16;
    tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    if (c + 1 < kChunks) {
      tmem_ld16(tmem_addr + (c + 1) * 16, v);
      prefetch(c + 1, (c + 1) & 1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] += __uint_as_float(bs[c & 1][0].v[j]);
      f[8 + j] += __uint_as_float(bs[c & 1][1].v[j]);
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
    } else if (p.act == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
    if (p.drop.thresh != 0u) {   // TF:267 / TF:310: dropout(dense(x)) then + residual; same element -> mask map as EPI 0
      const uint32_t e = (uint32_t)r * (uint32_t)p.N + (uint32_t)n;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bool k0, k1;
        drop_keep2(e + 2 * j, dseed, p.drop.thresh, k0, k1);
        f[2 * j] = k0 ? f[2 * j] * p.drop.scale : 0.0f;
        f[2 * j + 1] = k1 ? f[2 * j + 1] * p.drop.scale : 0.0f;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4* slot = reinterpret_cast<float4*>(buf + lane * 128 + (((4 * c + q) ^ (lane & 7)) << 4));
      float4 o = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      if (use_res) {
        const float4 rr = *slot;
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      *slot = o;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && row0 < p.M && col0 < p.N) {
    tma_store_2d(tm_out, buf, col0, row0);
    bulk_commit();
  }
}

template <int BN, int A_MN, int B_MN, int CTA2, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
            const __grid_constant__ CUtensorMap tmap_aux, const GemmParams p) {
  using L = Cfg<BN, CTA2, EPI>;
  constexpr int kStages = L::kStages;
  constexpr int kNCta = CTA2 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;           // one per epilogue warp (EPI = 1 with an aux operand)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2 * kEpiWarps);
  SchedSmem* ss = reinterpret_cast<SchedSmem*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (EPI) {
      tma_prefetch_desc(&tmap_out);
      if (p.out2 != nullptr) tma_prefetch_desc(&tmap_out2);
      if (p.aux_mode != 0 || (EPI == 2 && p.residual != nullptr)) tma_prefetch_desc(&tmap_aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps * kNCta);
    }
    if (EPI) {
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
    }
    for (int i = 0; i < kSched; ++i) {
      mbar_init(&ss->full[i], 1);
      // consumers per cluster: leader = producer + MMA + epilogue warps + the scheduler itself; peer = producer + epilogue
      mbar_init(&ss->empty[i], (kEpiWarps + 3) + (CTA2 ? kEpiWarps + 1 : 0));
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc2(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is on-chip prologue (barriers, TMEM, descriptor prefetch) and overlaps the previous kernel's tail
  // under programmatic dependent launch; no global memory is touched before this point
  pdl_wait();
  pdl_trigger();

  const int tiles = p.tiles_m * p.tiles_n;
  const int total_work = p.tail_split > 1 ? p.tail_from + (tiles - p.tail_from) * p.tail_split : tiles * p.splits;
  const int kb_per_split = (p.kblocks + p.splits - 1) / p.splits;
  const int unit = blockIdx.x / kNCta;
  const int nunits = gridDim.x / kNCta;
  TileIter<CTA2> ti{ss, 0, nunits, total_work, rank, p.dynamic != 0, p.raster ? p.per : 0, p.tiles_m, p.tiles_n};

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (each CTA loads its A rows and its B half)
    constexpr uint32_t kBytesA = L::kABytes;
    constexpr uint32_t kBytesB = B_MN ? L::kBBoxes * 64 * BK * 2 : L::kBNH * BK * 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = unit; w >= 0; w = ti.next(w, lane)) {
      int split, tm, tn, sub;
      decode_work(p, w, tiles, nunits, split, tm, tn, sub);
      const int m0 = tm * (BM * kNCta) + rank * BM;
      // sub-tile: the same boxes, loaded at the sub-tile's columns; the MMA reads only the first rows of each half
      const int n0 = sub < 0 ? tn * BN + rank * L::kBNH
                             : tn * BN + sub * (BN / p.tail_split) + rank * (BN / p.tail_split / kNCta);
      const int kb0 = split * kb_per_split;
      const int kb1 = min(p.kblocks, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], (kBytesA + kBytesB) * kNCta);
          const int k0 = kb * BK;
          if (A_MN == 0) {
            if (CTA2) tma_load_2d_2cta(sa, &tmap_a, &full_bar[stage], k0, m0);
            else tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              if (CTA2) tma_load_2d_2cta(sa + i * (BK * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
              else tma_load_2d(sa + i * (BK * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
            }
          }
          if (B_MN == 0) {
            if (CTA2) tma_load_2d_2cta(sb, &tmap_b, &full_bar[stage], k0, n0);
            else tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < L::kBBoxes; ++i) {
              if (CTA2) tma_load_2d_2cta(sb + i * (BK * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
              else tma_load_2d(sb + i * (BK * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA of the pair only)
    if (leader) {
      const uint32_t idesc_full = umma_idesc_bf16(BM * kNCta, BN, A_MN, B_MN);
      const uint32_t idesc_sub = umma_idesc_bf16(BM * kNCta, BN / (p.tail_split > 1 ? p.tail_split : 1), A_MN, B_MN);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = unit; w >= 0; w = ti.next(w, lane)) {
        int split, tm_unused, tn_unused, sub;
        decode_work(p, w, tiles, nunits, split, tm_unused, tn_unused, sub);
        const uint32_t idesc = sub < 0 ? idesc_full : idesc_sub;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * L::kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * L::kStageBytes;
            const uint32_t sb = sa + L::kABytes;
            const uint64_t adesc = A_MN ? umma_desc_sw128(sa, BK * 128, 1024) : umma_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(sb, BK * 128, 1024) : umma_desc_sw128(sb, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = adesc + (A_MN ? (uint64_t)(k * 128) : (uint64_t)(k * 2));
              const uint64_t bd = bdesc + (B_MN ? (uint64_t)(k * 128) : (uint64_t)(k * 2));
              if (CTA2) umma_bf16_2cta(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (CTA2) {
              umma_commit_2cta(&empty_bar[stage]);                    // frees the slot in both CTAs
              if (kb == kb1 - 1) umma_commit_2cta(&tfull_bar[acc]);   // accumulators complete -> both epilogues
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ tile scheduler (leader CTA, one lane): cluster launch control
    if (leader && p.dynamic && lane == 0) {
      // up to kSched - 1 queries in flight: the answer to a try_cancel takes longer than one short tile
      int issued = 0, consumed = 0;
      bool more = true;
      for (;;) {
        while (more && issued - consumed < kSched - 1) {
          const int slot = issued & (kSched - 1);
          if (issued >= kSched) mbar_wait(&ss->empty[slot], (uint32_t)(issued / kSched - 1) & 1u);
          mbar_expect_tx(&ss->full[slot], 16);
          if (CTA2) mbar_expect_tx_remote(&ss->full[slot], 1, 16);
          clc_try_cancel<CTA2>(&ss->resp[slot], &ss->full[slot]);
          ++issued;
        }
        if (consumed == issued) break;
        const int slot = consumed & (kSched - 1);
        mbar_wait(&ss->full[slot], (uint32_t)(consumed / kSched) & 1u);
        const int x = clc_decode(&ss->resp[slot]);
        fence_proxy_async_smem();
        mbar_arrive_relaxed(&ss->empty[slot]);
        ++consumed;
        if (x < 0) more = false;   // nothing left to cancel; the queries still in flight are drained before exit
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (this CTA's 128 rows of the tile)
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int slice = ew >> 2;            // which quarter of the BN columns
    constexpr int kColsPerWarp = BN / 4;
    constexpr int kChunks = kColsPerWarp / 16;
    uint8_t* stg = smem + L::kEpiOffset + ew * L::kEpiWarpBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t aux_phase = 0;
    // -> false when this warp's column slice lies outside the (sub-)tile of work item w
    auto tile_origin = [&](int w, int& row0, int& col0) -> bool {
      int split_unused, tm, tn, sub;
      decode_work(p, w, tiles, nunits, split_unused, tm, tn, sub);
      row0 = tm * (BM * kNCta) + rank * BM + quad * 32;
      col0 = tn * BN + (sub > 0 ? sub * (BN / p.tail_split) : 0) + slice * kColsPerWarp;
      return sub < 0 || slice * p.tail_split < 4;
    };
    auto load_aux = [&](int w) {   // lane 0: aux tile of work item w -> this warp's staging buffer
      int r0, c0;
      if (!tile_origin(w, r0, c0)) return;
      mbar_expect_tx(&aux_bar[ew], L::kEpiWarpBytes);
      tma_load_2d(stg, &tmap_aux, &aux_bar[ew], c0, r0);
    };
    auto load_res = [&](int w, int b) {   // lane 0 (EPI = 2): fp32 residual tile of work item w -> buffer b
      int r0, c0;
      if (!tile_origin(w, r0, c0)) return;
      mbar_expect_tx(&aux_bar[2 * ew + b], 4096);
      tma_load_2d(stg + b * 4096, &tmap_aux, &aux_bar[2 * ew + b], c0, r0);
    };
    const bool res_tma = EPI == 2 && p.residual != nullptr;
    if (EPI == 1 && p.aux_mode != 0 && lane == 0) load_aux(unit);
    if (res_tma && lane == 0) load_res(unit, 0);
    int w_next = ti.next(unit, lane);   // one work item ahead: the aux / residual tile of the next item is prefetched
    int nact = 0;                       // work items in which this warp had columns to process
    float cs0 = 0.0f, cs1 = 0.0f;   // raster = 1: this lane's two column sums, carried across the warp's tiles
    int cs_col0 = -1;
    for (int w = unit; w >= 0; w = w_next, w_next = (w >= 0 ? ti.next(w, lane) : -1)) {
      const int split = (p.raster || p.tail_split > 1) ? 0 : w / tiles;
      int row0, col0;
      const bool active = tile_origin(w, row0, col0);
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * L::kAccStride + slice * kColsPerWarp);
      if (!active) {
        // column slice beyond a sub-tile: nothing to read, but the accumulator stage is released only once its MMAs
        // have completed (an early arrival could otherwise be counted in the NEXT phase of tempty)
        mbar_wait(&tfull_bar[acc], acc_phase);
      } else if constexpr (EPI == 2) {
        cs_col0 = col0;
        const int b = nact & 1;
        if (res_tma && lane == 0 && w_next >= 0) {
          bulk_wait_read0();            // the previous tile's store out of the other buffer has drained it
          load_res(w_next, b ^ 1);      // lands while this tile is processed
        }
        epilogue_tile_f32_tma<kChunks>(p, &tmap_out, stg + b * 4096, &aux_bar[2 * ew + b], (uint32_t)(nact >> 1) & 1u, taddr,
                                       &tfull_bar[acc], acc_phase, row0, col0, lane);
      } else if constexpr (EPI == 1) {
        cs_col0 = col0;
        epilogue_tile_bf16_tma<kChunks>(p, &tmap_out, &tmap_out2, stg, &aux_bar[ew], aux_phase, taddr, &tfull_bar[acc],
                                        acc_phase, row0, col0, split == 0, lane, cs0, cs1);
      } else if (p.out_f32) {
        epilogue_tile_f32<kChunks>(p, stg, taddr, &tfull_bar[acc], acc_phase, row0, col0, split == 0, lane);
      } else {
        epilogue_tile_bf16<kChunks>(p, taddr, &tfull_bar[acc], acc_phase, row0, col0, split == 0, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2 && !leader) mbar_arrive_remote(&tempty_bar[acc], 0);
        else mbar_arrive_relaxed(&tempty_bar[acc]);
      }
      if (EPI == 1 && p.aux_mode != 0) {
        if (active) aux_phase ^= 1;
        if (lane == 0 && w_next >= 0) {
          bulk_wait_read0();          // the store issued above has drained the buffer
          load_aux(w_next);           // lands while the MMA warp works on the next tile
        }
      }
      if (active) ++nact;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI == 1 && p.colsum != nullptr && p.raster && cs_col0 >= 0 && lane < kColsPerWarp / 2) {
      const int n = cs_col0 + 2 * lane;
      if (n < p.N) {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.colsum + n), "f"(cs0) : "memory");
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.colsum + n + 1), "f"(cs1) : "memory");
      }
    }
    if (EPI && lane == 0) bulk_wait_read0();   // shared memory must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 2) {
    if (CTA2) tmem_dealloc2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
struct OutMaps {
  CUtensorMap out, out2, aux;
};

template <int BN, int A_MN, int B_MN, int CTA2, int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& om, const GemmParams& p, int grid,
                  cudaStream_t st) {
  auto kfn = gemm_kernel<BN, A_MN, B_MN, CTA2, EPI>;
  using L = Cfg<BN, CTA2, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kfn, ta, tb, om.out, om.out2, om.aux, p));
  return 0;
}

template <int BN, int CTA2, int EPI>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& om,
                        const GemmParams& p, int grid, cudaStream_t st) {
  switch ((a_mn ? 2 : 0) | (b_mn ? 1 : 0)) {
    case 0: return launch<BN, 0, 0, CTA2, EPI>(ta, tb, om, p, grid, st);
    case 1: return launch<BN, 0, 1, CTA2, EPI>(ta, tb, om, p, grid, st);
    case 2: return launch<BN, 1, 0, CTA2, EPI>(ta, tb, om, p, grid, st);
    default: return launch<BN, 1, 1, CTA2, EPI>(ta, tb, om, p, grid, st);
  }
}

template <int BN, int CTA2>
static int launch_epi(int epi, int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& om,
                      const GemmParams& p, int grid, cudaStream_t st) {
  if (epi == 1) return launch_major<BN, CTA2, 1>(a_mn, b_mn, ta, tb, om, p, grid, st);
  if constexpr (BN == 128) {
    if (epi == 2) return launch_major<BN, CTA2, 2>(a_mn, b_mn, ta, tb, om, p, grid, st);
  }
  return launch_major<BN, CTA2, 0>(a_mn, b_mn, ta, tb, om, p, grid, st);
}

struct TileChoice {
  int cta2, bn, splits;
};

// Number of column sub-tiles (1, 2 or 4) the `rem` tiles of the partial last wave are cut into (decode_work): sub-tile
// widths 128 / 96 / 64; an MN-major B operand is staged in 64-column boxes, so a CTA's share must be whole boxes; every
// unit gets at most one sub-tile (rem * S <= units).
static bool tail_mn_partial() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VS_GEMM_TAIL_MN");   // experiment: let the MMA read part of a 64-column MN-major box
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static int tail_split_for(int bn, int cta2, int b_mn, int rem, int units, int max_s) {
  if (rem <= 0) return 1;
  for (int S = max_s; S >= 2; S >>= 1) {
    const int wsub = bn / S, per_cta = wsub / (cta2 ? 2 : 1);
    if (wsub != 128 && wsub != 96 && wsub != 64) continue;
    if (b_mn && per_cta % 64 != 0 && !tail_mn_partial()) continue;
    if (rem * S > units) continue;
    return S;
  }
  return 1;
}
static int tail_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VS_GEMM_TAIL");   // 0: off (A/B runs), 2: halves only, default 4
    v = e ? atoi(e) : 4;
    if (v != 0 && v != 2) v = 4;
  }
  return v;
}

// Cost model: a 256xBN pair tile (cta_group::2) and a 128xBN single-CTA tile take about the same time per k-block
// (the single CTA runs at half the tensor rate), so cost ~ waves * BN * k-blocks-per-work-item.
static TileChoice choose_tiles(int M, int N, int kblocks, int nsm, bool allow_split, int forced_split, int forced_cfg,
                               bool allow_192, int b_mn) {
  TileChoice best{0, 256, 1};
  double best_cost = 1e30;
  const int cand[5][2] = {{1, 256}, {1, 192}, {1, 128}, {0, 256}, {0, 128}};
  for (int c = 0; c < 5; ++c) {
    if (forced_cfg >= 0 && c != forced_cfg) continue;
    if (forced_cfg < 0 && !allow_192 && cand[c][1] == 192) continue;
    const int cta2 = cand[c][0], bn = cand[c][1];
    if (bn > 128 && N <= 128 && forced_cfg < 0) continue;
    const int tm = (M + (cta2 ? 255 : 127)) / (cta2 ? 256 : 128);
    const int tn = (N + bn - 1) / bn;
    const int units = cta2 ? nsm / 2 : nsm;
    const int max_s = forced_split > 0 ? forced_split : (allow_split ? 16 : 1);
    for (int s = (forced_split > 0 ? forced_split : 1); s <= max_s; ++s) {
      if (s > 1 && kblocks / s < 8 && forced_split <= 0) break;
      const int work = tm * tn * s;
      const int waves = (work + units - 1) / units;
      const int kb = (kblocks + s - 1) / s;
      // per-work-item time ~ columns x (k-blocks + fill/drain overhead); 128-wide tiles re-read A more often.
      // Constants fitted to the measured table of tools/gemm_bench.py over the ViT-B/16 shapes x 5 configurations
      // (r01: the first constants, overhead 6 / no width penalty, mis-picked pair128 for the QKV projection
      // (47.3 vs 42.1 us) and for the fc2 weight gradient (52.2 vs 48.2 us)).
      double weff = double(waves);
      if (s == 1 && tail_env() && work % units != 0) {   // a split last wave costs 1/S of a tile time (+ its fill / drain)
        const int S = tail_split_for(bn, cta2, b_mn, work % units, units, tail_env());
        if (S > 1) weff = double(work / units) + 1.0 / S + 0.1;
      }
      double cost = weff * bn * (bn == 128 ? 1.05 : 1.0) * (kb + 1.5);
      if (!cta2) cost *= 1.02;  // prefer pairs on ties (less L2 traffic)
      if (cost < best_cost) { best_cost = cost; best = {cta2, bn, s}; }
    }
  }
  return best;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_gemm_bf16(const vs_gemm_desc* d, void* stream) {
  VS_CHECK_ARG(d != nullptr, "vs_gemm_bf16: null descriptor");
  VS_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "vs_gemm_bf16: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
  VS_CHECK_ARG(d->N % 16 == 0, "vs_gemm_bf16: N=%d must be a multiple of 16 (epilogue chunks)", d->N);
  VS_CHECK_ARG(d->A && d->B && d->out, "vs_gemm_bf16: null operand");
  VS_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "vs_gemm_bf16: lda/ldb must be multiples of 8 elements");
  VS_CHECK_ARG(((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0) && ((uintptr_t)d->out % 16 == 0),
               "vs_gemm_bf16: operands must be 16-byte aligned");
  VS_CHECK_ARG(d->ldo % (d->out_dtype ? 4 : 16) == 0, "vs_gemm_bf16: ldo alignment (bf16 rows must be 32-byte multiples)");
  VS_CHECK_ARG(d->out_dtype == 1 || (uintptr_t)d->out % 32 == 0, "vs_gemm_bf16: bf16 output must be 32-byte aligned");
  VS_CHECK_ARG(d->bias == nullptr || (uintptr_t)d->bias % 32 == 0, "vs_gemm_bf16: bias must be 32-byte aligned");
  VS_CHECK_ARG(!d->accumulate || d->out_dtype == 1, "vs_gemm_bf16: accumulate requires fp32 output");
  VS_CHECK_ARG(d->act >= 0 && d->act <= 2 && d->aux_mode >= 0 && d->aux_mode <= 2, "vs_gemm_bf16: bad act/aux_mode");
  VS_CHECK_ARG(d->aux_mode == 0 || (d->aux != nullptr && d->ldaux % 16 == 0 && (uintptr_t)d->aux % 32 == 0),
               "vs_gemm_bf16: aux missing/misaligned (32-byte rows)");
  VS_CHECK_ARG(d->out2 == nullptr || (d->ldo2 % 16 == 0 && (uintptr_t)d->out2 % 32 == 0), "vs_gemm_bf16: out2 alignment");
  VS_CHECK_ARG(d->residual == nullptr || d->ldr % 4 == 0, "vs_gemm_bf16: ldr alignment");
  VS_CHECK_ARG(d->split_k <= 1 || d->accumulate, "vs_gemm_bf16: split_k > 1 requires accumulate");
  VS_CHECK_ARG(d->out_colsum == nullptr || d->out_dtype == 0, "vs_gemm_bf16: out_colsum is defined for bf16 outputs");

  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_gemm_bf16: no CUDA device");

  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.kblocks = (d->K + BK - 1) / BK;
  // tile_cfg: 0 automatic; 1..5 force {pair256, pair192, pair128, single256, single128} (tests / tuning)
  int forced_cfg = (d->tile_cfg >= 1 && d->tile_cfg <= 5) ? d->tile_cfg - 1 : -1;
  if (forced_cfg < 0) {
    // tuning knob: VS_GEMM_TILE="M,N,K,a_mn,b_mn:cfg;..." forces a tile configuration (1..5) for the listed shapes, so
    // that candidates can be compared inside the real training step rather than in an isolated micro-benchmark
    static int n_over = -1;
    static int over[16][6];
    if (n_over < 0) {
      n_over = 0;
      const char* e = getenv("VS_GEMM_TILE");
      while (e && *e && n_over < 16) {
        int v[6];
        if (sscanf(e, "%d,%d,%d,%d,%d:%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6) {
          for (int i = 0; i < 6; ++i) over[n_over][i] = v[i];
          ++n_over;
        }
        e = strchr(e, ';');
        if (e) ++e;
      }
    }
    for (int i = 0; i < n_over; ++i)
      if (over[i][0] == d->M && over[i][1] == d->N && over[i][2] == d->K && over[i][3] == (d->a_mn_major != 0) &&
          over[i][4] == (d->b_mn_major != 0) && over[i][5] >= 1 && over[i][5] <= 5)
        forced_cfg = over[i][5] - 1;
  }
  // bf16 outputs leave through shared memory + TMA stores (EPI = 1) on the 256- and 128-column tiles; VS_GEMM_EPI=direct
  // keeps the register-direct stores (A/B comparisons), VS_GEMM_EPI=tma is the default
  static int epi_env = -1;
  if (epi_env < 0) {
    const char* e = getenv("VS_GEMM_EPI");
    epi_env = (e && strcmp(e, "direct") == 0) ? 0 : 1;
  }
  const bool want_tma = epi_env == 1 && d->out_dtype == 0;
  TileChoice tc = choose_tiles(d->M, d->N, p.kblocks, nsm, d->accumulate != 0, d->split_k, forced_cfg, !want_tma,
                                d->b_mn_major != 0);
  // fp32 outputs (+ residual) of the 128-column tiles: TMA residual prefetch + TMA store (EPI = 2)
  const bool f32_tma = epi_env == 1 && d->out_dtype == 1 && !d->accumulate && d->row_tokens == 0 && tc.bn == 128 &&
                       d->aux_mode == 0 && d->out2 == nullptr && d->ldo % 4 == 0 && (d->residual == nullptr || d->ldr % 4 == 0);
  const int epi = (want_tma && tc.bn != 192) ? 1 : (f32_tma ? 2 : 0);
  // fused into the staged-tile epilogue; otherwise (or with VS_GEMM_COLSUM=separate, for A/B runs) a pass after the GEMM
  static int colsum_fused = -1;
  if (colsum_fused < 0) {
    const char* e = getenv("VS_GEMM_COLSUM");
    colsum_fused = (e && strcmp(e, "separate") == 0) ? 0 : 1;
  }
  p.colsum = (epi == 1 && colsum_fused) ? d->out_colsum : nullptr;
  int splits = tc.splits;
  if (splits > p.kblocks) splits = p.kblocks;
  // every split must own at least one k-block
  while (splits > 1 && (splits - 1) * ((p.kblocks + splits - 1) / splits) >= p.kblocks) --splits;
  const int BN = tc.bn;
  p.tiles_m = (d->M + (tc.cta2 ? 255 : 127)) / (tc.cta2 ? 256 : 128);
  p.tiles_n = (d->N + BN - 1) / BN;
  p.splits = splits;
  p.out = d->out; p.ldo = d->ldo; p.out_f32 = d->out_dtype; p.accumulate = d->accumulate;
  p.bias = d->bias; p.act = d->act;
  p.out2 = (__nv_bfloat16*)d->out2; p.ldo2 = d->ldo2;
  p.aux = (const __nv_bfloat16*)d->aux; p.ldaux = d->ldaux; p.aux_mode = d->aux_mode;
  p.residual = d->residual; p.ldr = d->ldr; p.row_tokens = d->row_tokens;
  p.drop.thresh = 0; p.drop.scale = 1.0f; p.drop.seed = nullptr; p.drop.site = 0;
  if (d->dropout_p > 0.0f) {
    VS_CHECK_ARG(d->dropout_p < 1.0f && d->dropout_seed != nullptr && d->out_dtype == 1 && !d->accumulate,
                 "vs_gemm_bf16: dropout needs 0<p<1, a device seed pointer and a non-accumulating fp32 output");
    VS_CHECK_ARG((long long)d->M * d->N < (1LL << 32), "vs_gemm_bf16: dropout index space exceeds 2^32");
    p.drop.thresh = (uint32_t)(d->dropout_p * 65536.0f + 0.5f);
    p.drop.scale = 1.0f / (1.0f - (float)p.drop.thresh / 65536.0f);
    p.drop.seed = d->dropout_seed;
    p.drop.site = d->dropout_site;
  }

  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!d->a_mn_major) { dims[0] = d->K; dims[1] = d->M; box[0] = BK; box[1] = BM; }
    else                { dims[0] = d->M; dims[1] = d->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)d->lda * 2;
    int rc = make_tmap_bf16(&ta, d->A, 2, dims, strides, box);
    if (rc) return rc;
    const int bnh = tc.cta2 ? BN / 2 : BN;
    if (!d->b_mn_major) { dims[0] = d->K; dims[1] = d->N; box[0] = BK; box[1] = bnh; }
    else                { dims[0] = d->N; dims[1] = d->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)d->ldb * 2;
    rc = make_tmap_bf16(&tb, d->B, 2, dims, strides, box);
    if (rc) return rc;
  }
  OutMaps om;
  om.out = ta; om.out2 = ta; om.aux = ta;   // placeholders (never dereferenced) unless EPI = 1
  if (epi == 2) {
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M}, strides[1];
    uint32_t box[2] = {32u, 32u};
    strides[0] = (uint64_t)d->ldo * 4;
    int rc = make_tmap_sw(&om.out, d->out, 4, 2, dims, strides, box, 128);
    if (rc) return rc;
    if (d->residual) {
      strides[0] = (uint64_t)d->ldr * 4;
      rc = make_tmap_sw(&om.aux, d->residual, 4, 2, dims, strides, box, 128);
      if (rc) return rc;
    }
  } else if (epi) {
    // one [32 rows x BN/4 columns] box per epilogue warp; 128-byte rows (BN = 256) or 64-byte rows (BN = 128)
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M}, strides[1];
    uint32_t box[2] = {(uint32_t)(BN / 4), 32u};
    const int sw = BN == 256 ? 128 : 64;
    strides[0] = (uint64_t)d->ldo * 2;
    int rc = make_tmap_bf16_sw(&om.out, d->out, 2, dims, strides, box, sw);
    if (rc) return rc;
    if (d->out2) {
      strides[0] = (uint64_t)d->ldo2 * 2;
      rc = make_tmap_bf16_sw(&om.out2, d->out2, 2, dims, strides, box, sw);
      if (rc) return rc;
    }
    if (d->aux_mode) {
      strides[0] = (uint64_t)d->ldaux * 2;
      rc = make_tmap_bf16_sw(&om.aux, d->aux, 2, dims, strides, box, sw);
      if (rc) return rc;
    }
  }
  const int total = p.tiles_m * p.tiles_n * splits;
  cudaStream_t st = (cudaStream_t)stream;
  // VS_GEMM_SCHED=clc: one cluster per work item, tiles handed out by cluster launch control (adapts to the SMs the
  // kernel actually gets, e.g. next to NCCL's CTAs); VS_GEMM_SCHED=static: persistent grid, fixed round-robin assignment
  static int sched_env = -1;
  if (sched_env < 0) {
    const char* e = getenv("VS_GEMM_SCHED");
    sched_env = (e && strcmp(e, "clc") == 0) ? 1 : 0;
  }
  p.dynamic = sched_env;
  // column-persistent tile order when the kernel also sums the columns of its output (decode_work): needs at least one
  // unit per column tile and a single k split; VS_GEMM_RASTER=0 keeps the default order (one reduction per warp and tile)
  p.raster = 0;
  p.per = 0;
  int units_eff = 0;
  {
    static int raster_env = -1;
    if (raster_env < 0) {
      const char* e = getenv("VS_GEMM_RASTER");
      raster_env = (e && e[0] == '0') ? 0 : 1;
    }
    const int units = tc.cta2 ? nsm / 2 : nsm;
    if (p.colsum != nullptr && raster_env && !p.dynamic && splits == 1 && p.tiles_n <= units) {
      int per = units / p.tiles_n;
      if (per > p.tiles_m) per = p.tiles_m;
      p.raster = 1;
      p.per = per;
      units_eff = per * p.tiles_n;
    }
  }
  // partial last wave of the static schedule -> narrower work items (decode_work).  Every unit gets at most ONE sub-tile,
  // as its last item (rem * S <= units), which the epilogue's prefetch logic relies on.  Sub-tile widths 128 / 96 / 64;
  // an MN-major B operand is staged in 64-column boxes, so a CTA's share of the sub-tile must be whole boxes
  // (tail_split_for).  VS_GEMM_TAIL=0 switches it off (A/B runs), =2 limits the split to halves.
  p.tail_from = total;
  p.tail_split = 1;
  int total_items = total;
  {
    const int units = tc.cta2 ? nsm / 2 : nsm;
    const int tiles = p.tiles_m * p.tiles_n;
    const int rem = tiles % units;
    if (tail_env() && splits == 1 && !p.dynamic && !p.raster && rem > 0) {
      const int S = tail_split_for(BN, tc.cta2, d->b_mn_major != 0, rem, units, tail_env());
      if (S > 1) {
        p.tail_split = S;
        p.tail_from = tiles - rem;
        total_items = p.tail_from + rem * S;
      }
    }
  }
  int rc = 0;
  if (tc.cta2) {
    const int pairs = nsm / 2;
    const int grid = p.raster ? 2 * units_eff : (p.dynamic ? 2 * total : 2 * (total_items < pairs ? total_items : pairs));
    if (BN == 256) rc = launch_epi<256, 1>(epi, d->a_mn_major, d->b_mn_major, ta, tb, om, p, grid, st);
    else if (BN == 192) rc = launch_major<192, 1, 0>(d->a_mn_major, d->b_mn_major, ta, tb, om, p, grid, st);
    else rc = launch_epi<128, 1>(epi, d->a_mn_major, d->b_mn_major, ta, tb, om, p, grid, st);
  } else {
    const int grid = p.raster ? units_eff : (p.dynamic ? total : (total_items < nsm ? total_items : nsm));
    if (BN == 256) rc = launch_epi<256, 0>(epi, d->a_mn_major, d->b_mn_major, ta, tb, om, p, grid, st);
    else rc = launch_epi<128, 0>(epi, d->a_mn_major, d->b_mn_major, ta, tb, om, p, grid, st);
  }
  if (rc == 0 && d->out_colsum != nullptr && p.colsum == nullptr)
    rc = vs_colsum_bf16(d->out, d->ldo, d->M, d->N, d->out_colsum, 1, stream);
  return rc;
}
