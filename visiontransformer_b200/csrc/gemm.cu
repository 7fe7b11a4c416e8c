// Persistent, warp-specialised bf16 GEMM on tcgen05 tensor cores for sm_100a.
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// One CTA per SM loops over (m-tile, n-tile, k-split) work items.  Roles:
//   warp 0      : TMA producer  — streams 128xBK A tiles and BNxBK B tiles into a 4-stage 128B-swizzled smem ring
//   warp 1      : MMA issuer    — one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into TMEM accumulators
//   warp 2      : TMEM allocator (512 columns = two BN=256 fp32 accumulator stages)
//   warps 4..11 : epilogue      — tcgen05.ld the accumulator (thread = output row), fused bias / GELU / ReLU /
//                                 GELU' / residual / row-remap, bf16 or fp32 (or atomic fp32) stores
// The two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Both operands can be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the latter lets the
// dgrad (dY·W) and wgrad (dYᵀ·X) GEMMs of the backward pass read the forward tensors in place, without transposes
// (replaces autograd's mm_backward for TF:216-218,262,290,305).
#include "common.cuh"
#include "../../include/vitseg.h"

namespace vs {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kblocks;  // kblocks = ceil(K / BK)
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
};

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kAccStages = (BN <= 256) ? 2 : 1;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.tiles_m * p.tiles_n;
  const int total_work = tiles * p.splits;
  const int kb_per_split = (p.kblocks + p.splits - 1) / p.splits;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (whole warp converged; one elected lane issues,
    // so descriptors / coordinates stay in uniform registers)
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles;
      const int t = w - split * tiles;
      const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
      const int m0 = tm * BM, n0 = tn * BN;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(p.kblocks, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          const int k0 = kb * BK;
          if (A_MN == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d(sa + i * (BK * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * (BK * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (whole warp converged, elected lane issues:
    // a divergent single-thread loop made ptxas wrap every UTCHMMA in ELECT/R2UR sequences and the issue loop,
    // not the tensor pipe, became the limiter — ncu r01: tensor pipe 51 % active with the operands always ready)
    const uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(p.kblocks, kb0 + kb_per_split);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
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
            umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = ew >> 2;             // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles;
      const int t = w - split * tiles;
      const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
      const int m0 = tm * BM, n0 = tn * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int r = m0 + quad * 32 + lane;
      const bool row_ok = r < p.M;
      const long long orow = r;
      const long long rrow = (p.row_tokens > 0) ? (r % p.row_tokens) : r;
      const bool add_bias = (p.bias != nullptr) && (split == 0);
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        const int n = n0 + half * kColsPerWarp + c;
        if (n >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * BN + half * kColsPerWarp + c), v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (add_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        if (row_ok) {
          if (p.out2 != nullptr) {
            uint4* o2 = reinterpret_cast<uint4*>(p.out2 + orow * p.ldo2 + n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o2[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (p.aux_mode != 0) {
            const uint4* a4 = reinterpret_cast<const uint4*>(p.aux + (long long)r * p.ldaux + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 a = __ldg(a4 + j);
              const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 x = unpack_bf16(aw[q]);
                if (p.aux_mode == 1) {
                  f[8 * j + 2 * q] *= gelu_erf_grad(x.x);
                  f[8 * j + 2 * q + 1] *= gelu_erf_grad(x.y);
                } else {
                  f[8 * j + 2 * q] = x.x > 0.0f ? f[8 * j + 2 * q] : 0.0f;
                  f[8 * j + 2 * q + 1] = x.y > 0.0f ? f[8 * j + 2 * q + 1] : 0.0f;
                }
              }
            }
          }
          if (p.residual != nullptr && split == 0) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = r4[j];
              f[4 * j + 0] += x.x; f[4 * j + 1] += x.y; f[4 * j + 2] += x.z; f[4 * j + 3] += x.w;
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(f[4 * j]),
                             "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                             : "memory");
            } else {
              float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o4[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
template <int BN, int A_MN, int B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t st) {
  auto kfn = gemm_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN>::kTotal));
    attr_set = true;
  }
  kfn<<<grid, kThreads, SmemLayout<BN>::kTotal, st>>>(ta, tb, p);
  VS_CHECK_LAUNCH();
  return 0;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_gemm_bf16(const vs_gemm_desc* d, void* stream) {
  VS_CHECK_ARG(d != nullptr, "vs_gemm_bf16: null descriptor");
  VS_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "vs_gemm_bf16: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
  VS_CHECK_ARG(d->N % 32 == 0, "vs_gemm_bf16: N=%d must be a multiple of 32", d->N);
  VS_CHECK_ARG(d->A && d->B && d->out, "vs_gemm_bf16: null operand");
  VS_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "vs_gemm_bf16: lda/ldb must be multiples of 8 elements");
  VS_CHECK_ARG(((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0) && ((uintptr_t)d->out % 16 == 0),
               "vs_gemm_bf16: operands must be 16-byte aligned");
  VS_CHECK_ARG(d->ldo % (d->out_dtype ? 4 : 8) == 0, "vs_gemm_bf16: ldo alignment");
  VS_CHECK_ARG(!d->accumulate || d->out_dtype == 1, "vs_gemm_bf16: accumulate requires fp32 output");
  VS_CHECK_ARG(d->act >= 0 && d->act <= 2 && d->aux_mode >= 0 && d->aux_mode <= 2, "vs_gemm_bf16: bad act/aux_mode");
  VS_CHECK_ARG(d->aux_mode == 0 || (d->aux != nullptr && d->ldaux % 8 == 0), "vs_gemm_bf16: aux missing/misaligned");
  VS_CHECK_ARG(d->out2 == nullptr || d->ldo2 % 8 == 0, "vs_gemm_bf16: ldo2 alignment");
  VS_CHECK_ARG(d->residual == nullptr || d->ldr % 4 == 0, "vs_gemm_bf16: ldr alignment");

  const int nsm = sm_count();
  VS_CHECK_ARG(nsm > 0, "vs_gemm_bf16: no CUDA device");

  const int BN = (d->N <= 128) ? 128 : 256;
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.tiles_m = (d->M + BM - 1) / BM;
  p.tiles_n = (d->N + BN - 1) / BN;
  p.kblocks = (d->K + BK - 1) / BK;
  int splits = d->split_k;
  const int tiles = p.tiles_m * p.tiles_n;
  if (splits <= 0) {
    splits = 1;
    if (d->accumulate && tiles < nsm) {
      // pick the split count with the best wave efficiency, keeping >= 8 k-blocks per split
      double best = 0.0;
      for (int s = 1; s <= 16; ++s) {
        if (p.kblocks / s < 8 && s > 1) break;
        const int work = tiles * s;
        const double eff = double(work) / double(((work + nsm - 1) / nsm) * nsm);
        if (eff > best + 1e-9) { best = eff; splits = s; }
      }
    }
  }
  VS_CHECK_ARG(splits == 1 || d->accumulate, "vs_gemm_bf16: split_k > 1 requires accumulate");
  if (splits > p.kblocks) splits = p.kblocks;
  // every split must own at least one k-block
  while (splits > 1 && (splits - 1) * ((p.kblocks + splits - 1) / splits) >= p.kblocks) --splits;
  p.splits = splits;
  p.out = d->out; p.ldo = d->ldo; p.out_f32 = d->out_dtype; p.accumulate = d->accumulate;
  p.bias = d->bias; p.act = d->act;
  p.out2 = (__nv_bfloat16*)d->out2; p.ldo2 = d->ldo2;
  p.aux = (const __nv_bfloat16*)d->aux; p.ldaux = d->ldaux; p.aux_mode = d->aux_mode;
  p.residual = d->residual; p.ldr = d->ldr; p.row_tokens = d->row_tokens;

  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!d->a_mn_major) { dims[0] = d->K; dims[1] = d->M; box[0] = BK; box[1] = BM; }
    else                { dims[0] = d->M; dims[1] = d->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)d->lda * 2;
    int rc = make_tmap_bf16(&ta, d->A, 2, dims, strides, box);
    if (rc) return rc;
    if (!d->b_mn_major) { dims[0] = d->K; dims[1] = d->N; box[0] = BK; box[1] = BN; }
    else                { dims[0] = d->N; dims[1] = d->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)d->ldb * 2;
    rc = make_tmap_bf16(&tb, d->B, 2, dims, strides, box);
    if (rc) return rc;
  }
  const int total = tiles * splits;
  const int grid = total < nsm ? total : nsm;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = (BN == 256 ? 4 : 0) | (d->a_mn_major ? 2 : 0) | (d->b_mn_major ? 1 : 0);
  switch (key) {
    case 0: return launch<128, 0, 0>(ta, tb, p, grid, st);
    case 1: return launch<128, 0, 1>(ta, tb, p, grid, st);
    case 2: return launch<128, 1, 0>(ta, tb, p, grid, st);
    case 3: return launch<128, 1, 1>(ta, tb, p, grid, st);
    case 4: return launch<256, 0, 0>(ta, tb, p, grid, st);
    case 5: return launch<256, 0, 1>(ta, tb, p, grid, st);
    case 6: return launch<256, 1, 0>(ta, tb, p, grid, st);
    default: return launch<256, 1, 1>(ta, tb, p, grid, st);
  }
}
