// Shared device/host helpers for the sm_100a kernels of libvitseg.
// Thin inline-PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// UMMA shared-memory + instruction descriptors, and small numeric helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vs {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define VS_CHECK_ARG(cond, ...)                    \
  do {                                             \
    if (!(cond)) {                                 \
      vs::set_error(__VA_ARGS__);                  \
      return -1;                                   \
    }                                              \
  } while (0)
#define VS_CHECK_CUDA(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      vs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                              \
    }                                                                              \
  } while (0)
#define VS_CHECK_LAUNCH() VS_CHECK_CUDA(cudaGetLastError())

int sm_count();

// Build a TMA descriptor over a row-major bf16 tensor of up to 3 dims.
// dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dims[i+1].
// box[i] = box extent per dim; swizzle 128B; OOB elements are zero-filled.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// same with an explicit swizzle span in bytes (128, 64, 32 or 0 = none); box[0] * 2 must not exceed it
int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
// general form: elem_bytes 2 (bf16) or 4 (fp32)
int make_tmap_sw(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begins with pdl_wait() — griddepcontrol.wait: block until all
// prerequisite grids have COMPLETED and their memory is visible — followed by pdl_trigger(), which lets the next
// kernel of the stream be scheduled while this one is still running.  The next kernel's launch processing and its
// on-chip prologue (barrier init, TMEM allocation, tensor-map prefetch, placed BEFORE its own pdl_wait) then overlap
// this kernel's tail instead of forming a ~1-2 us bubble per launch (r01: 297 launches per training step).
// Correctness rule: a kernel touches global memory only after its own pdl_wait(), and every thread executes it
// before any early exit — then "B waited for A" holds transitively along the stream (C waits for all of B, all of B
// waited for all of A).  The attribute is set only with VS_PDL=1 (without it griddepcontrol.* are no-ops): inside the
// CUDA graph of the training step the programmatic edges measured 0.2 ms per step SLOWER than plain edges (r02).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// generic device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 256-bit global accesses (sm_100: LDG/STG.E.256).  One row per thread is the natural layout after tcgen05.ld, so a
// warp request touches 32 different lines; the LSU then costs one wavefront per (line, instruction).  Moving 32 B
// per instruction instead of 16 B halves the wavefronts per byte (gemm_epi_bench: the second bf16 output of fc1
// cost +31 us, the GELU' aux read +24 us with 128-bit accesses).  Addresses must be 32-byte aligned.
struct __align__(32) u32x8 { uint32_t v[8]; };
__device__ __forceinline__ void st_global_256(void* ptr, const u32x8& a) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]),
               "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
               : "memory");
}
__device__ __forceinline__ u32x8 ld_global_nc_256(const void* ptr) {
  u32x8 a;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.v[0]), "=r"(a.v[1]), "=r"(a.v[2]), "=r"(a.v[3]), "=r"(a.v[4]), "=r"(a.v[5]), "=r"(a.v[6]),
                 "=r"(a.v[7])
               : "l"(ptr));
  return a;
}

// exact-erf GELU (hidden_act="gelu", TF:290-299) and its derivative, fp32.
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. fp32-level): 2 MUFU ops (rcp, ex2) + ~8 FMAs per
// element instead of erff()'s ~30-instruction branchy path, which made the fc1 / fc2-dgrad GEMM epilogues slower
// than their MMA main loops.  The tail 0.5*(1+erf) is formed without cancellation: Phi(x<0) = 0.5*poly*e.
// half_tail(|x|) = 0.5 * erfc(|x| / sqrt(2)) = Phi(-|x|), and e = exp(-x^2/2).  Constants are pre-folded
// (p/sqrt(2) into the rcp argument, 0.5 into the polynomial, -0.5*log2(e) into the exponent): 11 instructions.
__device__ __forceinline__ float gelu_half_tail(float ax, float x, float& e) {
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164190f, ax, 1.0f)));          // 0.3275911 / sqrt(2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((x * x) * -0.72134752044448170368f));  // -0.5 * log2(e)
  float p = fmaf(0.5307027145f, t, -0.7265760135f);   // 0.5 * {1.061405429, -1.453152027, 1.421413741, ...}
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  e = ex;
  return (p * t) * ex;
}
// Forward GELU needs only the tail Phi(-|x|) = 2^Q(|x|): Q = log2(0.5 erfc(a / sqrt 2)) is smooth, and a degree-6
// polynomial (weighted minimax fit on [0, 6], weight a * Phi(-a): the quantity that reaches the output) gives
// |GELU error| <= 1.5e-7 and |Phi error| <= 2.6e-7 — the accuracy of the A-S form above with ONE MUFU op and 10
// instructions instead of 15 (the fc1 epilogue is bound by issue slots: K = 768 leaves ~24 instructions per element).
// Beyond a = 6 the tail is < 1e-9 and the argument is clamped (the polynomial is only valid on the fitted range).
__device__ __forceinline__ float gelu_erf(float x) {
  // x * Phi(x) = max(x, 0) - |x| * Phi(-|x|): no select, no cancellation in the tail
  const float ax = fabsf(x);
  const float a = fminf(ax, 6.0f);
  float q = fmaf(2.7676265744958073e-05f, a, -0.0007205978035926819f);
  q = fmaf(q, a, 0.007916657254099846f);
  q = fmaf(q, a, -0.053151555359363556f);
  q = fmaf(q, a, -0.4589695334434509f);
  q = fmaf(q, a, -1.1511367559432983f);
  q = fmaf(q, a, -0.9999993443489075f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
  return fmaf(-ax, h, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float h = gelu_half_tail(fabsf(x), x, e);
  const float cdf = x >= 0.0f ? 1.0f - h : h;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// ----------------------------------------------------------------------------------------------
// Counter-based dropout masks (hidden_dropout_prob / attention_probs_dropout_prob = 0.1, model/CE/classes.py:233-234).
// mask(site, element) = hash(element index, seed(step, site)): nothing is stored; backward regenerates the mask.
// A 32-bit murmur-style finaliser yields two 16-bit uniforms; keep iff u16 >= round(p * 65536).
// ----------------------------------------------------------------------------------------------
struct DropCfg {
  uint32_t thresh;        // 0 = dropout off
  float scale;            // 1 / (1 - p)
  const uint32_t* seed;   // device pointer: per-step counter (CUDA-graph replays advance it on the device)
  uint32_t site;
};
__device__ __forceinline__ uint32_t drop_hash(uint32_t x, uint32_t seed);
// (step, site) -> seed through two full avalanche rounds: seeds of neighbouring steps / sites share no linear relation
__device__ __forceinline__ uint32_t drop_seed(const DropCfg& d) {
  return drop_hash(d.site + 0x27D4EB2Fu, drop_hash(d.seed[0], 0x9E3779B9u));
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t x, uint32_t seed) {
  x ^= seed;
  x *= 0x9E3779B1u; x ^= x >> 16;
  x *= 0x85EBCA6Bu; x ^= x >> 13;
  x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
// per-element form (attention probabilities)
__device__ __forceinline__ bool drop_keep(uint32_t elem, uint32_t seed, uint32_t thresh) {
  return (drop_hash(elem, seed) & 0xFFFFu) >= thresh;
}
// pair form (hidden states): elements 2k and 2k+1 share one hash; `elem` must be even
__device__ __forceinline__ void drop_keep2(uint32_t elem, uint32_t seed, uint32_t thresh, bool& k0, bool& k1) {
  const uint32_t h = drop_hash(elem >> 1, seed);
  k0 = (h & 0xFFFFu) >= thresh;
  k1 = (h >> 16) >= thresh;
}

// quad form (attention probabilities): keys 4k .. 4k+3 of a query row share ONE hash evaluation.  Two avalanche rounds,
// then a 32 x 32 -> 64 bit multiply whose two result words supply four 16-bit uniforms: the low word folded with the
// high word (its own low bits are weak), the low half of the high word, and the mixed input's high half folded with the
// high word (the high half of the high word itself is bounded by the multiplier, i.e. not uniform).  14 instructions per
// four decisions instead of 2 x 12 with the pair form — the dropout decisions were 8.75 of the ~19 instructions per
// score of the attention forward (SASS r02) and the +8 / +12 us that dropout adds to the forward / backward kernels.
// Checked offline on 2^24 consecutive quads per seed (oracle-free numpy restatement, tools/dropout_quality.py): keep
// rate 0.90001, all four fields pass a 256-bin chi-square, pairwise joint drop rates 0.0100 +- 0.0001, lag-1 .. lag-200
// autocorrelations of the mask and of each field across neighbouring quads within 2 sigma of zero.
__device__ __forceinline__ void drop_keep4(uint32_t quad, uint32_t seed, uint32_t thresh, bool (&k)[4]) {
  uint32_t x = quad ^ seed;
  x *= 0x9E3779B1u; x ^= x >> 16;
  x *= 0x85EBCA6Bu; x ^= x >> 13;
  const unsigned long long w = (unsigned long long)x * 0xC2B2AE35u;
  const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w ^ hi;
  k[0] = (lo & 0xFFFFu) >= thresh;
  k[1] = (lo >> 16) >= thresh;
  k[2] = (hi & 0xFFFFu) >= thresh;
  k[3] = (((x >> 16) ^ hi) & 0xFFFFu) >= thresh;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Potentially blocking probe.  The suspend-time hint lets the hardware park the thread until the phase completes (or
// the hint expires) instead of answering "not yet" after a few dozen cycles: ncu r02 (short attention backward) showed
// 47 % of all executed instructions in try_wait poll loops of waiting warps (one poll per ~50 cycles and warp, six
// instructions each), taking issue slots from the warps that had work.
#ifndef VS_MBAR_SUSPEND_NS
#define VS_MBAR_SUSPEND_NS 20000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if VS_MBAR_SUSPEND_NS > 0
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)VS_MBAR_SUSPEND_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it answers)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("vitseg: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// proxy / tcgen05 fences
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA loads (tile mode, mbarrier completion)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
// Whole warp must call. Writes the TMEM base address to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base+i), columns [c, c+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (cf. PTX ISA "tcgen05 matrix descriptors"; field layout as in CUTLASS
// cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout_type [61,64) (2 = SWIZZLE_128B).
//
// Both operand flavours below live in shared memory as rows of 128 bytes (64 bf16) written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B; eight rows form one 1024-byte swizzle atom.
//   K-major  : a row is one M/N index, the 128 bytes run along K.  SBO = 1024 (next 8 M/N rows); LBO unused.
//              Advancing K by 16 elements = +32 bytes on the start address.
//   MN-major : a row is one K index, the 128 bytes run along M/N (64 elements).  SBO = 1024 (next 8 K rows);
//              LBO = byte distance between consecutive 64-element M/N blocks.  Advancing K by 16 = +2048 bytes.
// ----------------------------------------------------------------------------------------------
// The low word of a SWIZZLE_128B / SBO = 1024 descriptor (start address and LBO); its high word is the constant
// kUmmaDescHiSw128.  A byte offset (multiple of 16, inside the 256 KB shared-memory window) is applied by ADDING
// offset >> 4 to the low word, so an issuing thread that precomputes one low word per operand spends one add per
// descriptor instead of the shift / mask / or chain of umma_desc_sw128 (ncu r02: the single MMA-issuing warp of the short
// attention backward executed ~590 instructions per step, 10 per descriptor pair, and was the kernel's critical path).
constexpr uint32_t kUmmaDescHiSw128 = 0x40004040u;   // SBO 1024 >> 4 | version 1 << 14 | SWIZZLE_128B << 29
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulate.
// c_format F32 [4,6)=1 | a_format BF16 [7,10)=1 | b_format BF16 [10,13)=1 | a_major [15] | b_major [16] |
// N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                             uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// byte offset of 16-byte chunk `chunk` (0..7) in row `row` of a 128B-swizzled tile whose base is 1024-aligned
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}
#endif  // __CUDACC__

}  // namespace vs
