// knn_shortlist.cu -- stage 1 on the tensor cores: K1 (norms + FP16 copy), K2 (tcgen05 GEMM
// shortlist with the selection fused into the TMEM epilogue) and K3 (FP32 exact re-rank in
// the reference's operation order + top-k).
//
// Replaces the leaf loop of nanoflann's searchLevel
// (/root/reference/registration/nanoflann.hpp:1602-1622) for large query batches.  The
// Q x N distance matrix never touches HBM: the approximate score
//     s(q, x) = ||x||^2 - 2 <fp16(q), fp16(x)>        (D_apx = ||q||^2 + s)
// lives only in tensor memory; each epilogue thread owns one query row, keeps a running
// upper bound of the k-th smallest score and appends the 8-row groups that hold a score
// below it to that query's candidate list.  K3 selects among the listed scores, re-computes
// the survivors' distances exactly (bit-exact with L2_Adaptor::evalMetric,
// nanoflann.hpp:453-487) and picks the top-k by (d2, idx).
//
// Operand precision: FP16 (11-bit significand, u = 2^-11; same tensor rate as BF16, 8x
// tighter) after an exact power-of-two scaling (one scale for the database, one per query
// row) that keeps every element inside FP16's normal range.
//
// Exactness (proved in DESIGN.md "shortlist bound"): with dq = q - fp16(q), dx = x - fp16(x)
// (exact residuals, norms computed in K1),
//   |D_apx - D_ref| <= eps(q) = 2 (||dq|| Xmax + (1+u) ||q|| DXmax)     operand rounding (Cauchy-Schwarz)
//                              + 2^-11 ||q|| Xmax                        tensor-core FP32 accumulation
//                              + 2^-16 (||q|| + Xmax)^2 + tiny           FP32 norms / FFMA / reference rounding
// with Xmax = max ||x||, DXmax = max ||dx||.  Every true top-k row therefore has
// D_apx <= A_k + 2 eps, A_k = k-th smallest D_apx.  A thread's threshold is B + 2 eps with
// B >= A_k at all times (B = 32nd smallest score among 32 distinct rows already seen,
// k <= 32), hence the emitted groups contain every true top-k row.  Lists that overflow their
// capacity are detected and those queries are re-run through the exact scan on the GPU --
// never on a CPU.
//
// sm_100a only: tcgen05.mma (kind::f16, M=128, N=256, K=16, cta_group::1), accumulators in
// TMEM (2 x 256 columns, double buffered against the epilogue), operands staged by TMA
// (cp.async.bulk.tensor, 128B swizzle) and tracked with mbarriers.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "knn_kernels.cuh"
#include "knn_shortlist.cuh"

namespace gloc {

namespace {

// [emu-a-begin] (tests/cpp/gemm_emu_test.cpp compiles the marked regions for the host)
// ------------------------------------------------------------------ tile shape
constexpr int BM = 128;            // queries per tile = TMEM lanes = UMMA M
constexpr int BN = 256;            // DB rows per tile = UMMA N = TMEM columns per stage
constexpr int BK = 64;             // K elements per smem k-block (128 B of fp16: one swizzle row)
constexpr int UK = 16;             // UMMA K for 16-bit inputs
constexpr int kStagesB = 3;        // B ring depth
constexpr int kMaxKBlocks = 8;     // dim <= 512 keeps the whole query tile resident (128 KB)
constexpr int kABytesPerKB = BM * BK * 2;   // 16 KB
constexpr int kBBytes = BN * BK * 2;        // 32 KB
constexpr int kThreads = 384;      // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-11 epilogue
constexpr int kTmemCols = 512;
constexpr int kWarmChunks = 8;     // chunks whose scores all enter the threshold list (per warpgroup)
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kMaxKBlocks * kABytesPerKB +
                              (size_t)kStagesB * kBBytes + 256 /*barriers*/;
// CTA-pair variant (experimental, GLOC_KNN_PAIR=1): two CTAs of a cluster run one
// tcgen05.mma.cta_group::2 of M = 256 (each CTA its own 128-query tile) x N = 256; each CTA
// stages only its half of the database tile (128 rows, 16 KB per k-block), so the L2 -> SM
// traffic of the database halves and the same shared memory holds a ring twice as deep.
constexpr int kStagesBPair = 6;
constexpr int kBBytesPair = (BN / 2) * BK * 2;   // 16 KB
static_assert(kStagesBPair * kBBytesPair == kStagesB * kBBytes, "both variants use the same shared memory");
static_assert(2 * kStagesBPair + 6 <= 32, "barrier block is 256 bytes");

// shortlist error-bound constants (see header comment)
constexpr float kU = 1.f / 2048.f;          // FP16 unit roundoff
constexpr float kCAcc = 1.f / 2048.f;       // tensor-core accumulation, relative to ||q|| ||x||
constexpr float kC2 = 1.f / 65536.f;        // FP32 side computations, relative to (||q||+||x||)^2
constexpr float kInfl = 1.f + 1.f / 1024.f; // covers the rounding of the norms themselves

// monotone float <-> uint map so that unsigned atomicMin orders like the float
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// [emu-a-end]
// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (error surfaces on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// ---- cluster / CTA-pair flavours (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data into this CTA's shared memory, bytes counted on the barrier at
// `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map,
                                                 uint32_t bar_cluster_addr, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, FP16 inputs, FP32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// [emu-b-begin]
// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row atoms 1024 B apart (SBO).
// cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout_type SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 [4,6)=1, a/b_format F16 [7,10),
// [10,13)=0, a/b K-major [15],[16]=0, N>>3 [17,23), M>>4 [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kInstrDescPair = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
// [emu-b-end]

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// [emu-k1-begin]
// ------------------------------------------------------------------ K1: norms + FP16 copy
// power of two that maps max|v| into [2^13, 2^14): far from FP16 overflow (65504) and
// underflow (2^-14) for every element that matters
__host__ __device__ inline float pow2_scale_for(float max_abs) {
  if (!(max_abs > 0.f) || !(max_abs < 3.0e38f)) return 1.f;
  int e;
#ifdef __CUDA_ARCH__
  frexpf(max_abs, &e);
  return scalbnf(1.f, 14 - e);  // exact power of two
#else
  std::frexp(max_abs, &e);
  return std::ldexp(1.f, 14 - e);
#endif
}

// Database pass 1: ||x||^2 (FP32, tree order -- only the shortlist uses it), max ||x||^2 and
// max |x_i|.  One warp per row.
__global__ void knn_db_stats_kernel(const float* __restrict__ src, long long n, int dim,
                                    float* __restrict__ norms, unsigned* __restrict__ max_norm2_bits,
                                    unsigned* __restrict__ max_abs_bits) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)row * dim);
  float acc = 0.f, mx = 0.f;
  for (int i = lane; i < dim / 4; i += 32) {
    const float4 v = __ldg(s4 + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    norms[row] = acc;
    atomicMax(max_norm2_bits, __float_as_uint(acc));
    atomicMax(max_abs_bits, __float_as_uint(mx));
  }
}

// Database pass 2: FP16 copy of scale*x and the exact residual norm ||x - fp16(scale x)/scale||^2.
// Rows in [n, n_pad) are zero with norm = +inf so that padded tile columns are never listed.
__global__ void knn_db_convert_kernel(const float* __restrict__ src, long long n, long long n_pad,
                                      int dim, float scale, __half* __restrict__ dst,
                                      float* __restrict__ norms, unsigned* __restrict__ max_dx2_bits) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  __half2* d2 = reinterpret_cast<__half2*>(dst + (size_t)row * dim);
  if (row >= n) {
    for (int i = lane; i < dim / 2; i += 32) d2[i] = __floats2half2_rn(0.f, 0.f);
    if (lane == 0) norms[row] = __int_as_float(0x7f800000);
    return;
  }
  const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)row * dim);
  const float inv = 1.f / scale;
  float res = 0.f;
  for (int i = lane; i < dim / 4; i += 32) {
    const float4 v = __ldg(s4 + i);
    const __half2 h0 = __floats2half2_rn(v.x * scale, v.y * scale);
    const __half2 h1 = __floats2half2_rn(v.z * scale, v.w * scale);
    d2[2 * i] = h0;
    d2[2 * i + 1] = h1;
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const float r0 = v.x - f0.x * inv, r1 = v.y - f0.y * inv, r2 = v.z - f1.x * inv, r3 = v.w - f1.y * inv;
    res += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
  if (lane == 0) atomicMax(max_dx2_bits, __float_as_uint(res));
}

// Queries: one warp per row, everything in one pass with a per-row power-of-two scale (no
// host round trip): FP16 copy, ||q||^2, ||dq||^2 and 1/scale.
__global__ void knn_query_prep_kernel(const float* __restrict__ src, int n, int dim,
                                      __half* __restrict__ dst, float* __restrict__ qn,
                                      float* __restrict__ qe, float* __restrict__ qinv) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)row * dim);
  __half2* d2 = reinterpret_cast<__half2*>(dst + (size_t)row * dim);
  float4 v[kMaxKBlocks * BK / 128];  // dim <= 512: at most 4 float4 per lane
  float acc = 0.f, mx = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxKBlocks * BK / 128; ++j) {
    const int i = lane + 32 * j;
    v[j] = (i < dim / 4) ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const float scale = pow2_scale_for(mx), inv = 1.f / scale;
  float res = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxKBlocks * BK / 128; ++j) {
    const int i = lane + 32 * j;
    if (i < dim / 4) {
      const __half2 h0 = __floats2half2_rn(v[j].x * scale, v[j].y * scale);
      const __half2 h1 = __floats2half2_rn(v[j].z * scale, v[j].w * scale);
      d2[2 * i] = h0;
      d2[2 * i + 1] = h1;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const float r0 = v[j].x - f0.x * inv, r1 = v[j].y - f0.y * inv;
      const float r2 = v[j].z - f1.x * inv, r3 = v[j].w - f1.y * inv;
      res += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
  if (lane == 0) {
    qn[row] = acc;
    qe[row] = res;
    qinv[row] = inv;
  }
}

// [emu-k1-end]
// [emu-c-begin]
// ------------------------------------------------------------------ K2: GEMM shortlist
struct GemmArgs {
  int nq, n_qtiles, n_ranges, tiles_per_range, n_kb, k, cap, r_big;
  long long n_rows;                 // searchable rows
  const float* xn;                  // [n_pad] row norms (+inf beyond n_rows... see prep)
  const float* qn;                  // [nq] ||q||^2
  const float* qe;                  // [nq] ||q - fp16(q)||^2
  const float* qinv;                // [nq] 1 / (per-row power-of-two scale)
  float inv_sx;                     // 1 / (database power-of-two scale)
  const unsigned* max_norm2_bits;   // Xmax^2
  const unsigned* max_dx2_bits;     // DXmax^2
  unsigned* thr_ord;                // [nq] shared running threshold (ordered-uint of s-space)
  float* eps2;                      // [nq] 2*eps, written by the epilogue (read by K3)
  // candidate lists, [nq][2*n_ranges] lists (two epilogue warpgroups per unit) of cap groups
  unsigned* cand_g;                 // [list][cap]     base row of an 8-row group
  float4* cand_v;                   // [list][cap][2]  the group's 8 scores
  unsigned* unit_cnt;               // [nq][2*n_ranges]
};

// kPair = false: one CTA per SM, cta_group::1 (the shipped path).  kPair = true: clusters of two
// CTAs, cta_group::2 (see kStagesBPair); a.n_qtiles then counts PAIRS of query tiles and map_db
// has a 128-row box.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
knn_shortlist_gemm_kernel(const __grid_constant__ CUtensorMap map_q,
                          const __grid_constant__ CUtensorMap map_db, GemmArgs a) {
  constexpr int kStages = kPair ? kStagesBPair : kStagesB;
  constexpr int kStageBytes = kPair ? kBBytesPair : kBBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                          // n_kb x 16 KB (resident)
  unsigned char* sB = smem + (size_t)kMaxKBlocks * kABytesPerKB;     // kStages x kStageBytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kStages * kStageBytes);
  uint64_t* full_b = bars;                   // [kStages]
  uint64_t* empty_b = bars + kStages;        // [kStages]
  uint64_t* a_full = bars + 2 * kStages;     // [1]
  uint64_t* a_empty = a_full + 1;            // [1]
  uint64_t* tm_full = a_empty + 1;           // [2]
  uint64_t* tm_empty = tm_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = a.n_qtiles * a.n_ranges;
  // persistent worker = CTA (or CTA pair); in a pair only rank 0 ("leader") issues MMAs and owns
  // the barriers the operands and the accumulator hand-back are counted on
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
#define GLOC_WORKER (kPair ? (blockIdx.x >> 1) : blockIdx.x)
#define GLOC_N_WORKERS (kPair ? (gridDim.x >> 1) : gridDim.x)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_b + s, 1);
      mbar_init(empty_b + s, 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(tm_full + s, 1);
      mbar_init(tm_empty + s, kPair ? 16 : 8);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    if constexpr (kPair) {   // the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers are initialised, too
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tiles a unit runs -- every role computes the same schedule.
  auto unit_tiles = [&](int u, int& qt, int& rg, int& t_begin, int& t_count) {
    qt = u % a.n_qtiles;
    if constexpr (kPair) qt = 2 * qt + (int)cta_rank;   // this CTA's query tile of the pair
    rg = u / a.n_qtiles;
    t_begin = rg * a.tiles_per_range;
    const long long total_tiles = (a.n_rows + BN - 1) / BN;
    t_count = (int)min((long long)a.tiles_per_range, total_tiles - t_begin);
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, uphase = 0;
      // pair: both CTAs load their own halves; every byte is counted on the leader's barriers,
      // which the leader arms for both (the empty barriers are per CTA: multicast commits)
      const uint32_t a_full_ld = kPair ? map_to_cta(smem_u32(a_full), 0) : 0u;
      for (int u = GLOC_WORKER; u < n_units; u += GLOC_N_WORKERS) {
        int qt, rg, t_begin, t_count;
        unit_tiles(u, qt, rg, t_begin, t_count);
        mbar_wait(a_empty, uphase ^ 1);  // previous unit's MMAs no longer read the query tile
        if (leader) mbar_expect_tx(a_full, (uint32_t)(a.n_kb * kABytesPerKB) * (kPair ? 2u : 1u));
        for (int kb = 0; kb < a.n_kb; ++kb) {
          if constexpr (kPair) tma_load_2d_pair(sA + (size_t)kb * kABytesPerKB, &map_q, a_full_ld, kb * BK, qt * BM);
          else tma_load_2d(sA + (size_t)kb * kABytesPerKB, &map_q, a_full, kb * BK, qt * BM);
        }
        uphase ^= 1;
        for (int it = 0; it < t_count; ++it) {
          const int t = t_begin + it;
          for (int kb = 0; kb < a.n_kb; ++kb) {
            mbar_wait(empty_b + stage, phase ^ 1);
            if (leader) mbar_expect_tx(full_b + stage, kBBytes);   // 32 KB = both halves of a pair
            if constexpr (kPair)
              tma_load_2d_pair(sB + (size_t)stage * kStageBytes, &map_db,
                               map_to_cta(smem_u32(full_b + stage), 0), kb * BK,
                               t * BN + (int)cta_rank * (BN / 2));
            else
              tma_load_2d(sB + (size_t)stage * kStageBytes, &map_db, full_b + stage, kb * BK, t * BN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      mbar_wait(a_empty, uphase ^ 1);  // the last unit's commit has landed before the CTA exits
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0 && leader) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0, uphase = 0;
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA));
      const uint64_t b_desc0 = make_sw128_desc(smem_u32(sB));
      for (int u = GLOC_WORKER; u < n_units; u += GLOC_N_WORKERS) {
        int qt, rg, t_begin, t_count;
        unit_tiles(u, qt, rg, t_begin, t_count);
        mbar_wait(a_full, uphase);
        uphase ^= 1;
        for (int it = 0; it < t_count; ++it) {
          mbar_wait(tm_empty + as, aphase ^ 1);  // epilogue drained this accumulator stage
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < a.n_kb; ++kb) {
            mbar_wait(full_b + stage, phase);
            tcgen05_fence_after();
            // descriptors differ only in the 14-bit start-address field (units of 16 B)
            const uint64_t ad = a_desc0 + (uint64_t)(kb * (kABytesPerKB >> 4));
            const uint64_t bd = b_desc0 + (uint64_t)(stage * (kStageBytes >> 4));
#pragma unroll
            for (int k4 = 0; k4 < BK / UK; ++k4) {
              if constexpr (kPair)
                umma_f16_pair(d_tmem, ad + (uint64_t)(k4 * (UK * 2 >> 4)), bd + (uint64_t)(k4 * (UK * 2 >> 4)),
                              kInstrDescPair, (kb | k4) != 0 ? 1u : 0u);
              else
                umma_f16(d_tmem, ad + (uint64_t)(k4 * (UK * 2 >> 4)), bd + (uint64_t)(k4 * (UK * 2 >> 4)),
                         kInstrDesc, (kb | k4) != 0 ? 1u : 0u);
            }
            // B slot reusable once these MMAs retire (pair: in both CTAs)
            if constexpr (kPair) tcgen05_commit_pair(empty_b + stage); else tcgen05_commit(empty_b + stage);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          // accumulator ready for the epilogue (pair: each CTA's own 128 lanes)
          if constexpr (kPair) tcgen05_commit_pair(tm_full + as); else tcgen05_commit(tm_full + as);
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
        // query tile may be overwritten
        if constexpr (kPair) tcgen05_commit_pair(a_empty); else tcgen05_commit(a_empty);
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: selection out of TMEM
    // Thread = one query row.  Per 32-column chunk: scores s = ||x||^2 - 2 dot (32 FFMA) and
    // the minima of its four 8-column groups (FMNMX3 tree); a group whose minimum is under the
    // thread's threshold is appended, whole, to the candidate list (predicated stores, no
    // divergence).  Units that start without any bound also keep a sorted list of the 32
    // smallest scores seen, whose largest entry B = lst[31] bounds the 32nd (hence k-th,
    // k <= 32) smallest score of the whole database from above; the threshold is B + 2 eps,
    // shared between the units of a query through global memory.
    // Two warpgroups (warps 4-7 and 8-11) share every tile: group wg takes the chunks with
    // c % 2 == wg, so each scheduler has two epilogue warps to hide TMEM/L1 latency.  Both
    // threads of a query row keep their own list / candidate list (list index 2*range + wg)
    // and meet in the shared threshold.
    const int ew = warp & 3;                  // this warp's TMEM lane quadrant
    const int wg = (warp - 4) >> 2;           // 0 or 1
    const int row = ew * 32 + lane;           // query row inside the tile = TMEM lane
    int as = 0;
    uint32_t aphase = 0;
    const float xmax = sqrtf(__uint_as_float(*a.max_norm2_bits)) * kInfl;
    const float dxmax = sqrtf(__uint_as_float(*a.max_dx2_bits)) * kInfl;
    for (int u = GLOC_WORKER; u < n_units; u += GLOC_N_WORKERS) {
      int qt, rg, t_begin, t_count;
      unit_tiles(u, qt, rg, t_begin, t_count);
      const int q = qt * BM + row;
      const bool q_ok = q < a.nq;
      float eps2 = 0.f, cm = 0.f;
      if (q_ok) {
        const float qnorm = sqrtf(a.qn[q]) * kInfl, dq = sqrtf(a.qe[q]) * kInfl;
        const float e = 2.f * (dq * xmax + (1.f + kU) * qnorm * dxmax) + kCAcc * qnorm * xmax +
                        kC2 * (qnorm + xmax) * (qnorm + xmax) + 1e-30f;
        eps2 = 2.f * e;
        if (rg == 0 && wg == 0) a.eps2[q] = eps2;
        cm = -2.f * a.inv_sx * a.qinv[q];   // undoes both power-of-two scales (exact)
      }
      float thr = q_ok ? ord2f(a.thr_ord[q]) : -INFINITY;   // shared across this query's units
      float lst[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) lst[j] = INFINITY;
      unsigned cnt = 0;
      // scores emitted since the last flush and not yet in lst (FIFO; extra ones are only
      // emitted -- lst then holds a subset of the rows seen, which is still a valid bound)
      float pend[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pend[i] = INFINITY;
      int npend = 0;
      // No bound yet anywhere in the warp (first units of a query): the first kWarmChunks
      // chunks insert every score into lst with one warp-uniform instruction stream.
      int warm_left = __any_sync(0xffffffffu, q_ok && thr == INFINITY) ? kWarmChunks : 0;
      // Only units that start without any bound build a list of their own (warp-uniform): a
      // unit that starts under another unit's bound sees too few scores below it to ever fill
      // 32 entries, so maintaining one would be pure overhead.  It still emits every score
      // under the shared bound, which is all exactness needs.
      const bool own_list = warm_left > 0;
      const int n_lists = 2 * a.n_ranges;
      const size_t list_base = ((size_t)(q_ok ? q : 0) * n_lists + 2 * rg + wg) * (size_t)a.cap;
      unsigned* const cand_g = a.cand_g + list_base;
      float4* const cand_v = a.cand_v + 2 * list_base;

      // sorted insertion of one score; the largest of the 33 values drops out (+inf: no-op)
      auto lst_insert = [&](float w) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float lo = fminf(lst[i], w);
          w = fmaxf(lst[i], w);
          lst[i] = lo;
        }
      };
      // warp-uniform: every lane inserts its pending scores (lanes with fewer insert +inf)
      auto flush_pending = [&]() {
        const int mp = __reduce_max_sync(0xffffffffu, npend);
#pragma unroll 1
        for (int p = 0; p < mp; ++p) {
          const float w = pend[0];
#pragma unroll
          for (int i = 0; i < 7; ++i) pend[i] = pend[i + 1];
          pend[7] = INFINITY;
          lst_insert(w);
        }
        npend = 0;
        thr = fminf(thr, lst[31] + eps2);
      };

      for (int it = 0; it < t_count; ++it) {
        const int t = t_begin + it;
        // bound published by the other units of this query: read now, used after the tile
        unsigned seen = 0xFFFFFFFFu;
        if (q_ok) seen = *reinterpret_cast<volatile unsigned*>(a.thr_ord + q);
        mbar_wait(tm_full + as, aphase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
        const float4* xn4 = reinterpret_cast<const float4*>(a.xn + (size_t)t * BN);
        // rows >= n_rows (search limit inside this tile) must neither be short-listed nor
        // tighten the bound; rows >= n_total already carry ||x||^2 = +inf
        const int valid_cols = (int)min((long long)BN, a.n_rows - (long long)t * BN);
#pragma unroll 1
        for (int c = wg; c < BN / 32; c += 2) {
          float4 xr[8];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) xr[j4] = __ldg(xn4 + c * 8 + j4);
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          float sc[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            sc[4 * j4 + 0] = fmaf(cm, __uint_as_float(v[4 * j4 + 0]), xr[j4].x);
            sc[4 * j4 + 1] = fmaf(cm, __uint_as_float(v[4 * j4 + 1]), xr[j4].y);
            sc[4 * j4 + 2] = fmaf(cm, __uint_as_float(v[4 * j4 + 2]), xr[j4].z);
            sc[4 * j4 + 3] = fmaf(cm, __uint_as_float(v[4 * j4 + 3]), xr[j4].w);
          }
          if (valid_cols < BN) {  // warp-uniform, last tile of a limited search only
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j >= valid_cols) sc[j] = INFINITY;
          }
          const bool warm = warm_left > 0;   // warp-uniform
          if (warm) {
            --warm_left;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float r0 = sc[8 * g], r1 = sc[8 * g + 1], r2 = sc[8 * g + 2], r3 = sc[8 * g + 3];
              float r4 = sc[8 * g + 4], r5 = sc[8 * g + 5], r6 = sc[8 * g + 6], r7 = sc[8 * g + 7];
#pragma unroll 1
              for (int jj = 0; jj < 8; ++jj) {   // rolled: one copy of the insertion per group
                lst_insert(r0);
                r0 = r1; r1 = r2; r2 = r3; r3 = r4; r4 = r5; r5 = r6; r6 = r7;
              }
            }
            thr = fminf(thr, lst[31] + eps2);
          }
          float mg[4];   // minima of the four 8-column groups
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            mg[g] = fminf(fminf(fminf(sc[8 * g], sc[8 * g + 1]), fminf(sc[8 * g + 2], sc[8 * g + 3])),
                          fminf(fminf(sc[8 * g + 4], sc[8 * g + 5]), fminf(sc[8 * g + 6], sc[8 * g + 7])));
          }
          // Emission, group-granular and fully predicated (no divergence, no per-score work):
          // a lane whose 8-column group holds a score under its threshold stores the group's
          // base row and all 8 scores (one aligned 32-byte sector); K3 filters the individual
          // scores.  With ~1e3 hits per query some lane of the warp hits in nearly every
          // group, so any per-score or lane-divergent handling here would run all the time.
          const unsigned col0 = (unsigned)(t * BN + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (mg[g] <= thr) {   // thr = -inf on rows beyond nq
              if (cnt < (unsigned)a.cap) {
                cand_g[cnt] = col0 + 8 * g;
                cand_v[2 * cnt] = make_float4(sc[8 * g], sc[8 * g + 1], sc[8 * g + 2], sc[8 * g + 3]);
                cand_v[2 * cnt + 1] = make_float4(sc[8 * g + 4], sc[8 * g + 5], sc[8 * g + 6], sc[8 * g + 7]);
              }
              ++cnt;
            }
          }
          if (own_list && !warm) {
            // units that maintain their own bound: scores entering the sorted list go through
            // the pending FIFO (visited only where some lane has one: warp-uniform branch)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (__any_sync(0xffffffffu, mg[g] < lst[31])) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float vj = sc[8 * g + jj];
                  if (vj < lst[31] && npend < 8) {
#pragma unroll
                    for (int i = 7; i > 0; --i) pend[i] = pend[i - 1];
                    pend[0] = vj;
                    ++npend;
                  }
                }
              }
            }
          }
          // flush as soon as some lane's pending buffer is full (warp-uniform, no divergence)
          if (own_list && __any_sync(0xffffffffu, npend >= 8)) flush_pending();
        }
        if (own_list && __any_sync(0xffffffffu, npend > 0)) flush_pending();
        // all of this warp's TMEM reads of the stage are complete: hand it back to the MMA
        tcgen05_fence_before();
        __syncwarp();
        if constexpr (kPair) {   // the leader's MMA thread waits for the epilogues of both CTAs
          if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(tm_empty + as), 0));
        } else {
          if (lane == 0) mbar_arrive(tm_empty + as);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        if (q_ok) {  // publish / pick up the bound shared by all units of this query
          const unsigned mine = f2ord(thr);
          if (mine < seen) atomicMin(a.thr_ord + q, mine);
          else thr = ord2f(seen);
        }
      }
      if (q_ok) a.unit_cnt[(size_t)q * n_lists + 2 * rg + wg] = cnt;
    }
  }

  tcgen05_fence_before();
  if constexpr (kPair) cluster_sync_all();   // neither CTA leaves while its peer can still reach it
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(kTmemCols)
                   : "memory");
  }
}
#undef GLOC_WORKER
#undef GLOC_N_WORKERS
// [emu-c-end]

// [emu-k3-begin]
// ------------------------------------------------------------------ K3: select + exact re-rank
constexpr int kEntMax = 4096;    // group entries per query handled by K3 (more: exact-scan fallback)
constexpr int kCandMax = 2048;   // scores under the GEMM's final bound, per query
constexpr int kFinalMax = 256;   // candidates re-ranked exactly, per query
constexpr int kRerankThreads = 128;
constexpr int kMaxLists = 64;    // 2 * n_ranges

struct RerankArgs {
  const float* db;
  const float* q;
  int nq, dim, k, n_ranges, cap;   // n_ranges here = lists per query (2 per GEMM range)
  const unsigned* cand_g;
  const float* cand_v;
  const unsigned* unit_cnt;
  const unsigned* thr_ord;
  const float* eps2;
  uint64_t offset;
  uint64_t* out_idx;
  float* out_d2;
  int* overflow_list;     // compacted ids of queries that must be re-run exactly
  int* overflow_count;
  unsigned long long* rows_reranked;   // [0] rows re-ranked, [1] overflowed queries (cumulative)
};

// One CTA per query.  Select the rows to re-rank, re-rank them exactly, write the top-k.
// The selection is a radix select (k-th smallest approximate score) over the scores of the
// emitted groups, which stay in registers; only the exact re-rank touches the database.
__global__ void __launch_bounds__(kRerankThreads)
knn_shortlist_rerank_kernel(RerankArgs a) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int groups = a.dim / 4, gstride = groups + 1;
  float* G = reinterpret_cast<float*>(sm_raw);                            // [32][dim/4 + 1]
  unsigned* ckey = reinterpret_cast<unsigned*>(sm_raw);                   // [kCandMax] candidate keys  } dead before
  unsigned* crow = ckey + kCandMax;                                       // [kCandMax] candidate rows  } G is written
  const size_t g_bytes = max((size_t)32 * gstride * 4, (size_t)kCandMax * 8);
  unsigned* hist = reinterpret_cast<unsigned*>(sm_raw + g_bytes);         // [4][256]
  unsigned* fin_i = hist + 4 * 256;                                       // [kFinalMax]
  float* fin_d = reinterpret_cast<float*>(fin_i + kFinalMax);             // [kFinalMax]
  float* qs = fin_d + kFinalMax;                                          // [dim]
  __shared__ int n_fin, n_cand, bad;
  __shared__ unsigned s_cnt[kMaxLists], s_off[kMaxLists + 1], s_red[3][kRerankThreads / 32], s_sel[2];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { n_fin = 0; n_cand = 0; bad = 0; }
  // stage 0: everything that only depends on q, issued together
  unsigned my_cnt = 0;
  if (tid < a.n_ranges) my_cnt = a.unit_cnt[(size_t)q * a.n_ranges + tid];
  const float tau = ord2f(a.thr_ord[q]);
  const float eps2 = a.eps2[q];
  for (int i = tid; i < groups; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(a.q + (size_t)q * a.dim) + i);
  for (int i = tid; i < 4 * 256; i += kRerankThreads) hist[i] = 0u;
  __syncthreads();
  if (tid < a.n_ranges) {
    if (my_cnt > (unsigned)a.cap) bad = 1;
    s_cnt[tid] = min(my_cnt, (unsigned)a.cap);
  }
  __syncthreads();
  if (warp == 0) {   // exclusive prefix over <= 64 lists
    const unsigned c0 = lane < a.n_ranges ? s_cnt[lane] : 0u;
    const unsigned c1 = lane + 32 < a.n_ranges ? s_cnt[lane + 32] : 0u;
    unsigned x0 = c0, x1 = c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
      if (lane >= o) { x0 += y0; x1 += y1; }
    }
    const unsigned t0 = __shfl_sync(0xffffffffu, x0, 31);
    s_off[lane] = x0 - c0;
    s_off[lane + 32] = t0 + x1 - c1;
    if (lane == 31) {
      s_off[kMaxLists] = t0 + x1;
      if (t0 + x1 > (unsigned)kEntMax) bad = 1;
    }
  }
  __syncthreads();
  if (!bad) {
    // stage 1: one thread per group entry (base row + 8 scores): the scores under the GEMM's
    // final bound (the only ones that can be in the top-k; typically a small fraction of what
    // was emitted under the looser running bounds) are compacted into shared memory
    const int total = (int)s_off[kMaxLists];
    const size_t qbase = (size_t)q * a.n_ranges * (size_t)a.cap;
    auto entry_at = [&](int e) -> size_t {
      int lo = 0, hi = a.n_ranges - 1;   // largest r with s_off[r] <= e
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= (unsigned)e) lo = mid; else hi = mid - 1;
      }
      return qbase + (size_t)lo * a.cap + ((unsigned)e - s_off[lo]);
    };
    // k-th smallest of the cnt keys in ckey (radix select on key - kmin, 8 bits per pass
    // starting at the highest bit in which the keys differ).  `hist` must be zero on entry.
    auto radix_kth = [&](int cnt, unsigned kmin, unsigned kmax) -> unsigned {
      const unsigned range = kmax - kmin;
      const int nbits = 32 - __clz(range | 1u);
      const int passes = (nbits + 7) >> 3;       // 1..4
      unsigned prefix = 0u, kk = (unsigned)a.k;  // rank (1-based) inside the current prefix class
      for (int ps = 0; ps < passes; ++ps) {
        const int shift = 8 * (passes - 1 - ps);
        unsigned* h = hist + 256 * ps;
        for (int i = tid; i < cnt; i += kRerankThreads) {
          const unsigned d = ckey[i] - kmin;
          if (ps == 0 || (d >> (shift + 8)) == prefix) atomicAdd(&h[(d >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {   // smallest bin whose cumulative count reaches kk
          unsigned c[8], sum = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) { c[i] = h[lane * 8 + i]; sum += c[i]; }
          unsigned incl = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
          }
          const unsigned excl = incl - sum;
          if (excl < kk && kk <= incl) {   // exactly one lane
            unsigned run = excl;
            int bsel = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (run < kk && kk <= run + c[i]) { bsel = i; s_sel[1] = kk - run; }
              run += c[i];
            }
            s_sel[0] = (unsigned)(lane * 8 + bsel);
          }
        }
        __syncthreads();
        prefix = (prefix << 8) | s_sel[0];
        kk = s_sel[1];
      }
      return kmin + prefix;
    };
    float tau_cur = tau;
    int cnt = 0;
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int attempt = 0;; ++attempt) {
      kmin = 0xFFFFFFFFu;
      kmax = 0u;
      for (int e0 = tid; e0 < total; e0 += 2 * kRerankThreads) {   // two entries in flight per thread
        unsigned base[2] = {0u, 0u};
        float4 v[2][2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u * kRerankThreads;
          ok[u] = e < total;
          if (ok[u]) {
            const size_t at = entry_at(e);
            base[u] = a.cand_g[at];
            v[u][0] = reinterpret_cast<const float4*>(a.cand_v)[2 * at];
            v[u][1] = reinterpret_cast<const float4*>(a.cand_v)[2 * at + 1];
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!ok[u]) continue;
          const float sc[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w,
                               v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (sc[j] <= tau_cur && sc[j] < INFINITY) {   // +inf = masked / padded row
              const int pos = atomicAdd(&n_cand, 1);
              if (pos < kCandMax) {
                const unsigned k32 = f2ord(sc[j]);
                ckey[pos] = k32;
                crow[pos] = base[u] + j;
                kmin = min(kmin, k32);
                kmax = max(kmax, k32);
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
      }
      if (lane == 0) { s_red[1][warp] = kmin; s_red[2][warp] = kmax; }
      __syncthreads();
      cnt = n_cand;
      kmin = 0xFFFFFFFFu; kmax = 0u;
#pragma unroll
      for (int w2 = 0; w2 < kRerankThreads / 32; ++w2) {
        kmin = min(kmin, s_red[1][w2]);
        kmax = max(kmax, s_red[2][w2]);
      }
      if (cnt <= kCandMax || attempt == 1) break;
      // More candidates than fit (near-duplicate rows pass the bound together): the k-th
      // smallest of the kCandMax stored ones -- k distinct rows -- bounds A_k from above, so
      // collect again under that tighter bound.
      tau_cur = fminf(tau_cur, ord2f(radix_kth(kCandMax, kmin, kmax)) + eps2);
      __syncthreads();
      for (int i = tid; i < 4 * 256; i += kRerankThreads) hist[i] = 0u;
      if (tid == 0) n_cand = 0;
      __syncthreads();
    }
    if (cnt > kCandMax) {
      if (tid == 0) bad = 1;
    } else {
      // stage 2: A_k = k-th smallest candidate score, then keep the candidates with
      // s <= A_k + 2 eps.
      float tau2 = INFINITY;
      if (cnt >= a.k) tau2 = ord2f(radix_kth(cnt, kmin, kmax)) + eps2;
      for (int i = tid; i < cnt; i += kRerankThreads) {
        if (ord2f(ckey[i]) <= tau2) {
          const int pos = atomicAdd(&n_fin, 1);
          if (pos < kFinalMax) fin_i[pos] = crow[i];
        }
      }
    }
    __syncthreads();
    if (n_fin > kFinalMax && tid == 0) bad = 1;
    __syncthreads();
  }
  if (bad) {
    if (tid == 0) {
      const int pos = atomicAdd(a.overflow_count, 1);
      a.overflow_list[pos] = q;
      if (a.rows_reranked) atomicAdd(a.rows_reranked + 1, 1ull);
    }
    return;
  }
  const int nf = n_fin;
  // stage 3: exact distances, reference operation order (nanoflann.hpp:453-487): the
  // per-group sums ((d0^2+d1^2)+d2^2)+d3^2 are independent -- one warp per row, a lane takes
  // every 32nd group (coalesced 512 B reads), 4 rows x 4 groups in flight per lane; the
  // running sum over groups is the serial chain, one thread per row.
  for (int b0 = 0; b0 < nf; b0 += 32) {
    const int nb = min(32, nf - b0);
    for (int r0 = warp * 4; r0 < nb; r0 += (kRerankThreads / 32) * 4) {
      for (int g0 = 0; g0 < groups; g0 += 128) {
        float4 x[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4* row = reinterpret_cast<const float4*>(a.db + (size_t)fin_i[b0 + min(r0 + u, nb - 1)] * a.dim);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = g0 + lane + 32 * j;
            x[u][j] = g < groups ? __ldg(row + g) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r0 + u >= nb) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = g0 + lane + 32 * j;
            if (g >= groups) continue;
            const float4 qq = *reinterpret_cast<const float4*>(qs + 4 * g);
            const float d0 = __fsub_rn(qq.x, x[u][j].x), d1 = __fsub_rn(qq.y, x[u][j].y);
            const float d2 = __fsub_rn(qq.z, x[u][j].z), d3 = __fsub_rn(qq.w, x[u][j].w);
            float sg = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
            sg = __fadd_rn(sg, __fmul_rn(d2, d2));
            sg = __fadd_rn(sg, __fmul_rn(d3, d3));
            G[(r0 + u) * gstride + g] = sg;
          }
        }
      }
    }
    __syncthreads();
    if (tid < nb) {
      float r = 0.f;
      const float* gr = G + tid * gstride;
#pragma unroll 8
      for (int g = 0; g < groups; ++g) r = __fadd_rn(r, gr[g]);
      fin_d[b0 + tid] = r;
    }
    __syncthreads();
  }
  // stage 4: top-k by (d2, idx)
  uint64_t* oi = a.out_idx + (size_t)q * a.k;
  float* od = a.out_d2 + (size_t)q * a.k;
  for (int i = nf + tid; i < a.k; i += kRerankThreads) {   // fewer finalists than k: empty slots
    oi[i] = 0xFFFFFFFFFFFFFFFFull;
    od[i] = 3.402823466e+38f;
  }
  for (int i = tid; i < nf; i += kRerankThreads) {
    const uint64_t ki = pack_key(fin_d[i], fin_i[i]);
    int rank = 0;
    for (int j = 0; j < nf; ++j) rank += (pack_key(fin_d[j], fin_i[j]) < ki) ? 1 : 0;
    if (rank < a.k) {
      oi[rank] = (uint64_t)fin_i[i] + a.offset;
      od[rank] = fin_d[i];
    }
  }
  if (tid == 0 && a.rows_reranked) atomicAdd(a.rows_reranked, (unsigned long long)nf);
}

__global__ void knn_fill_u32_kernel(unsigned* p, size_t n, unsigned v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// [emu-k3-end]
// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// fp16 matrix [rows][dim] (K-major), box = BK x box_rows, 128B swizzle, OOB rows read as 0
bool make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t dim, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {dim, rows};
  cuuint64_t gstride[1] = {dim * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    const size_t want = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want; else p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

struct ShortlistState {
  // database-derived (valid for rows [0, prepared_rows))
  Buf db_h, xn, dbstats;   // dbstats: [0] max ||x||^2, [1] max |x_i|, [2] max ||dx||^2 (float bits)
  size_t prepared_rows = 0, prepared_pad = 0;
  const float* prepared_src = nullptr;
  float scale_x = 1.f;
  // per-call workspaces
  Buf q_h, qn, qe, qinv, thr, eps2, cand_g, cand_v, unit_cnt, ovf_list, ovf_count, rows_ctr, partial;
};

bool shortlist_supported(size_t dim, size_t k) {
  return dim % BK == 0 && dim >= BK && dim <= (size_t)kMaxKBlocks * BK && k >= 1 && k <= 32;
}

bool shortlist_applicable(size_t dim, size_t n_rows, size_t nq, size_t k) {
  // the GEMM wins once the batch fills tensor tiles; tiny batches stay on the exact scan
  return shortlist_supported(dim, k) && nq >= 64 && n_rows >= 1024;
}

void shortlist_invalidate(ShortlistState* s, size_t first_dirty_row) {
  if (s) s->prepared_rows = std::min(s->prepared_rows, first_dirty_row);
}

void shortlist_destroy(ShortlistState* s) {
  if (!s) return;
  for (Buf* b : {&s->db_h, &s->xn, &s->dbstats, &s->q_h, &s->qn, &s->qe, &s->qinv, &s->thr, &s->eps2,
                 &s->cand_g, &s->cand_v, &s->unit_cnt, &s->ovf_list, &s->ovf_count, &s->rows_ctr,
                 &s->partial})
    b->release();
  delete s;
}

namespace {

struct Plan {
  int n_ranges, tiles_per_range, r_big, cap;
};

// n_qtiles: query tiles (or pairs of them) a unit covers; sms: persistent workers
Plan make_plan_units(long long n_rows, int n_qtiles, int sms) {
  const long long tiles = (n_rows + BN - 1) / BN;
  const long long max_r = std::max<long long>(1, std::min<long long>(16, tiles / 8));
  double best = -1;
  long long best_r = 1;
  for (long long r = 1; r <= max_r; ++r) {
    const long long units = (long long)n_qtiles * r;
    const long long waves = (units + sms - 1) / sms;
    const double eff = (double)units / (double)(waves * sms);
    const double score = eff - 0.004 * (double)r;
    if (score > best + 1e-12) { best = score; best_r = r; }
  }
  Plan p;
  p.tiles_per_range = (int)((tiles + best_r - 1) / best_r);
  p.n_ranges = (int)((tiles + p.tiles_per_range - 1) / p.tiles_per_range);
  p.r_big = std::min<int>(p.n_ranges, (sms + n_qtiles - 1) / n_qtiles);
  p.cap = 512;
  return p;
}
Plan make_plan(long long n_rows, int nq, int sms) {
  return make_plan_units(n_rows, (nq + BM - 1) / BM, sms);
}

// GLOC_KNN_PAIR=1 (experimental): CTA-pair GEMM.  Returns the number of resident CTA pairs, 0
// when the variant is off or cannot be launched.
int pair_workers() {
  static int cached = -1;
  static int cached_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cached >= 0 && cached_dev == dev) return cached;
  cached_dev = dev;
  cached = 0;
  const char* e = getenv("GLOC_KNN_PAIR");
  if (!e || atoi(e) == 0) return 0;
  if (cudaFuncSetAttribute(knn_shortlist_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * sm_count(dev));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, knn_shortlist_gemm_kernel<true>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cached = n;
  return cached;
}

}  // namespace

int shortlist_pair_workers() { return pair_workers(); }

int shortlist_query(ShortlistState** sp, const ShortlistArgs& A, uint64_t* launches,
                    uint64_t* fallback, uint64_t* rows_reranked) {
  if (!*sp) *sp = new ShortlistState;
  ShortlistState* S = *sp;
  cudaStream_t st = A.stream;
  const int dim = (int)A.dim, k = (int)A.k;
  const int sms = sm_count(A.device);
  (void)fallback;
  (void)rows_reranked;

  // ---- K1 (cached): FP16 copy + norms of the database rows
  const size_t n_total = A.n_total;
  const size_t n_pad = (n_total + BN - 1) / BN * BN + BN;
  if (S->prepared_src != A.d_db || S->prepared_pad < n_pad) S->prepared_rows = 0;
  if (S->prepared_rows < n_total) {
    const bool fresh = S->prepared_rows == 0;
    if (fresh) {
      GLOC_CUDA_TRY(S->db_h.reserve(n_pad * dim * 2));
      GLOC_CUDA_TRY(S->xn.reserve(n_pad * sizeof(float)));
      GLOC_CUDA_TRY(S->dbstats.reserve(16));
      GLOC_CUDA_TRY(cudaMemsetAsync(S->dbstats.p, 0, 16, st));
      S->prepared_pad = std::min(S->db_h.bytes / ((size_t)dim * 2), S->xn.bytes / sizeof(float));
    }
    const size_t r0 = S->prepared_rows;
    unsigned* stats = (unsigned*)S->dbstats.p;
    const int wpb = 8;
    const long long new_rows = (long long)(n_total - r0);
    knn_db_stats_kernel<<<(unsigned)((new_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        A.d_db + r0 * dim, new_rows, dim, (float*)S->xn.p + r0, stats, stats + 1);
    GLOC_CUDA_TRY(cudaGetLastError());
    float max_abs = 0.f;  // one-time host round trip: the scale is a launch parameter
    GLOC_CUDA_TRY(cudaMemcpyAsync(&max_abs, stats + 1, 4, cudaMemcpyDeviceToHost, st));
    GLOC_CUDA_TRY(cudaStreamSynchronize(st));
    if (fresh) {
      S->scale_x = pow2_scale_for(max_abs);
    } else if (!(max_abs * S->scale_x < 32768.f)) {
      S->prepared_rows = 0;  // appended rows outgrew the scale: convert everything again
      return shortlist_query(sp, A, launches, fallback, rows_reranked);
    }
    const long long conv_rows = (long long)(n_pad - r0);
    knn_db_convert_kernel<<<(unsigned)((conv_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        A.d_db + r0 * dim, new_rows, conv_rows, dim, S->scale_x, (__half*)S->db_h.p + r0 * dim,
        (float*)S->xn.p + r0, stats + 2);
    GLOC_CUDA_TRY(cudaGetLastError());
    *launches += 2;
    S->prepared_rows = n_total;
    S->prepared_src = A.d_db;
  }

  const int pairs = pair_workers();   // 0: the shipped one-CTA-per-SM kernel
  CUtensorMap map_db;
  if (!make_map(&map_db, S->db_h.p, A.n_rows, (uint64_t)dim, pairs ? BN / 2 : BN))
    return fail(GLOC_ERR_CUDA, "cuTensorMapEncodeTiled(db) failed");

  // queries per pass: bounded so that the candidate workspace (2 lists per unit, 36 B per
  // group slot) stays under ~16 GB
  size_t kChunk = 65536;
  {
    const int nq0 = (int)std::min<size_t>(A.nq, kChunk);
    const Plan p0 = pairs ? make_plan_units((long long)A.n_rows, ((nq0 + BM - 1) / BM + 1) / 2, pairs)
                          : make_plan((long long)A.n_rows, nq0, sms);
    const size_t per_query = (size_t)p0.n_ranges * 2 * p0.cap * 36;
    kChunk = std::max<size_t>(BM, std::min<size_t>(kChunk, ((size_t)16 << 30) / per_query / BM * BM));
  }
  for (size_t q0 = 0; q0 < A.nq; q0 += kChunk) {
    const int nq = (int)std::min(kChunk, A.nq - q0);
    const int n_qtiles = (nq + BM - 1) / BM;
    const Plan plan = pairs ? make_plan_units((long long)A.n_rows, (n_qtiles + 1) / 2, pairs)
                            : make_plan((long long)A.n_rows, nq, sms);
    const size_t lists = (size_t)nq * plan.n_ranges * 2;  // two epilogue warpgroups per unit
    GLOC_CUDA_TRY(S->q_h.reserve((size_t)n_qtiles * BM * dim * 2));
    GLOC_CUDA_TRY(S->qn.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->qe.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->qinv.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->thr.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->eps2.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->cand_g.reserve(lists * plan.cap * 4));
    GLOC_CUDA_TRY(S->cand_v.reserve(lists * plan.cap * 32));
    GLOC_CUDA_TRY(S->unit_cnt.reserve(lists * 4));
    GLOC_CUDA_TRY(S->ovf_list.reserve((size_t)nq * 4));
    GLOC_CUDA_TRY(S->ovf_count.reserve(4));
    if (!S->rows_ctr.p) {
      GLOC_CUDA_TRY(S->rows_ctr.reserve(16));
      GLOC_CUDA_TRY(cudaMemsetAsync(S->rows_ctr.p, 0, 16, st));
    }
    const float* dq = A.d_q + q0 * dim;

    // K1 for the queries
    {
      const int wpb = 8;
      knn_query_prep_kernel<<<(nq + wpb - 1) / wpb, wpb * 32, 0, st>>>(
          dq, nq, dim, (__half*)S->q_h.p, (float*)S->qn.p, (float*)S->qe.p, (float*)S->qinv.p);
      GLOC_CUDA_TRY(cudaGetLastError());
      knn_fill_u32_kernel<<<(nq + 255) / 256, 256, 0, st>>>((unsigned*)S->thr.p, (size_t)nq,
                                                          0xFF800000u);  // f2ord(+inf)
      GLOC_CUDA_TRY(cudaGetLastError());
      GLOC_CUDA_TRY(cudaMemsetAsync(S->ovf_count.p, 0, 4, st));
      *launches += 2;
    }
    CUtensorMap map_q;
    if (!make_map(&map_q, S->q_h.p, (uint64_t)nq, (uint64_t)dim, BM))
      return fail(GLOC_ERR_CUDA, "cuTensorMapEncodeTiled(queries) failed");

    // K2
    GemmArgs g;
    g.nq = nq;
    g.n_qtiles = pairs ? (n_qtiles + 1) / 2 : n_qtiles;   // units per range
    g.n_ranges = plan.n_ranges;
    g.tiles_per_range = plan.tiles_per_range;
    g.n_kb = dim / BK;
    g.k = k;
    g.cap = plan.cap;
    g.r_big = plan.r_big;
    g.n_rows = (long long)A.n_rows;
    g.xn = (const float*)S->xn.p;
    g.qn = (const float*)S->qn.p;
    g.qe = (const float*)S->qe.p;
    g.qinv = (const float*)S->qinv.p;
    g.inv_sx = 1.f / S->scale_x;
    g.max_norm2_bits = (const unsigned*)S->dbstats.p;
    g.max_dx2_bits = (const unsigned*)S->dbstats.p + 2;
    g.thr_ord = (unsigned*)S->thr.p;
    g.eps2 = (float*)S->eps2.p;
    g.cand_g = (unsigned*)S->cand_g.p;
    g.cand_v = (float4*)S->cand_v.p;
    g.unit_cnt = (unsigned*)S->unit_cnt.p;
    static unsigned long long attr_mask = 0;
    if (first_use_on_current_device(attr_mask)) {
      GLOC_CUDA_TRY(cudaFuncSetAttribute(knn_shortlist_gemm_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes));
    }
    const int n_units = g.n_qtiles * plan.n_ranges;
    cudaError_t ge;
    if (pairs) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (unsigned)std::min(n_units, pairs));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = kSmemBytes;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (A.prof) A.prof->begin(st);
      ge = cudaLaunchKernelEx(&cfg, knn_shortlist_gemm_kernel<true>, map_q, map_db, g);
      if (A.prof) A.prof->end(st);
    } else {
      const int grid = std::min(n_units, sms);
      if (A.prof) A.prof->begin(st);
      knn_shortlist_gemm_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(map_q, map_db, g);
      ge = cudaGetLastError();
      if (A.prof) A.prof->end(st);
    }
    GLOC_CUDA_TRY(ge);

    // K3
    RerankArgs r;
    r.db = A.d_db;
    r.q = dq;
    r.nq = nq;
    r.dim = dim;
    r.k = k;
    r.n_ranges = plan.n_ranges * 2;   // lists per query
    r.cap = plan.cap;
    r.cand_g = g.cand_g;
    r.cand_v = (const float*)g.cand_v;
    r.unit_cnt = g.unit_cnt;
    r.thr_ord = g.thr_ord;
    r.eps2 = g.eps2;
    r.offset = A.offset;
    r.out_idx = A.d_idx + q0 * k;
    r.out_d2 = A.d_d2 + q0 * k;
    r.overflow_list = (int*)S->ovf_list.p;
    r.overflow_count = (int*)S->ovf_count.p;
    r.rows_reranked = (unsigned long long*)S->rows_ctr.p;
    const size_t rr_smem = std::max((size_t)kCandMax * 8, (size_t)32 * (dim / 4 + 1) * 4) + (size_t)4 * 256 * 4 +
                           (size_t)kFinalMax * 8 + (size_t)dim * 4;
    static unsigned long long attr2_mask = 0;
    if (first_use_on_current_device(attr2_mask)) {
      GLOC_CUDA_TRY(cudaFuncSetAttribute(knn_shortlist_rerank_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    knn_shortlist_rerank_kernel<<<nq, kRerankThreads, rr_smem, st>>>(r);
    GLOC_CUDA_TRY(cudaGetLastError());
    *launches += 2;

    if (getenv("GLOC_DEBUG_SHORTLIST")) {  // tuning aid: candidate-list statistics of this chunk
      cudaStreamSynchronize(st);
      std::vector<unsigned> hc(lists), ht(nq);
      std::vector<float> he(nq);
      int ovf = 0;
      cudaMemcpy(hc.data(), S->unit_cnt.p, lists * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ht.data(), S->thr.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(he.data(), S->eps2.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&ovf, S->ovf_count.p, 4, cudaMemcpyDeviceToHost);
      unsigned long long sum = 0, mx = 0, over = 0;
      for (unsigned c : hc) { sum += c; mx = std::max<unsigned long long>(mx, c); over += c > (unsigned)plan.cap; }
      unsigned long long sum_first = 0;
      for (int qi = 0; qi < nq; ++qi) sum_first += hc[(size_t)qi * plan.n_ranges * 2] + hc[(size_t)qi * plan.n_ranges * 2 + 1];
      fprintf(stderr, "[shortlist] nq=%d ranges=%d tiles/range=%d r_big=%d cap=%d | emitted/query=%.1f "
                      "(range0 %.1f) max/list=%llu lists_over_cap=%llu | overflowed queries=%d | thr[0]=%08x eps2[0]=%g\n",
              nq, plan.n_ranges, plan.tiles_per_range, plan.r_big, plan.cap, (double)sum / nq,
              (double)sum_first / nq, mx, over, ovf, ht[0], he[0]);
    }

    // overflowed queries: exact scan on the device, sized for the worst case, count read on
    // the device (surplus CTAs exit immediately; usually every CTA does)
    {
      // small-tile configuration (16 queries x 256 rows per CTA) over many row ranges: a
      // handful of overflowed queries still spreads over the whole GPU
      const int BNs = 256;
      const int n_r = (int)std::max<long long>(1, std::min<long long>(32, (long long)A.n_rows / (BNs * 4LL)));
      long long rpr = ((long long)A.n_rows + n_r - 1) / n_r;
      rpr = (rpr + BNs - 1) / BNs * BNs;
      const int n_ranges = (int)(((long long)A.n_rows + rpr - 1) / rpr);
      GLOC_CUDA_TRY(S->partial.reserve((size_t)nq * n_ranges * k * 8));
      GLOC_CUDA_TRY(launch_knn_exact_scan(A.d_db, (long long)A.n_rows, dim, dq, nq, k, n_ranges, rpr,
                                          (uint64_t*)S->partial.p, st, (const int*)S->ovf_list.p,
                                          (const int*)S->ovf_count.p, /*force_small=*/true));
      GLOC_CUDA_TRY(launch_knn_finalize((const uint64_t*)S->partial.p, nq, n_ranges, k, A.offset,
                                        A.d_idx + q0 * k, A.d_d2 + q0 * k, st,
                                        (const int*)S->ovf_list.p, (const int*)S->ovf_count.p));
      *launches += 2;
    }
  }
  return GLOC_OK;
}

// Cumulative device-side counters (rows re-ranked exactly, queries that overflowed and were
// re-run by the exact scan).  Synchronises the device.
int shortlist_counters(ShortlistState* s, uint64_t* rows, uint64_t* overflow) {
  *overflow = 0;
  *rows = 0;
  if (!s || !s->rows_ctr.p) return GLOC_OK;
  unsigned long long c[2] = {0, 0};
  GLOC_CUDA_TRY(cudaDeviceSynchronize());
  GLOC_CUDA_TRY(cudaMemcpy(c, s->rows_ctr.p, 16, cudaMemcpyDeviceToHost));
  *rows = c[0];
  *overflow = c[1];
  return GLOC_OK;
}

}  // namespace gloc
