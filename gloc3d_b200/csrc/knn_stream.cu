// knn_stream.cu -- stage 1 for one to four queries per call: the HBM-bound regime.
//
// The reference issues ONE query per call (RpyPCLoopDetector::detect,
// /root/reference/registration/loop_detector.cpp:42-45 and :73-79 in SLAM mode).  With so
// few queries every database byte is used for a handful of flops, so the search is a single
// streaming pass over the float32 rows at HBM speed -- no tensor cores, no FP16 copy.
//
// Exactness is the same contract as everywhere (L2_Adaptor::evalMetric,
// nanoflann.hpp:453-487): r = 0; r += ((d0^2 + d1^2) + d2^2) + d3^2 per group of 4 dims, in
// order, float32, no FMA.  The 128 group sums of a (row, query) pair are independent, the
// running sum over them is a serial chain.  A warp therefore works in two phases on a batch
// of 32/QN rows: (1) lanes = dims -- coalesced 512-byte row reads, every lane produces the
// sums of groups lane, lane+32, ... into a shared-memory tile; (2) lanes = (row, query)
// pairs -- each lane adds up its 128 group sums in the reference's order (bank-conflict
// free: tile stride 129).  The top-k lives in registers (lane i holds the i-th best key of
// each query), inserts are warp shuffles.
#include <algorithm>

#include "knn_kernels.cuh"

namespace gloc {

namespace {

constexpr int kStreamWarps = 4;      // warps per CTA
constexpr int kStreamMaxJ = 4;       // float4 groups per lane: dim <= 512
constexpr int kStreamCandMax = 2048; // keys gathered by the final merge
constexpr int kStreamHeadCache = 2048;   // list heads cached in shared memory by the final merge

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// Insert `key` (warp-uniform, smaller than the current k-th best) into the ascending list
// held one entry per lane.
__device__ __forceinline__ void list_insert(uint64_t& L, uint64_t key, int lane) {
  const unsigned larger = __ballot_sync(0xffffffffu, L > key);   // a suffix of the lanes
  const int pos = larger ? __ffs(larger) - 1 : 32;
  const uint64_t up = __shfl_up_sync(0xffffffffu, L, 1);
  if (lane > pos) L = up;
  if (lane == pos) L = key;
}

template <int QN>
__global__ void __launch_bounds__(kStreamWarps * 32)
knn_stream_kernel(const float* __restrict__ db, long long n_rows, int dim,
                  const float* __restrict__ q, int nq, int k, uint64_t* __restrict__ partial) {
  constexpr int RB = 32 / QN;   // rows per warp batch; lane = row_in_batch * QN + query
  extern __shared__ __align__(16) float stream_smem[];
  __shared__ uint64_t merge_buf[kStreamWarps][QN][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = dim >> 2, gs = groups + 1;
  float* S = stream_smem + (size_t)warp * 32 * gs;

  float4 qv[QN][kStreamMaxJ];
#pragma unroll
  for (int qi = 0; qi < QN; ++qi)
#pragma unroll
    for (int j = 0; j < kStreamMaxJ; ++j) {
      const int g = lane + 32 * j;
      qv[qi][j] = (qi < nq && g < groups)
                      ? __ldg(reinterpret_cast<const float4*>(q + (size_t)qi * dim) + g)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  uint64_t L[QN], worst[QN];
#pragma unroll
  for (int qi = 0; qi < QN; ++qi) L[qi] = worst[qi] = kEmptyKey;

  const long long n_batches = (n_rows + RB - 1) / RB;
  const long long w_total = (long long)gridDim.x * kStreamWarps;
  for (long long b = (long long)blockIdx.x * kStreamWarps + warp; b < n_batches; b += w_total) {
    const long long r0 = b * RB;
    // ---- phase 1: group sums, 4 rows in flight per lane
#pragma unroll 1
    for (int sub = 0; sub < RB; sub += 4) {
      float4 x[4][kStreamMaxJ];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long row = min(r0 + sub + u, n_rows - 1);
        const float4* rp = reinterpret_cast<const float4*>(db + (size_t)row * dim);
#pragma unroll
        for (int j = 0; j < kStreamMaxJ; ++j) {
          const int g = lane + 32 * j;
          x[u][j] = g < groups ? ldg_stream(rp + g) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int qi = 0; qi < QN; ++qi) {
          float* srow = S + (size_t)((sub + u) * QN + qi) * gs;
#pragma unroll
          for (int j = 0; j < kStreamMaxJ; ++j) {
            const int g = lane + 32 * j;
            if (g < groups) {
              const float d0 = __fsub_rn(qv[qi][j].x, x[u][j].x), d1 = __fsub_rn(qv[qi][j].y, x[u][j].y);
              const float d2 = __fsub_rn(qv[qi][j].z, x[u][j].z), d3 = __fsub_rn(qv[qi][j].w, x[u][j].w);
              float sg = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
              sg = __fadd_rn(sg, __fmul_rn(d2, d2));
              sg = __fadd_rn(sg, __fmul_rn(d3, d3));
              srow[g] = sg;
            }
          }
        }
    }
    __syncwarp();
    // ---- phase 2: the serial chain, one (row, query) pair per lane
    float r = 0.f;
    {
      const float* sr = S + (size_t)lane * gs;
#pragma unroll 16
      for (int g = 0; g < groups; ++g) r = __fadd_rn(r, sr[g]);
    }
    __syncwarp();
    const long long row = r0 + lane / QN;
    const int my_q = lane % QN;
    const uint64_t key = (row < n_rows && my_q < nq) ? pack_key(r, (uint32_t)row) : kEmptyKey;
    uint64_t my_worst = worst[0];
#pragma unroll
    for (int qi = 1; qi < QN; ++qi)
      if (my_q == qi) my_worst = worst[qi];
    unsigned m = __ballot_sync(0xffffffffu, key < my_worst);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
      const int c = src % QN;
#pragma unroll
      for (int qi = 0; qi < QN; ++qi) {
        if (c == qi && kk < worst[qi]) {   // warp-uniform
          list_insert(L[qi], kk, lane);
          worst[qi] = __shfl_sync(0xffffffffu, L[qi], k - 1);
        }
      }
    }
  }

  // ---- CTA merge: warp 0 folds the other warps' lists into its own
#pragma unroll
  for (int qi = 0; qi < QN; ++qi) merge_buf[warp][qi][lane] = L[qi];
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < kStreamWarps; ++w)
#pragma unroll
      for (int qi = 0; qi < QN; ++qi)
        for (int i = 0; i < k; ++i) {
          const uint64_t kk = merge_buf[w][qi][i];
          if (kk >= worst[qi]) break;   // ascending: nothing further can enter
          list_insert(L[qi], kk, lane);
          worst[qi] = __shfl_sync(0xffffffffu, L[qi], k - 1);
        }
#pragma unroll
    for (int qi = 0; qi < QN; ++qi)
      if (qi < nq && lane < k)
        partial[((size_t)qi * gridDim.x + blockIdx.x) * k + lane] = L[qi];
  }
}

// Global top-k of n_lists ascending lists of k keys per query.  The k smallest list heads
// are k distinct keys, so the k-th smallest head distance T bounds the answer's k-th
// distance from above: only keys with distance <= T (at most ~k per list with head <= T) are
// gathered and ranked.  One CTA per query.
__global__ void __launch_bounds__(256)
knn_stream_merge_kernel(const uint64_t* __restrict__ partial, int n_lists, int k,
                        uint64_t idx_offset, uint64_t* __restrict__ out_idx,
                        float* __restrict__ out_d2, int* __restrict__ overflow) {
  __shared__ uint64_t cand[kStreamCandMax];
  __shared__ int n_cand;
  __shared__ unsigned s_red[2][8], s_cnt[2][8], s_heads[8];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t* P = partial + (size_t)q * n_lists * k;
  if (tid == 0) n_cand = 0;
  // k-th smallest head distance (bit pattern order == value order for d2 >= 0): bisection
  // over the heads cached in shared memory (0xFFFFFFFF = empty list)
  __shared__ unsigned s_hd[kStreamHeadCache];
  const bool cached = n_lists <= kStreamHeadCache;
  unsigned lo = 0xFFFFFFFFu, hi = 0u;
  int heads = 0;
  for (int l = tid; l < n_lists; l += 256) {
    const uint64_t h = P[(size_t)l * k];
    const unsigned d = h != kEmptyKey ? (unsigned)(h >> 32) : 0xFFFFFFFFu;
    if (cached) s_hd[l] = d;
    if (h != kEmptyKey) {
      lo = min(lo, d);
      hi = max(hi, d);
      ++heads;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    heads += __shfl_xor_sync(0xffffffffu, heads, o);
  }
  if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; s_heads[warp] = (unsigned)heads; }
  __syncthreads();
  lo = 0xFFFFFFFFu; hi = 0u; heads = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { lo = min(lo, s_red[0][w]); hi = max(hi, s_red[1][w]); heads += (int)s_heads[w]; }
  unsigned T = 0xFFFFFFFFu;   // fewer than k non-empty lists: take everything
  if (heads >= k) {
    int round = 0;
    while (lo < hi) {
      const unsigned mid = lo + ((hi - lo) >> 1);
      int c = 0;
      if (cached) {
        for (int l = tid; l < n_lists; l += 256) c += s_hd[l] <= mid ? 1 : 0;   // mid < 0xFFFFFFFF
      } else {
        for (int l = tid; l < n_lists; l += 256) {
          const uint64_t h = P[(size_t)l * k];
          c += (h != kEmptyKey && (unsigned)(h >> 32) <= mid) ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) s_cnt[round & 1][warp] = (unsigned)c;
      __syncthreads();
      int total = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) total += (int)s_cnt[round & 1][w];
      if (total >= k) hi = mid; else lo = mid + 1;
      ++round;
    }
    T = lo;
  }
  __syncthreads();
  for (int e = tid; e < n_lists * k; e += 256) {
    const uint64_t key = P[e];
    if (key != kEmptyKey && (unsigned)(key >> 32) <= T) {
      const int pos = atomicAdd(&n_cand, 1);
      if (pos < kStreamCandMax) cand[pos] = key;
    }
  }
  __syncthreads();
  const int nc = n_cand;
  if (nc > kStreamCandMax) {   // pathological ties: the generic merge kernel re-does the call
    if (tid == 0) atomicExch(overflow, (int)gridDim.x);
    return;
  }
  uint64_t* oi = out_idx + (size_t)q * k;
  float* od = out_d2 + (size_t)q * k;
  for (int i = nc + tid; i < k; i += 256) {
    oi[i] = 0xFFFFFFFFFFFFFFFFull;
    od[i] = 3.402823466e+38f;
  }
  for (int i = tid; i < nc; i += 256) {
    const uint64_t ki = cand[i];
    int rank = 0;
    for (int j = 0; j < nc; ++j) rank += cand[j] < ki ? 1 : 0;
    if (rank < k) {
      oi[rank] = (uint64_t)(uint32_t)(ki & 0xFFFFFFFFull) + idx_offset;
      od[rank] = __uint_as_float((uint32_t)(ki >> 32));
    }
  }
}

}  // namespace

bool stream_applicable(size_t dim, size_t nq, size_t k) {
  return nq >= 1 && nq <= 4 && k <= 32 && dim % 4 == 0 && dim >= 4 && dim <= 128 * kStreamMaxJ;
}

int stream_grid(int device, size_t dim) {
  const size_t smem = (size_t)kStreamWarps * 32 * (dim / 4 + 1) * sizeof(float);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)200 * 1024) / std::max<size_t>(smem, 1)));
  return sm_count(device) * per_sm;
}

// partial: [nq][grid][k] keys.  *overflow (device int, zeroed by the caller) becomes nq when
// the final merge met more than kStreamCandMax tied keys; the generic list merge then runs
// (its grid exits immediately otherwise).
cudaError_t launch_knn_stream(const float* db, long long n_rows, int dim, const float* q, int nq,
                              int k, int grid, uint64_t* partial, uint64_t idx_offset,
                              uint64_t* out_idx, float* out_d2, int* overflow,
                              EventProfiler* prof, cudaStream_t stream) {
  const size_t smem = (size_t)kStreamWarps * 32 * (dim / 4 + 1) * sizeof(float);
  cudaError_t e = cudaSuccess;
#define GLOC_STREAM(QN)                                                                       \
  do {                                                                                        \
    static unsigned long long attr_mask = 0;                                                  \
    if (first_use_on_current_device(attr_mask)) {                                             \
      e = cudaFuncSetAttribute(knn_stream_kernel<QN>,                                         \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);      \
      if (e != cudaSuccess) return e;                                                         \
    }                                                                                         \
    if (prof) prof->begin(stream);                                                            \
    knn_stream_kernel<QN><<<grid, kStreamWarps * 32, smem, stream>>>(db, n_rows, dim, q, nq,  \
                                                                     k, partial);             \
    e = cudaGetLastError();                                                                   \
    if (prof) prof->end(stream);                                                              \
  } while (0)
  if (nq == 1) GLOC_STREAM(1);
  else if (nq == 2) GLOC_STREAM(2);
  else GLOC_STREAM(4);
#undef GLOC_STREAM
  if (e != cudaSuccess) return e;
  knn_stream_merge_kernel<<<nq, 256, 0, stream>>>(partial, grid, k, idx_offset, out_idx, out_d2,
                                                  overflow);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_knn_finalize(partial, nq, grid, k, idx_offset, out_idx, out_d2, stream, nullptr,
                             overflow);
}

}  // namespace gloc
