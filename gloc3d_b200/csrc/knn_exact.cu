// knn_exact.cu -- K3 as a full scan: exact float32 squared-L2 distances in the
// reference's operation order + fused per-query top-k, plus the list-merge
// kernels (range merge, K4 shard merge).
//
// Replaces the leaf loop of nanoflann's searchLevel
// (/root/reference/registration/nanoflann.hpp:1602-1622), i.e.
// L2_Adaptor::evalMetric (:453-487) + KNNResultSet::addPoint (:200-233), for
// every row of the database (the KD-tree prunes almost nothing at D=512).
//
// Arithmetic contract (bit-exact with the reference's x86-64 Release build):
//   r = 0; for each group of 4 dims: r = r + (((d0*d0 + d1*d1) + d2*d2) + d3*d3)
//   then the 0-3 tail dims one by one: r = r + d*d.   float32, round-to-nearest,
//   NO fused multiply-add -> only __fsub_rn/__fmul_rn/__fadd_rn are used (nvcc
//   never contracts those intrinsics).
//
// Roofline: this kernel is FP32-issue bound (12 non-fusable ops per 4 dims per
// (query,row) pair = 3 ops/dim), not HBM bound, for any query batch above ~16.
#include <algorithm>

#include "knn_kernels.cuh"

namespace gloc {

namespace {

constexpr int DK = 32;       // dims staged per shared-memory chunk
constexpr int DKP = DK + 4;  // padded row stride: conflict-free LDS.128 across rows

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage `rows` x DK floats (dims [c0, c0+DK) of rows [r0, r0+rows)) into smem,
// zero-filling rows >= r_end and dims >= dim.
template <int ROWS, int NT>
__device__ __forceinline__ void stage_chunk(float* dst, const float* __restrict__ src,
                                            long long r0, long long r_end, int dim, int c0,
                                            bool vec_ok, int tid,
                                            const int* __restrict__ rowmap = nullptr) {
  constexpr int SLOTS = ROWS * (DK / 4);
#pragma unroll
  for (int s = tid; s < SLOTS; s += NT) {
    const int row = s / (DK / 4);
    const int col = c0 + (s % (DK / 4)) * 4;
    float* d = dst + row * DKP + (s % (DK / 4)) * 4;
    const long long gr = r0 + row;
    if (gr < r_end && col < dim) {
      const float* g = src + (size_t)(rowmap ? (long long)rowmap[gr] : gr) * dim + col;
      if (vec_ok && col + 4 <= dim) {
        cp_async16(d, g);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        v.x = g[0];
        if (col + 1 < dim) v.y = g[1];
        if (col + 2 < dim) v.z = g[2];
        if (col + 3 < dim) v.w = g[3];
        *reinterpret_cast<float4*>(d) = v;
      }
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Merge `cnt` (<= 32) unsorted distinct keys (one per lane) into the ascending
// list L[0..k).  Whole warp participates.
template <int KCAP>
__device__ __forceinline__ void warp_merge_into_list(uint64_t* L, int k, uint64_t e, int cnt,
                                                     int lane) {
  constexpr int PER = KCAP / 32;
  int rank_e;
  {
    int lo = 0, hi = k;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (L[mid] < e) lo = mid + 1; else hi = mid;
    }
    rank_e = lo;
  }
  uint64_t le[PER];
  int np[PER];
#pragma unroll
  for (int s = 0; s < PER; ++s) {
    const int i = lane + 32 * s;
    le[s] = (i < k) ? L[i] : kEmptyKey;
    np[s] = i;
  }
  for (int t = 0; t < cnt; ++t) {
    const uint64_t other = __shfl_sync(0xffffffffu, e, t);
    rank_e += (other < e) ? 1 : 0;
#pragma unroll
    for (int s = 0; s < PER; ++s) np[s] += (other < le[s]) ? 1 : 0;
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < PER; ++s) {
    const int i = lane + 32 * s;
    if (i < k && np[s] < k) L[np[s]] = le[s];
  }
  if (lane < cnt && rank_e < k) L[rank_e] = e;
  __syncwarp();
}

template <int BQ, int BN, int TQ, int TN, int KCAP>
struct ScanCfg {
  static constexpr int TYN = BQ / TQ;
  static constexpr int TXN = BN / TN;
  static constexpr int NT = TYN * TXN;
  static constexpr int QCAP = TXN;  // one push per (query, tx) per phase at most
  static constexpr size_t kSmem = (size_t)2 * (BQ + BN) * DKP * sizeof(float) +
                                  (size_t)BQ * KCAP * 8 + (size_t)BQ * QCAP * 8 +
                                  (size_t)BQ * sizeof(int);
  static_assert(QCAP <= 32 || QCAP == 64, "queue merged one warp-load at a time");
};

template <int BQ, int BN, int TQ, int TN, int KCAP>
__global__ void __launch_bounds__(ScanCfg<BQ, BN, TQ, TN, KCAP>::NT, 1)
knn_exact_scan_kernel(const float* __restrict__ db, long long n_rows, int dim,
                      const float* __restrict__ q, int nq, int k, int n_qtiles,
                      long long rows_per_range, int n_ranges,
                      uint64_t* __restrict__ partial, const int* __restrict__ qmap,
                      const int* __restrict__ nq_dev) {
  using Cfg = ScanCfg<BQ, BN, TQ, TN, KCAP>;
  constexpr int NT = Cfg::NT, TXN = Cfg::TXN, TYN = Cfg::TYN, QCAP = Cfg::QCAP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);           // [2][BQ][DKP]
  float* Xs = Qs + 2 * BQ * DKP;                            // [2][BN][DKP]
  uint64_t* list = reinterpret_cast<uint64_t*>(Xs + 2 * BN * DKP);  // [BQ][KCAP]
  uint64_t* queue = list + BQ * KCAP;                       // [BQ][QCAP]
  int* qcnt = reinterpret_cast<int*>(queue + BQ * QCAP);    // [BQ]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % TXN, ty = tid / TXN;
  // n_qtiles is the number of query tiles in the GRID; a CTA strides over the real tiles.
  // Fallback launches bound the grid and pass the real query count (and the compacted query
  // ids) in device memory: surplus CTAs leave before any barrier.
  const int rg = blockIdx.x / n_qtiles;
  if (nq_dev != nullptr) nq = *nq_dev;
  for (int qt = blockIdx.x % n_qtiles; (long long)qt * BQ < nq; qt += n_qtiles) {
  const long long q0 = (long long)qt * BQ;
  const long long row_begin = (long long)rg * rows_per_range;
  const long long row_end = min(n_rows, row_begin + rows_per_range);
  const bool vec_ok = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  const int n_chunks = (dim + DK - 1) / DK;
  const int full_groups = dim / 4, tail = dim % 4;

  __syncthreads();  // previous query tile of this CTA fully written out
  for (int i = tid; i < BQ * KCAP; i += NT) list[i] = kEmptyKey;
  for (int i = tid; i < BQ; i += NT) qcnt[i] = 0;
  __syncthreads();

  for (long long x0 = row_begin; x0 < row_end; x0 += BN) {
    float acc[TQ][TN];
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    stage_chunk<BQ, NT>(Qs, q, q0, nq, dim, 0, vec_ok, tid, qmap);
    stage_chunk<BN, NT>(Xs, db, x0, row_end, dim, 0, vec_ok, tid);
    cp_async_commit();
    for (int c = 0; c < n_chunks; ++c) {
      const int st = c & 1;
      if (c + 1 < n_chunks) {
        stage_chunk<BQ, NT>(Qs + (st ^ 1) * BQ * DKP, q, q0, nq, dim, (c + 1) * DK, vec_ok, tid,
                            qmap);
        stage_chunk<BN, NT>(Xs + (st ^ 1) * BN * DKP, db, x0, row_end, dim, (c + 1) * DK, vec_ok,
                            tid);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* Qc = Qs + st * BQ * DKP;
      const float* Xc = Xs + st * BN * DKP;
#pragma unroll 2
      for (int g = 0; g < DK / 4; ++g) {
        const int gg = c * (DK / 4) + g;
        if (gg < full_groups) {
          float4 qa[TQ], xb[TN];
#pragma unroll
          for (int i = 0; i < TQ; ++i)
            qa[i] = *reinterpret_cast<const float4*>(Qc + (ty + i * TYN) * DKP + g * 4);
#pragma unroll
          for (int j = 0; j < TN; ++j)
            xb[j] = *reinterpret_cast<const float4*>(Xc + (tx + j * TXN) * DKP + g * 4);
#pragma unroll
          for (int i = 0; i < TQ; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
              const float d0 = __fsub_rn(qa[i].x, xb[j].x);
              const float d1 = __fsub_rn(qa[i].y, xb[j].y);
              const float d2 = __fsub_rn(qa[i].z, xb[j].z);
              const float d3 = __fsub_rn(qa[i].w, xb[j].w);
              float s = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
              s = __fadd_rn(s, __fmul_rn(d2, d2));
              s = __fadd_rn(s, __fmul_rn(d3, d3));
              acc[i][j] = __fadd_rn(acc[i][j], s);
            }
        } else if (gg == full_groups && tail > 0) {
          // nanoflann.hpp:481-485: the last 0-3 components, one by one
          for (int t = 0; t < tail; ++t) {
#pragma unroll
            for (int i = 0; i < TQ; ++i)
#pragma unroll
              for (int j = 0; j < TN; ++j) {
                const float d = __fsub_rn(Qc[(ty + i * TYN) * DKP + g * 4 + t],
                                          Xc[(tx + j * TXN) * DKP + g * 4 + t]);
                acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(d, d));
              }
          }
        }
      }
      __syncthreads();
    }

    // Fused top-k: TN phases; in phase j every thread offers column tx + j*TXN.
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const long long row = x0 + tx + j * TXN;
      int pushed = 0;
      if (row < row_end) {
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
          const int qr = ty + i * TYN;
          if (q0 + qr < nq) {
            const uint64_t key = pack_key(acc[i][j], (uint32_t)row);
            if (key < list[qr * KCAP + k - 1]) {
              const int pos = atomicAdd(&qcnt[qr], 1);
              queue[qr * QCAP + pos] = key;
              pushed = 1;
            }
          }
        }
      }
      if (__syncthreads_or(pushed)) {
        for (int qr = warp; qr < BQ; qr += NT / 32) {
          const int cnt = qcnt[qr];
          if (cnt > 0) {
            for (int b = 0; b < cnt; b += 32) {
              const int c2 = min(32, cnt - b);
              const uint64_t e = (lane < c2) ? queue[qr * QCAP + b + lane] : kEmptyKey;
              warp_merge_into_list<KCAP>(list + qr * KCAP, k, e, c2, lane);
            }
            if (lane == 0) qcnt[qr] = 0;
          }
        }
        __syncthreads();
      }
    }
  }

  // partial[q][range][k]
  for (int i = tid; i < BQ * k; i += NT) {
    const int qr = i / k, s = i % k;
    if (q0 + qr < nq)
      partial[((size_t)(q0 + qr) * n_ranges + rg) * k + s] = list[qr * KCAP + s];
  }
  }  // query-tile loop
}

// ---------------------------------------------------------------------------
// Merge n_lists ascending key lists per query into the final (idx, d2) arrays.
// One warp per query; rank of an element = its position in its own list + the
// lower bounds in every other list (keys are distinct: different rows).
__global__ void knn_finalize_kernel(const uint64_t* __restrict__ partial, int nq, int n_lists,
                                    int k, uint64_t idx_offset, uint64_t* __restrict__ out_idx,
                                    float* __restrict__ out_d2, const int* __restrict__ qmap,
                                    const int* __restrict__ nq_dev) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (nq_dev != nullptr) nq = *nq_dev;
  if (warp >= nq) return;
  const uint64_t* P = partial + (size_t)warp * n_lists * k;
  const size_t orow = qmap ? (size_t)qmap[warp] : (size_t)warp;
  uint64_t* oi = out_idx + orow * k;
  float* od = out_d2 + orow * k;
  for (int i = lane; i < k; i += 32) {
    oi[i] = 0xFFFFFFFFFFFFFFFFull;
    od[i] = 3.402823466e+38f;  // FLT_MAX
  }
  __syncwarp();
  const int total = n_lists * k;
  for (int e = lane; e < total; e += 32) {
    const uint64_t key = P[e];
    if (key == kEmptyKey) continue;
    const int seg = e / k;
    int rank = e % k;
    for (int s = 0; s < n_lists && rank < k; ++s) {
      if (s == seg) continue;
      const uint64_t* Ls = P + (size_t)s * k;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (Ls[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      oi[rank] = (uint64_t)(uint32_t)(key & 0xFFFFFFFFull) + idx_offset;
      od[rank] = __uint_as_float((uint32_t)(key >> 32));
    }
  }
}

__device__ __forceinline__ bool pair_less(float da, uint64_t ia, float db, uint64_t ib) {
  return da < db || (da == db && ia < ib);
}

// K4: merge g shard lists [g][nq][k] of (idx, d2), each ascending by (d2, idx),
// UINT64_MAX = empty slot.  One warp per query.
__global__ void knn_merge_pairs_kernel(const uint64_t* __restrict__ idx,
                                       const float* __restrict__ d2, int g, int nq, int k,
                                       uint64_t* __restrict__ out_idx,
                                       float* __restrict__ out_d2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nq) return;
  uint64_t* oi = out_idx + (size_t)warp * k;
  float* od = out_d2 + (size_t)warp * k;
  for (int i = lane; i < k; i += 32) {
    oi[i] = 0xFFFFFFFFFFFFFFFFull;
    od[i] = 3.402823466e+38f;
  }
  __syncwarp();
  const int total = g * k;
  for (int e = lane; e < total; e += 32) {
    const int seg = e / k, pos = e % k;
    const size_t base = ((size_t)seg * nq + warp) * k;
    const uint64_t ii = idx[base + pos];
    if (ii == 0xFFFFFFFFFFFFFFFFull) continue;
    const float dd = d2[base + pos];
    int rank = pos;
    for (int s = 0; s < g && rank < k; ++s) {
      if (s == seg) continue;
      const size_t b2 = ((size_t)s * nq + warp) * k;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint64_t im = idx[b2 + mid];
        const bool less = (im != 0xFFFFFFFFFFFFFFFFull) && pair_less(d2[b2 + mid], im, dd, ii);
        if (less) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      oi[rank] = ii;
      od[rank] = dd;
    }
  }
}


// K4 fused with its collective (row-sharded search): every rank's local top-k lists
// sit in a buffer that every other rank can address over NVLink (peer memory).  One kernel per
// rank (a) tells every peer "my lists of this epoch are complete" with a remote store into the
// peer's flag word -- stream order has put the lists in memory before this kernel started --,
// (b) waits until every peer has said the same, (c) pulls the lists of ITS queries straight out of
// the peers' buffers into shared memory and (d) merges them.  No NCCL call, no staging copy.
//   fbufs[p]  rank p's flag words (64 x u32) as this rank addresses them
//   bufs[p]   rank p's list buffer of this call (two alternate: a rank may be one call ahead of a
//             peer that is still reading): [256 B][idx][d2]
//   block     first list of this rank's queries inside every peer's buffer (sliced batch:
//             rank * nq; the same query everywhere: 0)
// One warp per query; lists of k <= 128 entries.
__global__ void knn_p2p_gather_merge_kernel(void* const* __restrict__ fbufs, void* const* __restrict__ bufs,
                                            int n_ranks, int rank,
                                            unsigned epoch, size_t idx_off, size_t d2_off, size_t block,
                                            int nq, int k, uint64_t* __restrict__ out_idx,
                                            float* __restrict__ out_d2, int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char p2p_smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (a) signal: idempotent, so every block does it -- no block waits on another block of this grid
  if (threadIdx.x < n_ranks) {
    __threadfence_system();
    volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(fbufs[threadIdx.x]) + rank;
    *flag = epoch;
  }
  // (b) wait for every peer (bounded: a peer that died must not hang this GPU)
  if (threadIdx.x < n_ranks) {
    volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(fbufs[rank]) + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(*mine - epoch) < 0) {
      if (clock64() - t0 > 8000000000ll) {   // ~4 s
        atomicExch(err, 1);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const int q = blockIdx.x * wpb + w;
  if (q >= nq) return;
  uint64_t* li = reinterpret_cast<uint64_t*>(p2p_smem) + (size_t)w * n_ranks * k;
  float* ld = reinterpret_cast<float*>(p2p_smem + (size_t)wpb * n_ranks * k * 8) + (size_t)w * n_ranks * k;
  // (c) gather: k contiguous entries per peer, read around the caches (the buffers are rewritten every call)
  for (int p = 0; p < n_ranks; ++p) {
    const char* base = reinterpret_cast<const char*>(bufs[p]);
    const uint64_t* pi = reinterpret_cast<const uint64_t*>(base + idx_off) + (block + q) * k;
    const float* pd = reinterpret_cast<const float*>(base + d2_off) + (block + q) * k;
    for (int i = lane; i < k; i += 32) {
      li[p * k + i] = __ldcv(pi + i);
      ld[p * k + i] = __ldcv(pd + i);
    }
  }
  __syncwarp();
  // (d) merge by rank counting, as knn_merge_pairs_kernel
  uint64_t* oi = out_idx + (size_t)q * k;
  float* od = out_d2 + (size_t)q * k;
  for (int i = lane; i < k; i += 32) {
    oi[i] = 0xFFFFFFFFFFFFFFFFull;
    od[i] = 3.402823466e+38f;
  }
  __syncwarp();
  const int total = n_ranks * k;
  for (int e = lane; e < total; e += 32) {
    const int seg = e / k, pos = e % k;
    const uint64_t ii = li[e];
    if (ii == 0xFFFFFFFFFFFFFFFFull) continue;
    const float dd = ld[e];
    int r = pos;
    for (int s2 = 0; s2 < n_ranks && r < k; ++s2) {
      if (s2 == seg) continue;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint64_t im = li[s2 * k + mid];
        const bool less = (im != 0xFFFFFFFFFFFFFFFFull) && pair_less(ld[s2 * k + mid], im, dd, ii);
        if (less) lo = mid + 1; else hi = mid;
      }
      r += lo;
    }
    if (r < k) {
      oi[r] = ii;
      od[r] = dd;
    }
  }
}

template <int BQ, int BN, int TQ, int TN, int KCAP>
cudaError_t launch_scan(const float* db, long long n_rows, int dim, const float* q, int nq, int k,
                        int n_ranges, long long rows_per_range, uint64_t* partial,
                        const int* qmap, const int* nq_dev, cudaStream_t stream) {
  using Cfg = ScanCfg<BQ, BN, TQ, TN, KCAP>;
  auto kern = knn_exact_scan_kernel<BQ, BN, TQ, TN, KCAP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::kSmem);
  if (e != cudaSuccess) return e;
  int n_qtiles = (nq + BQ - 1) / BQ;
  if (nq_dev != nullptr) n_qtiles = std::min(n_qtiles, 64);  // fallback: bounded grid, CTAs stride
  kern<<<n_qtiles * n_ranges, Cfg::NT, Cfg::kSmem, stream>>>(db, n_rows, dim, q, nq, k, n_qtiles,
                                                            rows_per_range, n_ranges, partial, qmap,
                                                            nq_dev);
  return cudaGetLastError();
}

}  // namespace

int exact_scan_tile_q(int nq) { return nq <= 64 ? 16 : 128; }
int exact_scan_tile_n(int nq) { return nq <= 64 ? 256 : 128; }

cudaError_t launch_knn_exact_scan(const float* db, long long n_rows, int dim, const float* q,
                                  int nq, int k, int n_ranges, long long rows_per_range,
                                  uint64_t* partial, cudaStream_t stream, const int* qmap,
                                  const int* nq_dev, bool force_small) {
  const bool small = force_small || nq <= 64;
#define GLOC_SCAN(KC)                                                                       \
  (small ? launch_scan<16, 256, 4, 4, KC>(db, n_rows, dim, q, nq, k, n_ranges,              \
                                          rows_per_range, partial, qmap, nq_dev, stream)    \
         : launch_scan<128, 128, 8, 8, KC>(db, n_rows, dim, q, nq, k, n_ranges,             \
                                           rows_per_range, partial, qmap, nq_dev, stream))
  if (k <= 32) return GLOC_SCAN(32);
  if (k <= 64) return GLOC_SCAN(64);
  return GLOC_SCAN(128);
#undef GLOC_SCAN
}

cudaError_t launch_knn_finalize(const uint64_t* partial, int nq, int n_lists, int k,
                                uint64_t idx_offset, uint64_t* out_idx, float* out_d2,
                                cudaStream_t stream, const int* qmap, const int* nq_dev) {
  const int threads = 128, wpb = threads / 32;
  knn_finalize_kernel<<<(nq + wpb - 1) / wpb, threads, 0, stream>>>(
      partial, nq, n_lists, k, idx_offset, out_idx, out_d2, qmap, nq_dev);
  return cudaGetLastError();
}

cudaError_t launch_knn_merge_pairs(const uint64_t* idx, const float* d2, int g, int nq, int k,
                                   uint64_t* out_idx, float* out_d2, cudaStream_t stream) {
  const int threads = 128, wpb = threads / 32;
  knn_merge_pairs_kernel<<<(nq + wpb - 1) / wpb, threads, 0, stream>>>(idx, d2, g, nq, k, out_idx,
                                                                       out_d2);
  return cudaGetLastError();
}

cudaError_t launch_knn_p2p_gather_merge(void* const* d_fbufs, void* const* d_bufs, int n_ranks, int rank, unsigned epoch, size_t idx_off,
                                        size_t d2_off, size_t block, int nq, int k, uint64_t* out_idx,
                                        float* out_d2, int* err, cudaStream_t stream) {
  const int threads = 128, wpb = threads / 32;
  const size_t smem = (size_t)wpb * n_ranks * k * 12;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  knn_p2p_gather_merge_kernel<<<(nq + wpb - 1) / wpb, threads, smem, stream>>>(
      d_fbufs, d_bufs, n_ranks, rank, epoch, idx_off, d2_off, block, nq, k, out_idx, out_d2, err);
  return cudaGetLastError();
}

}  // namespace gloc
