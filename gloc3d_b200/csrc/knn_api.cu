// knn_api.cu -- C ABI of stage 1 (retrieval).  See include/gloc3d.h for the
// reference interfaces each entry point replaces.
#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "comm.cuh"
#include "knn_kernels.cuh"
#include "knn_shortlist.cuh"

namespace gloc {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int sm_count(int device) {
  static std::mutex mu;
  static std::vector<int> cache;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= device) cache.resize(device + 1, 0);
  if (cache[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0)
      v = 148;
    cache[device] = v;
  }
  return cache[device];
}

// Device buffer that only grows.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) {
      cudaError_t e = cudaFree(p);
      p = nullptr;
      bytes = 0;
      if (e != cudaSuccess) return e;
    }
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      e = cudaMalloc(&p, need);
      want = need;
    }
    if (e == cudaSuccess) bytes = want; else p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace gloc

using namespace gloc;

struct gloc_knn_index {
  int device = 0;
  size_t dim = 0;
  float* d_db = nullptr;  // row-major n x dim float32 (the reference's KeyMat, contiguous)
  size_t n = 0, cap = 0;
  size_t search_limit = SIZE_MAX;
  uint64_t offset = 0;
  int mode = GLOC_KNN_AUTO;
  cudaStream_t stream = nullptr;  // used by the host-buffer entry points
  cudaStream_t copy_stream = nullptr;   // host<->device copies of large batches overlap the search
  cudaEvent_t ev_in[8] = {}, ev_out[8] = {};
  DevBuf partial;                 // exact-scan per-range lists
  DevBuf flag;                    // streaming scan: merge-overflow flag
  DevBuf stage_q, stage_idx, stage_d2;
  DevBuf sh_allq, sh_lidx, sh_ld2, sh_gidx, sh_gd2;   // row-sharded search: gathered queries, per-shard lists
  // peer-memory exchange (NVLink): this rank's lists where every rank can read them
  struct {
    int state = 0;               // 0 untried, 1 in use, -1 unavailable (NCCL exchange instead)
    gloc_comm* comm = nullptr;
    void* buf[2] = {nullptr, nullptr};
    void* d_ptrs = nullptr;      // device: [2][size] peer pointers (list buffers) + [size] (flags = buffer 0)
    int* d_err = nullptr;
    size_t cap_entries = 0;
    unsigned epoch = 0;
  } p2p;
  ShortlistState* sl = nullptr;   // tensor-shortlist state (bf16 copy, norms, workspaces)
  gloc_knn_stats stats{};
  EventProfiler prof;
};

namespace {

struct RangePlan {
  int n_ranges;
  long long rows_per_range;
};

// Split the searchable rows into ranges so that (query tiles x ranges) fills the SMs in
// whole waves; each (query tile, range) unit is one CTA of the exact scan.
RangePlan plan_ranges(long long n_search, int nq, int device) {
  const int BQ = exact_scan_tile_q(nq), BN = exact_scan_tile_n(nq);
  const int sms = sm_count(device);
  const long long n_qtiles = (nq + BQ - 1) / BQ;
  const long long max_ranges = std::max<long long>(1, std::min<long long>(64, n_search / (BN * 4LL)));
  double best_score = -1.0;
  long long best_r = 1;
  for (long long r = 1; r <= max_ranges; ++r) {
    const long long units = n_qtiles * r;
    const long long waves = (units + sms - 1) / sms;
    const double eff = (double)units / (double)(waves * sms);
    const double score = eff - 0.003 * (double)r;
    if (score > best_score + 1e-12) {
      best_score = score;
      best_r = r;
    }
  }
  long long rpr = (n_search + best_r - 1) / best_r;
  rpr = (rpr + BN - 1) / BN * BN;
  RangePlan p;
  p.rows_per_range = rpr;
  p.n_ranges = (int)((n_search + rpr - 1) / rpr);
  return p;
}

int exact_query_device(gloc_knn_index* ix, const float* d_q, size_t nq, size_t k,
                       uint64_t* d_idx, float* d_d2, size_t n_search, cudaStream_t stream) {
  // up to 16 queries: passes over the rows at HBM speed, four queries per pass (knn_stream.cu);
  // beyond that the register-tiled exact scan amortises the row reads better
  if (nq <= 16 && stream_applicable(ix->dim, std::min<size_t>(nq, 4), k) &&
      (reinterpret_cast<uintptr_t>(d_q) & 15) == 0 && std::getenv("GLOC_KNN_NO_STREAM") == nullptr) {
    const int grid = stream_grid(ix->device, ix->dim);
    GLOC_CUDA_TRY(ix->partial.reserve(4 * (size_t)grid * k * sizeof(uint64_t)));
    GLOC_CUDA_TRY(ix->flag.reserve(16));
    for (size_t q0 = 0; q0 < nq; q0 += 4) {
      const int cq = (int)std::min<size_t>(4, nq - q0);
      GLOC_CUDA_TRY(cudaMemsetAsync(ix->flag.p, 0, 4, stream));
      GLOC_CUDA_TRY(launch_knn_stream(ix->d_db, (long long)n_search, (int)ix->dim, d_q + q0 * ix->dim, cq,
                                      (int)k, grid, (uint64_t*)ix->partial.p, ix->offset, d_idx + q0 * k,
                                      d_d2 + q0 * k, (int*)ix->flag.p, &ix->prof, stream));
      ix->stats.kernel_launches += 3;
    }
    return GLOC_OK;
  }
  const size_t kChunk = 1u << 17;
  for (size_t q0 = 0; q0 < nq; q0 += kChunk) {
    const int cq = (int)std::min(kChunk, nq - q0);
    const RangePlan plan = plan_ranges((long long)n_search, cq, ix->device);
    GLOC_CUDA_TRY(ix->partial.reserve((size_t)cq * plan.n_ranges * k * sizeof(uint64_t)));
    ix->prof.begin(stream);
    cudaError_t le = launch_knn_exact_scan(ix->d_db, (long long)n_search, (int)ix->dim,
                                           d_q + q0 * ix->dim, cq, (int)k, plan.n_ranges,
                                           plan.rows_per_range, (uint64_t*)ix->partial.p, stream);
    ix->prof.end(stream);
    GLOC_CUDA_TRY(le);
    GLOC_CUDA_TRY(launch_knn_finalize((const uint64_t*)ix->partial.p, cq, plan.n_ranges, (int)k,
                                      ix->offset, d_idx + q0 * k, d_d2 + q0 * k, stream));
    ix->stats.kernel_launches += 2;
  }
  return GLOC_OK;
}

int grow_db(gloc_knn_index* ix, size_t need_rows) {
  if (need_rows <= ix->cap) return GLOC_OK;
  size_t new_cap = std::max(need_rows, ix->cap + ix->cap / 2);
  float* nd = nullptr;
  cudaError_t e = cudaMalloc((void**)&nd, new_cap * ix->dim * sizeof(float));
  if (e != cudaSuccess && new_cap != need_rows) {
    (void)cudaGetLastError();
    new_cap = need_rows;
    e = cudaMalloc((void**)&nd, new_cap * ix->dim * sizeof(float));
  }
  if (e != cudaSuccess) return fail(GLOC_ERR_NOMEM, std::string("cudaMalloc(db): ") + cudaGetErrorString(e));
  if (ix->n > 0) {
    e = cudaMemcpy(nd, ix->d_db, ix->n * ix->dim * sizeof(float), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);   // asynchronous to the host otherwise
    if (e != cudaSuccess) {
      cudaFree(nd);
      return fail(GLOC_ERR_CUDA, std::string("cudaMemcpy(db grow): ") + cudaGetErrorString(e));
    }
  }
  if (ix->d_db) cudaFree(ix->d_db);
  ix->d_db = nd;
  ix->cap = new_cap;
  return GLOC_OK;
}

int put_rows(gloc_knn_index* ix, const float* rows, size_t n, size_t at, cudaMemcpyKind kind) {
  if (n == 0) return GLOC_OK;
  int rc = grow_db(ix, at + n);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpy(ix->d_db + at * ix->dim, rows, n * ix->dim * sizeof(float), kind));
  // device-to-device and small pageable copies return before the data has landed, and the
  // searches run on non-blocking streams that do not order against the legacy stream
  GLOC_CUDA_TRY(cudaStreamSynchronize(0));
  return GLOC_OK;
}

}  // namespace

extern "C" {

int gloc_version(void) { return 100; }
const char* gloc_last_error(void) { return g_last_error.c_str(); }

int gloc_knn_pair_workers(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    (void)cudaGetLastError();
    return 0;
  }
  gloc::DeviceGuard g(device);
  return gloc::shortlist_pair_workers();
}

int gloc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
        major == 10)
      ++ok;
  }
  return ok;
}

int gloc_knn_create(gloc_knn_index** out, size_t dim, int device) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_knn_create: out is null");
  *out = nullptr;
  if (dim == 0 || dim > (1u << 20)) return fail(GLOC_ERR_INVALID, "gloc_knn_create: bad dim");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_knn_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(GLOC_ERR_INVALID, "gloc_knn_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return fail(GLOC_ERR_CUDA, "gloc_knn_create: device is not sm_100 (kernels are sm_100a only)");
  DeviceGuard g(device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_create: cudaSetDevice failed");
  gloc_knn_index* ix = new (std::nothrow) gloc_knn_index;
  if (!ix) return fail(GLOC_ERR_NOMEM, "gloc_knn_create: out of host memory");
  ix->device = device;
  ix->dim = dim;
  e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&ix->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->ev_out[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    delete ix;
    return fail(GLOC_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = ix;
  return GLOC_OK;
}

void gloc_knn_destroy(gloc_knn_index* ix) {
  if (!ix) return;
  DeviceGuard g(ix->device);
  if (ix->stream) {
    cudaStreamSynchronize(ix->stream);
    cudaStreamDestroy(ix->stream);
  }
  if (ix->copy_stream) {
    cudaStreamSynchronize(ix->copy_stream);
    cudaStreamDestroy(ix->copy_stream);
  }
  for (int i = 0; i < 8; ++i) {
    if (ix->ev_in[i]) cudaEventDestroy(ix->ev_in[i]);
    if (ix->ev_out[i]) cudaEventDestroy(ix->ev_out[i]);
  }
  if (ix->d_db) cudaFree(ix->d_db);
  ix->partial.release();
  ix->flag.release();
  ix->stage_q.release();
  ix->stage_idx.release();
  ix->stage_d2.release();
  for (DevBuf* b : {&ix->sh_allq, &ix->sh_lidx, &ix->sh_ld2, &ix->sh_gidx, &ix->sh_gd2}) b->release();
  // (the peers' mappings of p2p.buf are closed by gloc_comm_destroy; destroy the index after a
  // barrier of the host program, as with any buffer peers may still read)
  for (void* b : ix->p2p.buf)
    if (b) {
      if (ix->p2p.comm) comm_unmap_peers(ix->p2p.comm, b);
      cudaFree(b);
    }
  if (ix->p2p.d_ptrs) cudaFree(ix->p2p.d_ptrs);
  if (ix->p2p.d_err) cudaFree(ix->p2p.d_err);
  shortlist_destroy(ix->sl);
  delete ix;
}

int gloc_knn_set_db(gloc_knn_index* ix, const float* rows, size_t n) {
  if (!ix || (n > 0 && !rows)) return fail(GLOC_ERR_INVALID, "gloc_knn_set_db: null argument");
  DeviceGuard g(ix->device);
  GLOC_CUDA_TRY(cudaDeviceSynchronize());
  ix->n = 0;
  shortlist_invalidate(ix->sl, 0);
  int rc = put_rows(ix, rows, n, 0, cudaMemcpyHostToDevice);
  if (rc == GLOC_OK) ix->n = n;
  return rc;
}

int gloc_knn_set_db_device(gloc_knn_index* ix, const float* d_rows, size_t n) {
  if (!ix || (n > 0 && !d_rows)) return fail(GLOC_ERR_INVALID, "gloc_knn_set_db_device: null argument");
  DeviceGuard g(ix->device);
  GLOC_CUDA_TRY(cudaDeviceSynchronize());
  ix->n = 0;
  shortlist_invalidate(ix->sl, 0);
  int rc = put_rows(ix, d_rows, n, 0, cudaMemcpyDeviceToDevice);
  if (rc == GLOC_OK) ix->n = n;
  return rc;
}

int gloc_knn_append(gloc_knn_index* ix, const float* rows, size_t n) {
  if (!ix || (n > 0 && !rows)) return fail(GLOC_ERR_INVALID, "gloc_knn_append: null argument");
  DeviceGuard g(ix->device);
  GLOC_CUDA_TRY(cudaDeviceSynchronize());
  int rc = put_rows(ix, rows, n, ix->n, cudaMemcpyHostToDevice);
  if (rc == GLOC_OK) {
    shortlist_invalidate(ix->sl, ix->n);  // rows [0, n) keep their derived data
    ix->n += n;
  }
  return rc;
}

size_t gloc_knn_size(const gloc_knn_index* ix) { return ix ? ix->n : 0; }
size_t gloc_knn_dim(const gloc_knn_index* ix) { return ix ? ix->dim : 0; }

int gloc_knn_set_search_limit(gloc_knn_index* ix, size_t n_search) {
  if (!ix) return fail(GLOC_ERR_INVALID, "gloc_knn_set_search_limit: null index");
  ix->search_limit = n_search;
  return GLOC_OK;
}

int gloc_knn_set_index_offset(gloc_knn_index* ix, uint64_t offset) {
  if (!ix) return fail(GLOC_ERR_INVALID, "gloc_knn_set_index_offset: null index");
  ix->offset = offset;
  return GLOC_OK;
}

int gloc_knn_set_mode(gloc_knn_index* ix, int mode) {
  if (!ix || mode < GLOC_KNN_AUTO || mode > GLOC_KNN_SHORTLIST)
    return fail(GLOC_ERR_INVALID, "gloc_knn_set_mode: bad argument");
  ix->mode = mode;
  return GLOC_OK;
}

int gloc_knn_get_stats(const gloc_knn_index* ix, gloc_knn_stats* stats) {
  if (!ix || !stats) return fail(GLOC_ERR_INVALID, "gloc_knn_get_stats: null argument");
  *stats = ix->stats;
  if (ix->sl) {  // device-side counters of the shortlist path (synchronises the device)
    DeviceGuard g(ix->device);
    uint64_t rows = 0, ovf = 0;
    int rc = shortlist_counters(ix->sl, &rows, &ovf);
    if (rc != GLOC_OK) return rc;
    stats->shortlist_rows = rows;
    stats->fallback_queries = ovf;
  }
  return GLOC_OK;
}

int gloc_knn_set_profiling(gloc_knn_index* ix, int enabled) {
  if (!ix) return fail(GLOC_ERR_INVALID, "gloc_knn_set_profiling: null index");
  ix->prof.enabled = enabled != 0;
  return GLOC_OK;
}

int gloc_knn_get_profile(gloc_knn_index* ix, gloc_profile* out) {
  if (!ix || !out) return fail(GLOC_ERR_INVALID, "gloc_knn_get_profile: null argument");
  DeviceGuard g(ix->device);
  ix->prof.collect(&out->dominant_ms, &out->dominant_launches);
  return GLOC_OK;
}

int gloc_knn_query_device(gloc_knn_index* ix, const float* d_q, size_t nq, size_t k,
                          uint64_t* d_idx, float* d_d2, void* stream_v) {
  if (!ix) return fail(GLOC_ERR_INVALID, "gloc_knn_query: null index");
  if (nq == 0) return GLOC_OK;
  if (!d_q || !d_idx || !d_d2) return fail(GLOC_ERR_INVALID, "gloc_knn_query: null buffer");
  if (k == 0 || k > 128) return fail(GLOC_ERR_RANGE, "gloc_knn_query: k must be in [1, 128]");
  const size_t n_search = std::min(ix->n, ix->search_limit);
  if (n_search == 0)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_knn_query: the index holds no searchable rows");
  if (n_search >= (1ull << 32))
    return fail(GLOC_ERR_RANGE, "gloc_knn_query: more than 2^32-1 rows per index; shard the database");
  if (nq > (size_t)INT32_MAX) return fail(GLOC_ERR_RANGE, "gloc_knn_query: too many queries in one call");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_query: cudaSetDevice failed");
  cudaStream_t stream = (cudaStream_t)stream_v;
  int mode = ix->mode;
  if (mode == GLOC_KNN_AUTO)
    mode = shortlist_applicable(ix->dim, n_search, nq, k) ? GLOC_KNN_SHORTLIST : GLOC_KNN_EXACT_SCAN;
  int rc;
  if (mode == GLOC_KNN_SHORTLIST) {
    if (!shortlist_supported(ix->dim, k))
      return fail(GLOC_ERR_RANGE, "gloc_knn_query: GLOC_KNN_SHORTLIST needs dim % 64 == 0, dim <= 512, k <= 32 (GLOC_KNN_AUTO takes the exact scan otherwise)");
    ShortlistArgs a;
    a.device = ix->device;
    a.d_db = ix->d_db;
    a.n_rows = n_search;
    a.n_total = ix->n;
    a.dim = ix->dim;
    a.d_q = d_q;
    a.nq = nq;
    a.k = k;
    a.offset = ix->offset;
    a.d_idx = d_idx;
    a.d_d2 = d_d2;
    a.stream = stream;
    a.prof = &ix->prof;
    uint64_t launches = 0, fallback = 0, rows = 0;
    rc = shortlist_query(&ix->sl, a, &launches, &fallback, &rows);
    if (rc != GLOC_OK) return rc;
    ix->stats.kernel_launches += launches;
    ix->stats.shortlist_queries += nq;
  } else {
    rc = exact_query_device(ix, d_q, nq, k, d_idx, d_d2, n_search, stream);
    if (rc != GLOC_OK) return rc;
  }
  ix->stats.queries += nq;
  ix->stats.last_mode = (uint64_t)mode;
  return GLOC_OK;
}

int gloc_knn_query(gloc_knn_index* ix, const float* q, size_t nq, size_t k, uint64_t* out_idx,
                   float* out_d2) {
  if (!ix) return fail(GLOC_ERR_INVALID, "gloc_knn_query: null index");
  if (nq == 0) return GLOC_OK;
  if (!q || !out_idx || !out_d2) return fail(GLOC_ERR_INVALID, "gloc_knn_query: null buffer");
  if (k == 0 || k > 128) return fail(GLOC_ERR_RANGE, "gloc_knn_query: k must be in [1, 128]");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_query: cudaSetDevice failed");
  GLOC_CUDA_TRY(ix->stage_q.reserve(nq * ix->dim * sizeof(float)));
  GLOC_CUDA_TRY(ix->stage_idx.reserve(nq * k * sizeof(uint64_t)));
  GLOC_CUDA_TRY(ix->stage_d2.reserve(nq * k * sizeof(float)));
  // Large batches are cut into up to 8 chunks: the upload of chunk i+1 and the download of
  // chunk i-1 (copy stream) overlap the search of chunk i (compute stream).  The results do
  // not depend on the chunking.
  const size_t kMinChunk = 4096;
  size_t n_chunks = std::min<size_t>(8, nq / kMinChunk);
  if (n_chunks < 2) n_chunks = 1;
  const size_t per = ((nq + n_chunks - 1) / n_chunks + 127) / 128 * 128;
  const float* dq = (const float*)ix->stage_q.p;
  uint64_t* di = (uint64_t*)ix->stage_idx.p;
  float* dd = (float*)ix->stage_d2.p;
  cudaStream_t cs = n_chunks > 1 ? ix->copy_stream : ix->stream;
  size_t c = 0;
  for (size_t q0 = 0; q0 < nq; q0 += per, ++c) {
    const size_t n = std::min(per, nq - q0);
    GLOC_CUDA_TRY(cudaMemcpyAsync((void*)(dq + q0 * ix->dim), q + q0 * ix->dim, n * ix->dim * sizeof(float),
                                  cudaMemcpyHostToDevice, cs));
    if (n_chunks > 1) GLOC_CUDA_TRY(cudaEventRecord(ix->ev_in[c], cs));
  }
  c = 0;
  for (size_t q0 = 0; q0 < nq; q0 += per, ++c) {
    const size_t n = std::min(per, nq - q0);
    if (n_chunks > 1) GLOC_CUDA_TRY(cudaStreamWaitEvent(ix->stream, ix->ev_in[c], 0));
    int rc = gloc_knn_query_device(ix, dq + q0 * ix->dim, n, k, di + q0 * k, dd + q0 * k, ix->stream);
    if (rc != GLOC_OK) return rc;
    if (n_chunks > 1) {
      GLOC_CUDA_TRY(cudaEventRecord(ix->ev_out[c], ix->stream));
      GLOC_CUDA_TRY(cudaStreamWaitEvent(cs, ix->ev_out[c], 0));
    }
    GLOC_CUDA_TRY(cudaMemcpyAsync(out_idx + q0 * k, di + q0 * k, n * k * sizeof(uint64_t),
                                  cudaMemcpyDeviceToHost, cs));
    GLOC_CUDA_TRY(cudaMemcpyAsync(out_d2 + q0 * k, dd + q0 * k, n * k * sizeof(float),
                                  cudaMemcpyDeviceToHost, cs));
  }
  GLOC_CUDA_TRY(cudaStreamSynchronize(cs));
  if (n_chunks > 1) GLOC_CUDA_TRY(cudaStreamSynchronize(ix->stream));
  return GLOC_OK;
}

int gloc_knn_merge_topk_device(const uint64_t* d_idx, const float* d_d2, size_t g, size_t nq,
                               size_t k, uint64_t* d_out_idx, float* d_out_d2, int device,
                               void* stream) {
  if (nq == 0 || g == 0) return GLOC_OK;
  if (!d_idx || !d_d2 || !d_out_idx || !d_out_d2)
    return fail(GLOC_ERR_INVALID, "gloc_knn_merge_topk_device: null buffer");
  if (k == 0 || k > 4096 || g > 4096 || nq > (size_t)INT32_MAX)
    return fail(GLOC_ERR_RANGE, "gloc_knn_merge_topk_device: bad sizes");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_merge_topk_device: cudaSetDevice failed");
  GLOC_CUDA_TRY(launch_knn_merge_pairs(d_idx, d_d2, (int)g, (int)nq, (int)k, d_out_idx, d_out_d2,
                                       (cudaStream_t)stream));
  return GLOC_OK;
}
}  // extern "C"

namespace {

// Peer-memory list buffers of a row-sharded index: (re)allocated collectively -- every rank sees the
// same (ranks, queries, k), so every rank grows in the same call.  Returns false when the GPUs
// cannot address each other (then the lists travel through NCCL instead).
bool p2p_prepare(gloc_knn_index* ix, gloc_comm* comm, size_t entries, cudaStream_t stream) {
  auto& P = ix->p2p;
  if (P.state < 0 || std::getenv("GLOC_SHARD_NO_P2P") != nullptr) return false;
  if (P.state == 1 && P.comm == comm && entries <= P.cap_entries) return true;
  // growth (or first use): nobody may still be reading the old buffers
  cudaStreamSynchronize(stream);
  const size_t cap = std::max<size_t>(entries + entries / 4, 65536);
  const size_t bytes = 256 + cap * 12;
  void* nb[2] = {nullptr, nullptr};
  void** peers[2] = {nullptr, nullptr};
  bool ok = cudaMalloc(&nb[0], bytes) == cudaSuccess && cudaMalloc(&nb[1], bytes) == cudaSuccess;
  if (ok) ok = cudaMemset(nb[0], 0, 256) == cudaSuccess && cudaMemset(nb[1], 0, 256) == cudaSuccess &&
               cudaDeviceSynchronize() == cudaSuccess;
  // the mapping is collective: called even after a local failure (with a null pointer it fails everywhere)
  int rc0 = comm_map_peers(comm, ok ? nb[0] : nullptr, &peers[0]);
  int rc1 = rc0 == GLOC_OK ? comm_map_peers(comm, nb[1], &peers[1]) : rc0;
  if (rc0 != GLOC_OK || rc1 != GLOC_OK) {
    if (rc0 == GLOC_OK) comm_unmap_peers(comm, nb[0]);
    for (void* b : nb)
      if (b) cudaFree(b);
    (void)cudaGetLastError();
    P.state = -1;
    return false;
  }
  for (void* b : P.buf)
    if (b) {
      comm_unmap_peers(P.comm, b);
      cudaFree(b);
    }
  const size_t n = (size_t)comm->size;
  if (!P.d_ptrs && cudaMalloc(&P.d_ptrs, 3 * 64 * sizeof(void*)) != cudaSuccess) { P.state = -1; return false; }
  if (!P.d_err && (cudaMalloc((void**)&P.d_err, 4) != cudaSuccess || cudaMemset(P.d_err, 0, 4) != cudaSuccess)) { P.state = -1; return false; }
  std::vector<void*> table(3 * 64, nullptr);
  for (size_t i = 0; i < n; ++i) {
    table[i] = peers[0][i];          // list buffers of even calls
    table[64 + i] = peers[1][i];     // list buffers of odd calls
    table[128 + i] = peers[0][i];    // flag words: always buffer 0
  }
  if (cudaMemcpy(P.d_ptrs, table.data(), table.size() * sizeof(void*), cudaMemcpyHostToDevice) != cudaSuccess) {
    P.state = -1;
    return false;
  }
  P.buf[0] = nb[0];
  P.buf[1] = nb[1];
  P.cap_entries = cap;
  P.comm = comm;
  P.epoch = 0;
  P.state = 1;
  return true;
}

}  // namespace

extern "C" {

// Row-sharded exact top-k (BASELINE configs[3]): every rank holds rows
// [offset, offset + n) and calls this collectively.
//   replicated == 0: every rank passes ITS slice of the batch (nq_local queries, the same count on
//     every rank).  Queries are all-gathered over NVLink (one PCIe upload per query instead of one
//     per rank), every rank searches the whole batch on its shard, the local top-k lists go
//     straight to the rank that owns the query (all-to-all), which merges its N lists (K4): the
//     merge and the result download are sharded as well.  Output: this rank's slice.
//   replicated != 0: every rank passes the same nq queries (online localisation: one query);
//     local search, all-gather of the lists, merge on every rank.  Output: the whole result.
// Results are identical to a single-GPU search of the whole database: same (d2, idx) order.
int gloc_knn_query_sharded_device(gloc_knn_index* ix, gloc_comm* comm, const float* d_q, size_t nq,
                                  size_t k, uint64_t* d_out_idx, float* d_out_d2, int replicated,
                                  void* stream_v) {
  if (!ix || !comm) return fail(GLOC_ERR_INVALID, "gloc_knn_query_sharded: null argument");
  if (nq == 0) return GLOC_OK;
  if (!d_q || !d_out_idx || !d_out_d2) return fail(GLOC_ERR_INVALID, "gloc_knn_query_sharded: null buffer");
  if (comm->device != ix->device) return fail(GLOC_ERR_INVALID, "gloc_knn_query_sharded: communicator and index live on different devices");
  if (comm->size == 1) return gloc_knn_query_device(ix, d_q, nq, k, d_out_idx, d_out_d2, stream_v);
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_query_sharded: cudaSetDevice failed");
  cudaStream_t stream = (cudaStream_t)stream_v;
  const size_t N = (size_t)comm->size, dim = ix->dim;
  int rc;
  // tuning aid: GLOC_SHARD_TIMING=1 prints the device time of every stage of this call (rank 0)
  const bool timing = std::getenv("GLOC_SHARD_TIMING") != nullptr && comm->rank == 0;
  cudaEvent_t tev[5] = {};
  auto mark = [&](int i) {
    if (timing) {
      if (!tev[i]) cudaEventCreate(&tev[i]);
      cudaEventRecord(tev[i], stream);
    }
  };
  mark(0);
  // ---- exchange fused into the merge: the search writes its lists where the peers can read them
  //      (peer memory over NVLink), one kernel signals, waits, gathers and merges
  if (comm->size <= 64 && k <= 128 && (size_t)comm->size * k * 12 * 4 <= 48 * 1024 &&
      p2p_prepare(ix, comm, (replicated ? nq : N * nq) * k, stream)) {
    auto& P = ix->p2p;
    const unsigned epoch = ++P.epoch;
    const int par = (int)(epoch & 1u);
    char* buf = (char*)P.buf[par];
    const size_t idx_off = 256, d2_off = 256 + P.cap_entries * 8;
    const float* dq = d_q;
    size_t n_search = nq;
    if (!replicated) {
      GLOC_CUDA_TRY(ix->sh_allq.reserve(N * nq * dim * sizeof(float)));
      rc = comm_all_gather(comm, d_q, ix->sh_allq.p, nq * dim * sizeof(float), stream);
      if (rc != GLOC_OK) return rc;
      dq = (const float*)ix->sh_allq.p;
      n_search = N * nq;
    }
    mark(1);
    rc = gloc_knn_query_device(ix, dq, n_search, k, (uint64_t*)(buf + idx_off), (float*)(buf + d2_off), stream);
    if (rc != GLOC_OK) return rc;
    mark(2);
    mark(3);
    void* const* tab = (void* const*)P.d_ptrs;
    GLOC_CUDA_TRY(launch_knn_p2p_gather_merge(tab + 128, tab + 64 * par, comm->size, comm->rank, epoch, idx_off, d2_off,
                                              replicated ? 0 : (size_t)comm->rank * nq, (int)nq, (int)k, d_out_idx,
                                              d_out_d2, P.d_err, stream));
    ix->stats.kernel_launches++;
    mark(4);
    if (timing) {
      cudaEventSynchronize(tev[4]);
      float t[4] = {0, 0, 0, 0};
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
      fprintf(stderr, "[shard] ranks=%d nq=%zu %s (peer memory): gather queries=%.3f ms  local search (%zu queries x %zu rows)"
                      "=%.3f ms  signal + wait + gather + merge=%.3f ms\n", comm->size, nq,
              replicated ? "replicated" : "sliced", t[0], n_search, std::min(ix->n, ix->search_limit), t[1], t[3]);
      for (auto& e : tev) cudaEventDestroy(e);
    }
    return GLOC_OK;
  }
  if (replicated) {
    GLOC_CUDA_TRY(ix->sh_lidx.reserve(nq * k * sizeof(uint64_t)));
    GLOC_CUDA_TRY(ix->sh_ld2.reserve(nq * k * sizeof(float)));
    GLOC_CUDA_TRY(ix->sh_gidx.reserve(N * nq * k * sizeof(uint64_t)));
    GLOC_CUDA_TRY(ix->sh_gd2.reserve(N * nq * k * sizeof(float)));
    mark(1);
    rc = gloc_knn_query_device(ix, d_q, nq, k, (uint64_t*)ix->sh_lidx.p, (float*)ix->sh_ld2.p, stream);
    if (rc != GLOC_OK) return rc;
    mark(2);
    rc = comm_all_gather(comm, ix->sh_lidx.p, ix->sh_gidx.p, nq * k * sizeof(uint64_t), stream);
    if (rc != GLOC_OK) return rc;
    rc = comm_all_gather(comm, ix->sh_ld2.p, ix->sh_gd2.p, nq * k * sizeof(float), stream);
    if (rc != GLOC_OK) return rc;
  } else {
    const size_t nq_all = N * nq;
    GLOC_CUDA_TRY(ix->sh_allq.reserve(nq_all * dim * sizeof(float)));
    GLOC_CUDA_TRY(ix->sh_lidx.reserve(nq_all * k * sizeof(uint64_t)));
    GLOC_CUDA_TRY(ix->sh_ld2.reserve(nq_all * k * sizeof(float)));
    GLOC_CUDA_TRY(ix->sh_gidx.reserve(nq_all * k * sizeof(uint64_t)));
    GLOC_CUDA_TRY(ix->sh_gd2.reserve(nq_all * k * sizeof(float)));
    rc = comm_all_gather(comm, d_q, ix->sh_allq.p, nq * dim * sizeof(float), stream);
    if (rc != GLOC_OK) return rc;
    mark(1);
    rc = gloc_knn_query_device(ix, (const float*)ix->sh_allq.p, nq_all, k, (uint64_t*)ix->sh_lidx.p,
                               (float*)ix->sh_ld2.p, stream);
    if (rc != GLOC_OK) return rc;
    mark(2);
    // block r of my lists = the queries rank r owns; I receive my queries' lists from every shard,
    // laid out [shard][nq][k]: exactly what the merge takes
    rc = comm_all_to_all(comm, ix->sh_lidx.p, ix->sh_gidx.p, nq * k * sizeof(uint64_t), stream);
    if (rc != GLOC_OK) return rc;
    rc = comm_all_to_all(comm, ix->sh_ld2.p, ix->sh_gd2.p, nq * k * sizeof(float), stream);
    if (rc != GLOC_OK) return rc;
  }
  mark(3);
  GLOC_CUDA_TRY(launch_knn_merge_pairs((const uint64_t*)ix->sh_gidx.p, (const float*)ix->sh_gd2.p, (int)N,
                                       (int)nq, (int)k, d_out_idx, d_out_d2, stream));
  ix->stats.kernel_launches++;
  mark(4);
  if (timing) {
    cudaEventSynchronize(tev[4]);
    float t[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
    fprintf(stderr, "[shard] ranks=%d nq=%zu %s: gather queries=%.3f ms  local search (%zu queries x %zu rows)=%.3f ms  "
                    "exchange lists=%.3f ms  merge=%.3f ms\n", comm->size, nq, replicated ? "replicated" : "sliced",
            t[0], replicated ? nq : N * nq, std::min(ix->n, ix->search_limit), t[1], t[2], t[3]);
    for (auto& e : tev) cudaEventDestroy(e);
  }
  return GLOC_OK;
}

// Same with HOST buffers: upload of this rank's queries, download of this rank's results.
int gloc_knn_query_sharded(gloc_knn_index* ix, gloc_comm* comm, const float* q, size_t nq, size_t k,
                           uint64_t* out_idx, float* out_d2, int replicated) {
  if (!ix || !comm) return fail(GLOC_ERR_INVALID, "gloc_knn_query_sharded: null argument");
  if (nq == 0) return GLOC_OK;
  if (!q || !out_idx || !out_d2) return fail(GLOC_ERR_INVALID, "gloc_knn_query_sharded: null buffer");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_knn_query_sharded: cudaSetDevice failed");
  GLOC_CUDA_TRY(ix->stage_q.reserve(nq * ix->dim * sizeof(float)));
  GLOC_CUDA_TRY(ix->stage_idx.reserve(nq * k * sizeof(uint64_t)));
  GLOC_CUDA_TRY(ix->stage_d2.reserve(nq * k * sizeof(float)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(ix->stage_q.p, q, nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
  int rc = gloc_knn_query_sharded_device(ix, comm, (const float*)ix->stage_q.p, nq, k, (uint64_t*)ix->stage_idx.p,
                                         (float*)ix->stage_d2.p, replicated, ix->stream);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_idx, ix->stage_idx.p, nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, ix->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_d2, ix->stage_d2.p, nq * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(ix->stream));
  return GLOC_OK;
}

}  // extern "C"

// internal accessors for the localizer (loc_api.cu)
namespace gloc {
int knn_device_of(const gloc_knn_index* ix) { return ix->device; }
uint64_t knn_offset_of(const gloc_knn_index* ix) { return ix->offset; }
size_t knn_searchable_rows(const gloc_knn_index* ix) { return std::min(ix->n, ix->search_limit); }
}  // namespace gloc
