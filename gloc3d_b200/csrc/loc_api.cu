// loc_api.cu -- the whole query path in one call: top-k retrieval -> the k candidates' map grids
// gathered on the device -> scan-match verification of every (query, candidate) pair -> the located
// frame and pose per query.  Replaces the evaluation loop of the reference driver:
//   GlocEvaluator::detect_all_query    global_localization.cpp:482-509  (RpyPCLoopDetector::detect,
//                                      loop_detector.cpp:22-46 -> InvKeyTree::query)
//   GlocEvaluator::global_registraion  global_localization.cpp:511-574  (candidates in retrieval
//                                      order, loop_detector_.match(...) :519-524, first match wins)
// with FastCorrelativeScanMatcher2D::MatchWithSearchParameters (2d/fast_..._2d.cpp:270-320) as the
// verifier.  Nothing of a batch goes through the host between the stages.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "comm.cuh"
#include "csm_store.cuh"

namespace gloc {
int knn_device_of(const gloc_knn_index* ix);
uint64_t knn_offset_of(const gloc_knn_index* ix);
size_t knn_searchable_rows(const gloc_knn_index* ix);
}  // namespace gloc

using namespace gloc;

namespace {

struct QueryScan {     // per query: its scan and initial pose (host libm quaternion, as match_batch)
  long long pt_begin;
  int n_pts;
  float w0, z0, tx, ty;
};

// pair (q, c) = candidate c of query q: the map grid of the retrieved row, the query's scan
__global__ void loc_make_pairs_kernel(const uint64_t* __restrict__ idx, int nq, int k, uint64_t offset,
                                      const int* __restrict__ grid_of_row, const QueryScan* __restrict__ qs,
                                      CsmPairDev* __restrict__ pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * k) return;
  const QueryScan s = qs[i / k];
  const uint64_t row = idx[i] - offset;
  CsmPairDev p;
  p.grid = 0;
  p.gid = grid_of_row ? grid_of_row[row] : (int)row;
  p.pt_begin = s.pt_begin;
  p.n_pts = s.n_pts;
  p.w0 = s.w0; p.z0 = s.z0; p.tx = s.tx; p.ty = s.ty;
  pairs[i] = p;
}

// Which rank verifies which pair of a wave (the balanced form of "the owner verifies"): pairs in
// (query, candidate) order, owner[i] = the rank that holds pair i's row (or -1: no such row).  A rank
// keeps the first quota = ceil(pairs / n_ranks) of the pairs it owns; the rest go, in order, to the
// ranks with room, lowest rank first.  Pure function of its arguments: every rank derives the same
// assignment from the same retrieval results without communication.
void loc_assign_pairs(const int32_t* owner, size_t n, int n_ranks, int32_t* verifier) {
  std::vector<size_t> owned((size_t)n_ranks, 0), kept((size_t)n_ranks, 0), room((size_t)n_ranks, 0);
  size_t total = 0;
  for (size_t i = 0; i < n; ++i)
    if (owner[i] >= 0 && owner[i] < n_ranks) {
      owned[(size_t)owner[i]]++;
      total++;
    }
  const size_t quota = (total + (size_t)n_ranks - 1) / (size_t)n_ranks;
  for (int t = 0; t < n_ranks; ++t) room[(size_t)t] = owned[(size_t)t] < quota ? quota - owned[(size_t)t] : 0;
  int next = 0;                                 // next rank with room
  for (size_t i = 0; i < n; ++i) {
    const int r = owner[i];
    if (r < 0 || r >= n_ranks) {
      verifier[i] = -1;
      continue;
    }
    int by = r;
    if (kept[(size_t)r] < quota) {
      kept[(size_t)r]++;
    } else {
      while (next < n_ranks && room[(size_t)next] == 0) ++next;
      if (next < n_ranks) {
        by = next;
        room[(size_t)next]--;
      }
    }
    verifier[i] = by;
  }
}

}  // namespace

struct gloc_localizer {
  gloc_knn_index* knn = nullptr;
  gloc_csm_store* csm = nullptr;
  int device = 0;
  std::vector<int32_t> h_map;   // row -> grid (empty: identity, the reference's db_grids_[db_idx])
  CsmBuf d_map, d_q, d_idx, d_d2, d_qs, d_pairs, d_pts, d_keys;
  gloc_loc_stats stats{};
  EventProfiler prof_total, prof_retrieval;   // device-side spans on the store's stream
  // what gloc_loc_share_grids learnt about the job (all ranks hold the same tables)
  struct Shared {
    bool on = false;
    int size = 0, rank = 0;
    size_t local_grids = 0, local_rows = 0;          // this rank's store / shard when the tables were built
    std::vector<uint64_t> row_lo, row_n;             // [size] global row range of every rank
    std::vector<size_t> n_grids;                     // [size]
    std::vector<size_t> foreign_base;                // [size] rank r's grids in the store's foreign table
    std::vector<std::vector<int32_t>> maps;          // [size] row -> grid of every rank (empty: identity)
    std::vector<void*> chunks;                       // local buffers mapped through the communicator
    std::vector<void*> dummies;                      // stand-ins where a peer has more arena chunks
  } shared;
};

namespace {

int loc_run(gloc_localizer* L, const float* d_queries, size_t nq, const float* d_pts,
            const int64_t* scan_offsets, const double* init_xyyaw, const gloc_loc_params* P,
            uint64_t* out_idx, float* out_d2, gloc_csm_result* cand_results, gloc_loc_result* results) {
  gloc_csm_store* st = L->csm;
  const int k = P->k;
  if (k < 1 || k > 128) return fail(GLOC_ERR_RANGE, "gloc_loc_localize: k must be in [1, 128]");
  if (P->depth < 1 || P->depth > kCsmMaxDepth)
    return fail(GLOC_ERR_RANGE, "gloc_loc_localize: depth must be in [1, 8]");
  if (P->n_lin < 0 || P->n_ang < 0) return fail(GLOC_ERR_INVALID, "gloc_loc_localize: negative window");
  const long long S = 2ll * P->n_ang + 1, W = 2ll * P->n_lin + 1;
  if (S * W * W >= (1ll << 32) || S > 65535)
    return fail(GLOC_ERR_RANGE, "gloc_loc_localize: search window too large (scans*(2*n_lin+1)^2 must be < 2^32)");
  if (P->policy != GLOC_LOC_VERIFY_ALL && P->policy != GLOC_LOC_FIRST_MATCH)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize: unknown policy");
  if (nq * (size_t)k > (size_t)INT32_MAX) return fail(GLOC_ERR_RANGE, "gloc_loc_localize: too many pairs in one call");
  // RpyPCLoopDetector::detect leaves its outputs untouched when the database is too small
  // (loop_detector.cpp:27-30); here a database with fewer than k searchable rows is an error
  const size_t rows = knn_searchable_rows(L->knn);
  if (rows < (size_t)k)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_loc_localize: fewer than k searchable rows in the database");
  const size_t n_grids = st->recs.size();
  if (n_grids == 0) return fail(GLOC_ERR_NOT_BUILT, "gloc_loc_localize: the grid store is empty");
  if (L->h_map.empty() ? gloc_knn_size(L->knn) > n_grids : gloc_knn_size(L->knn) > L->h_map.size())
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize: database rows without a map grid (add the grids or set the row -> grid table)");
  const int64_t total_pts = scan_offsets[nq];
  std::vector<QueryScan> hq(nq);
  for (size_t q = 0; q < nq; ++q) {
    const int64_t b = scan_offsets[q], e = scan_offsets[q + 1];
    if (b < 0 || e <= b || e > total_pts || e - b > INT32_MAX)
      return fail(GLOC_ERR_INVALID, "gloc_loc_localize: bad scan offsets (every query needs a non-empty scan)");
    const double* in = init_xyyaw ? init_xyyaw + 3 * q : nullptr;
    const float ha = 0.5f * (float)(in ? in[2] : 0.0);
    hq[q].pt_begin = b;
    hq[q].n_pts = (int)(e - b);
    hq[q].w0 = std::cos(ha);
    hq[q].z0 = std::sin(ha);
    hq[q].tx = (float)(in ? in[0] : 0.0);
    hq[q].ty = (float)(in ? in[1] : 0.0);
  }
  const CsmParams prm = csm_make_params(P->n_lin, P->n_ang, P->depth, P->min_score);
  std::vector<float2> rot;
  csm_host_rotations(P->n_ang, P->ang_step, &rot);
  // the candidates can be any grid of the store: plan for the largest
  CsmBatchPlan plan;
  csm_make_plan(st->max_nx, st->max_ny, st->n_graded == 0, P->n_lin, P->depth, &plan);

  cudaStream_t stream = st->stream;
  const size_t n_pairs = nq * (size_t)k;
  GLOC_CUDA_TRY(L->d_idx.reserve(n_pairs * sizeof(uint64_t)));
  GLOC_CUDA_TRY(L->d_d2.reserve(n_pairs * sizeof(float)));
  GLOC_CUDA_TRY(L->d_qs.reserve(nq * sizeof(QueryScan)));
  GLOC_CUDA_TRY(L->d_pairs.reserve(n_pairs * sizeof(CsmPairDev)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2), cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_qs.p, hq.data(), nq * sizeof(QueryScan), cudaMemcpyHostToDevice, stream));
  // ---- stage 1 on the store's stream: the pairs are made from its output without leaving the device
  L->prof_retrieval.begin(stream);
  int rc = gloc_knn_query_device(L->knn, d_queries, nq, (size_t)k, (uint64_t*)L->d_idx.p, (float*)L->d_d2.p, stream);
  L->prof_retrieval.end(stream);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_idx, L->d_idx.p, n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_d2, L->d_d2.p, n_pairs * sizeof(float), cudaMemcpyDeviceToHost, stream));
  const int* d_map = L->h_map.empty() ? nullptr : (const int*)L->d_map.p;
  const uint64_t offset = knn_offset_of(L->knn);
  std::vector<unsigned long long> hbest(n_pairs, 0ull);
  std::vector<unsigned char> verified(n_pairs, 0);

  if (P->policy == GLOC_LOC_VERIFY_ALL) {
    loc_make_pairs_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream>>>(
        (const uint64_t*)L->d_idx.p, (int)nq, k, offset, d_map, (const QueryScan*)L->d_qs.p,
        (CsmPairDev*)L->d_pairs.p);
    GLOC_CUDA_TRY(cudaGetLastError());
    L->stats.kernel_launches++;
    rc = csm_match_core(st, plan, d_pts, (CsmPairDev*)L->d_pairs.p, (int)n_pairs, prm,
                        (const float2*)st->rot.p, hbest.data());
    if (rc != GLOC_OK) return rc;
    std::fill(verified.begin(), verified.end(), 1);
    L->stats.pairs_verified += n_pairs;
  } else {
    // the reference's order of evaluation: candidate c is tried only when candidates 0..c-1 failed
    // (global_localization.cpp:519-524).  Waves of candidates [c0, c1) over the queries still
    // unlocated give the same located frame and pose with one device batch per wave.
    GLOC_CUDA_TRY(cudaStreamSynchronize(stream));   // out_idx is on the host now
    std::vector<char> done(nq, 0);
    std::vector<CsmPairDev> hp;
    std::vector<size_t> where;
    for (int c0 = 0; c0 < k;) {
      const int c1 = std::min(k, c0 == 0 ? 1 : 2 * c0);
      hp.clear();
      where.clear();
      for (size_t q = 0; q < nq; ++q) {
        if (done[q]) continue;
        for (int c = c0; c < c1; ++c) {
          const uint64_t row = out_idx[q * k + c] - offset;
          CsmPairDev p;
          p.grid = 0;
          p.gid = L->h_map.empty() ? (int)row : L->h_map[row];
          p.pt_begin = hq[q].pt_begin;
          p.n_pts = hq[q].n_pts;
          p.w0 = hq[q].w0; p.z0 = hq[q].z0; p.tx = hq[q].tx; p.ty = hq[q].ty;
          hp.push_back(p);
          where.push_back(q * k + c);
        }
      }
      if (hp.empty()) break;
      GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_pairs.p, hp.data(), hp.size() * sizeof(CsmPairDev),
                                    cudaMemcpyHostToDevice, stream));
      std::vector<unsigned long long> wb(hp.size());
      rc = csm_match_core(st, plan, d_pts, (CsmPairDev*)L->d_pairs.p, (int)hp.size(), prm,
                          (const float2*)st->rot.p, wb.data());
      if (rc != GLOC_OK) return rc;
      L->stats.pairs_verified += hp.size();
      L->stats.waves++;
      for (size_t i = 0; i < hp.size(); ++i) {
        hbest[where[i]] = wb[i];
        verified[where[i]] = 1;
        uint32_t sb = (uint32_t)(wb[i] >> 32);
        float sc;
        std::memcpy(&sc, &sb, 4);
        if (wb[i] != 0 && sc > P->min_score) done[where[i] / k] = 1;
      }
      c0 = c1;
    }
  }
  L->prof_total.end(stream);
  GLOC_CUDA_TRY(cudaStreamSynchronize(stream));

  // ---- decode: Candidate2D + pose per verified pair, then global_registraion's choice per query
  static const double zero3[3] = {0, 0, 0};
  for (size_t q = 0; q < nq; ++q) {
    gloc_loc_result& R = results[q];
    std::memset(&R, 0, sizeof(R));
    R.candidate = -1;
    R.best_candidate = -1;
    R.db_index = UINT64_MAX;
    const double* in = init_xyyaw ? init_xyyaw + 3 * q : zero3;
    float best_score = 0.f;
    for (int c = 0; c < k; ++c) {
      const size_t i = q * k + c;
      gloc_csm_result r;
      std::memset(&r, 0, sizeof(r));
      r.score = P->min_score;
      if (verified[i]) {
        const uint64_t row = out_idx[i] - offset;
        const int gid = L->h_map.empty() ? (int)row : L->h_map[row];
        csm_decode(hbest[i], prm, P->ang_step, st->recs[gid].resolution, in, P->min_score, &r);
        R.n_verified++;
      } else {
        r.reserved = -1;   // not evaluated (GLOC_LOC_FIRST_MATCH: an earlier candidate matched)
      }
      if (cand_results) cand_results[i] = r;
      if (r.found) {
        if (R.candidate < 0) {
          R.located = 1;
          R.candidate = c;
          R.db_index = out_idx[i];
          R.match = r;
        }
        if (R.best_candidate < 0 || r.score > best_score) {
          R.best_candidate = c;
          best_score = r.score;
        }
      }
    }
  }
  L->stats.queries += nq;
  return GLOC_OK;
}

// ---- the same path over a row-sharded database: every rank holds the descriptors
// of rows [offset, offset + n) AND their map grids.  Collective call, same arguments on every rank:
//   1. every rank searches the whole batch on its shard; the local top-k lists are all-gathered
//      and merged on every rank (gloc_knn_query_sharded_device, replicated queries);
//   2. a (query, candidate) pair is verified by the rank that owns the candidate's row -- its grid
//      lives there; the query's scan is already everywhere;
//   3. one all-reduce (max over ranks of the 64-bit result keys, 0 where a rank does not own the
//      pair) gives every rank every pair's result; decode and the per-query choice are replicated.
// GLOC_LOC_FIRST_MATCH runs the same three steps per wave of candidates.
int loc_run_sharded(gloc_localizer* L, gloc_comm* comm, const float* d_queries, size_t nq, const float* d_pts,
                    const int64_t* scan_offsets, const double* init_xyyaw, const gloc_loc_params* P,
                    uint64_t* out_idx, float* out_d2, gloc_csm_result* cand_results, gloc_loc_result* results) {
  gloc_csm_store* st = L->csm;
  const int k = P->k;
  if (k < 1 || k > 128) return fail(GLOC_ERR_RANGE, "gloc_loc_localize_sharded: k must be in [1, 128]");
  if (P->depth < 1 || P->depth > kCsmMaxDepth) return fail(GLOC_ERR_RANGE, "gloc_loc_localize_sharded: depth must be in [1, 8]");
  if (P->n_lin < 0 || P->n_ang < 0) return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: negative window");
  const long long S = 2ll * P->n_ang + 1, W = 2ll * P->n_lin + 1;
  if (S * W * W >= (1ll << 32) || S > 65535) return fail(GLOC_ERR_RANGE, "gloc_loc_localize_sharded: search window too large");
  if (P->policy != GLOC_LOC_VERIFY_ALL && P->policy != GLOC_LOC_FIRST_MATCH)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: unknown policy");
  if (nq * (size_t)k > (size_t)INT32_MAX) return fail(GLOC_ERR_RANGE, "gloc_loc_localize_sharded: too many pairs in one call");
  const size_t n_local = gloc_knn_size(L->knn), n_grids = st->recs.size();
  if (n_grids == 0) return fail(GLOC_ERR_NOT_BUILT, "gloc_loc_localize_sharded: the grid store of this shard is empty");
  if (L->h_map.empty() ? n_local > n_grids : n_local > L->h_map.size())
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: shard rows without a map grid");
  // balanced verification needs the tables of gloc_loc_share_grids, still describing this shard
  // (checked before the first collective of the call)
  const bool balanced = L->shared.on && L->shared.size == comm->size && L->shared.local_grids == n_grids &&
                        L->shared.local_rows == n_local;
  if (L->shared.on && !balanced)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: the shard changed since gloc_loc_share_grids (share again)");
  const double resolution = st->recs[0].resolution;   // one resolution per map (the decode is replicated)
  const int64_t total_pts = scan_offsets[nq];
  std::vector<QueryScan> hq(nq);
  for (size_t q = 0; q < nq; ++q) {
    const int64_t b = scan_offsets[q], e = scan_offsets[q + 1];
    if (b < 0 || e <= b || e > total_pts || e - b > INT32_MAX)
      return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: bad scan offsets (every query needs a non-empty scan)");
    const double* in = init_xyyaw ? init_xyyaw + 3 * q : nullptr;
    const float ha = 0.5f * (float)(in ? in[2] : 0.0);
    hq[q].pt_begin = b;
    hq[q].n_pts = (int)(e - b);
    hq[q].w0 = std::cos(ha);
    hq[q].z0 = std::sin(ha);
    hq[q].tx = (float)(in ? in[0] : 0.0);
    hq[q].ty = (float)(in ? in[1] : 0.0);
  }
  const CsmParams prm = csm_make_params(P->n_lin, P->n_ang, P->depth, P->min_score);
  std::vector<float2> rot;
  csm_host_rotations(P->n_ang, P->ang_step, &rot);
  CsmBatchPlan plan;
  // (with shared grids a rank may meet any rank's grid)
  csm_make_plan(std::max(st->max_nx, L->shared.on ? st->f_max_nx : 0), std::max(st->max_ny, L->shared.on ? st->f_max_ny : 0),
                st->n_graded + (L->shared.on ? st->f_graded : 0) == 0, P->n_lin, P->depth, &plan);
  cudaStream_t stream = st->stream;
  const size_t n_pairs = nq * (size_t)k;
  GLOC_CUDA_TRY(L->d_idx.reserve(n_pairs * sizeof(uint64_t)));
  GLOC_CUDA_TRY(L->d_d2.reserve(n_pairs * sizeof(float)));
  GLOC_CUDA_TRY(L->d_pairs.reserve(n_pairs * sizeof(CsmPairDev)));
  GLOC_CUDA_TRY(L->d_keys.reserve(2 * n_pairs * sizeof(uint64_t)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2), cudaMemcpyHostToDevice, stream));
  L->prof_retrieval.begin(stream);
  int rc = gloc_knn_query_sharded_device(L->knn, comm, d_queries, nq, (size_t)k, (uint64_t*)L->d_idx.p,
                                         (float*)L->d_d2.p, 1, stream);
  L->prof_retrieval.end(stream);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_idx, L->d_idx.p, n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_d2, L->d_d2.p, n_pairs * sizeof(float), cudaMemcpyDeviceToHost, stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(stream));
  const uint64_t lo = knn_offset_of(L->knn), hi = lo + n_local;
  std::vector<unsigned long long> hkeys(n_pairs, 0ull), hall(n_pairs, 0ull);
  std::vector<unsigned char> verified(n_pairs, 0);
  std::vector<char> done(nq, 0);
  std::vector<CsmPairDev> hp;
  std::vector<size_t> where, wave;
  std::vector<int32_t> owner_of, verifier;
  std::vector<unsigned long long> wb;
  unsigned long long* d_send = (unsigned long long*)L->d_keys.p;
  unsigned long long* d_recv = d_send + n_pairs;
  for (int c0 = 0; c0 < k;) {
    const int c1 = P->policy == GLOC_LOC_VERIFY_ALL ? k : std::min(k, c0 == 0 ? 1 : 2 * c0);
    hp.clear();
    where.clear();
    auto add_pair = [&](size_t q, size_t i, int gid) {
      CsmPairDev p;
      p.grid = 0;
      p.gid = gid;
      p.pt_begin = hq[q].pt_begin;
      p.n_pts = hq[q].n_pts;
      p.w0 = hq[q].w0; p.z0 = hq[q].z0; p.tx = hq[q].tx; p.ty = hq[q].ty;
      hp.push_back(p);
      where.push_back(i);
    };
    if (!balanced) {
      for (size_t q = 0; q < nq; ++q) {
        if (done[q]) continue;
        for (int c = c0; c < c1; ++c) {
          const size_t i = q * k + c;
          verified[i] = 1;                          // by its owner, somewhere in the job
          const uint64_t gi = out_idx[i];
          if (gi < lo || gi >= hi) continue;        // another rank's row (or an empty slot)
          add_pair(q, i, L->h_map.empty() ? (int)(gi - lo) : L->h_map[gi - lo]);
        }
      }
    } else {
      // Every rank holds the same retrieval results, so every rank derives the same assignment without
      // talking: the pairs of the wave in (q, c) order, each owned by the rank that holds its row; a rank
      // keeps the first `quota` of its own, the rest go, in order, to the ranks with room (lowest first).
      const auto& SH = L->shared;
      const int N = SH.size;
      owner_of.clear();
      wave.clear();
      for (size_t q = 0; q < nq; ++q) {
        if (done[q]) continue;
        for (int c = c0; c < c1; ++c) {
          const size_t i = q * k + c;
          verified[i] = 1;
          const uint64_t gi = out_idx[i];
          int r = -1;
          for (int t = 0; t < N; ++t)
            if (gi >= SH.row_lo[t] && gi - SH.row_lo[t] < SH.row_n[t]) { r = t; break; }
          if (r < 0) continue;                      // an empty slot (fewer than k rows in the job)
          wave.push_back(i);
          owner_of.push_back(r);
        }
      }
      verifier.resize(wave.size());
      loc_assign_pairs(owner_of.data(), wave.size(), N, verifier.data());
      for (size_t w = 0; w < wave.size(); ++w) {
        const int r = owner_of[w];
        if (verifier[w] != SH.rank) continue;
        const size_t i = wave[w], q = i / (size_t)k;
        const uint64_t row = out_idx[i] - SH.row_lo[(size_t)r];
        const int g_local = SH.maps[(size_t)r].empty() ? (int)row : SH.maps[(size_t)r][row];
        add_pair(q, i, r == SH.rank ? g_local : (int)(SH.local_grids + SH.foreign_base[(size_t)r] + (size_t)g_local));
        if (r != SH.rank) L->stats.pairs_migrated++;
      }
    }
    if (!hp.empty()) {
      GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_pairs.p, hp.data(), hp.size() * sizeof(CsmPairDev), cudaMemcpyHostToDevice, stream));
      wb.resize(hp.size());
      rc = csm_match_core(st, plan, d_pts, (CsmPairDev*)L->d_pairs.p, (int)hp.size(), prm, (const float2*)st->rot.p, wb.data());
      if (rc != GLOC_OK) return rc;
      for (size_t i = 0; i < hp.size(); ++i) hkeys[where[i]] = wb[i];
      L->stats.pairs_verified += hp.size();
    }
    L->stats.waves++;
    // every pair has exactly one owner: max over ranks of (key | 0) is the owner's key
    GLOC_CUDA_TRY(cudaMemcpyAsync(d_send, hkeys.data(), n_pairs * 8, cudaMemcpyHostToDevice, stream));
    rc = comm_all_reduce_max_u64(comm, d_send, d_recv, n_pairs, stream);
    if (rc != GLOC_OK) return rc;
    GLOC_CUDA_TRY(cudaMemcpyAsync(hall.data(), d_recv, n_pairs * 8, cudaMemcpyDeviceToHost, stream));
    GLOC_CUDA_TRY(cudaStreamSynchronize(stream));
    for (size_t q = 0; q < nq; ++q)
      for (int c = c0; c < c1 && !done[q]; ++c) {
        const unsigned long long key = hall[q * k + c];
        uint32_t sb = (uint32_t)(key >> 32);
        float sc;
        std::memcpy(&sc, &sb, 4);
        if (key != 0 && sc > P->min_score) done[q] = 1;
      }
    c0 = c1;
  }
  L->prof_total.end(stream);
  static const double zero3[3] = {0, 0, 0};
  for (size_t q = 0; q < nq; ++q) {
    gloc_loc_result& R = results[q];
    std::memset(&R, 0, sizeof(R));
    R.candidate = -1;
    R.best_candidate = -1;
    R.db_index = UINT64_MAX;
    const double* in = init_xyyaw ? init_xyyaw + 3 * q : zero3;
    float best_score = 0.f;
    for (int c = 0; c < k; ++c) {
      const size_t i = q * k + c;
      gloc_csm_result r;
      std::memset(&r, 0, sizeof(r));
      r.score = P->min_score;
      if (verified[i]) {
        csm_decode(hall[i], prm, P->ang_step, resolution, in, P->min_score, &r);
        R.n_verified++;
      } else {
        r.reserved = -1;
      }
      if (cand_results) cand_results[i] = r;
      if (r.found) {
        if (R.candidate < 0) {
          R.located = 1;
          R.candidate = c;
          R.db_index = out_idx[i];
          R.match = r;
        }
        if (R.best_candidate < 0 || r.score > best_score) {
          R.best_candidate = c;
          best_score = r.score;
        }
      }
    }
  }
  L->stats.queries += nq;
  return GLOC_OK;
}

}  // namespace

extern "C" {

int gloc_loc_create(gloc_localizer** out, gloc_knn_index* knn, gloc_csm_store* csm) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_loc_create: out is null");
  *out = nullptr;
  if (!knn || !csm) return fail(GLOC_ERR_INVALID, "gloc_loc_create: null index or store");
  if (knn_device_of(knn) != csm->device)
    return fail(GLOC_ERR_INVALID, "gloc_loc_create: index and grid store live on different devices");
  gloc_localizer* L = new (std::nothrow) gloc_localizer;
  if (!L) return fail(GLOC_ERR_NOMEM, "gloc_loc_create: out of host memory");
  L->knn = knn;
  L->csm = csm;
  L->device = csm->device;
  *out = L;
  return GLOC_OK;
}

void gloc_loc_destroy(gloc_localizer* L) {
  if (!L) return;
  DeviceGuard g(L->device);
  for (CsmBuf* b : {&L->d_map, &L->d_q, &L->d_idx, &L->d_d2, &L->d_qs, &L->d_pairs, &L->d_pts, &L->d_keys}) b->release();
  // (peer mappings of a still shared store belong to the communicator: gloc_loc_unshare_grids or its destruction)
  for (void* p : L->shared.dummies) cudaFree(p);
  delete L;
}

int gloc_loc_set_row_grids(gloc_localizer* L, const int32_t* grid_of_row, size_t n_rows) {
  if (!L) return fail(GLOC_ERR_INVALID, "gloc_loc_set_row_grids: null localizer");
  DeviceGuard g(L->device);
  if (L->shared.on) L->shared.local_rows = (size_t)-1;   // the peers hold the old table: share again
  if (!grid_of_row || n_rows == 0) {   // back to the identity (db_grids_[db_idx], loop_detector.h:36-39)
    L->h_map.clear();
    return GLOC_OK;
  }
  const int n_grids = gloc_csm_num_grids(L->csm);
  for (size_t i = 0; i < n_rows; ++i)
    if (grid_of_row[i] < 0 || grid_of_row[i] >= n_grids)
      return fail(GLOC_ERR_INVALID, "gloc_loc_set_row_grids: grid id out of range at row " + std::to_string(i));
  try {
    L->h_map.assign(grid_of_row, grid_of_row + n_rows);
  } catch (...) {
    return fail(GLOC_ERR_NOMEM, "gloc_loc_set_row_grids: out of host memory");
  }
  GLOC_CUDA_TRY(cudaStreamSynchronize(L->csm->stream));
  GLOC_CUDA_TRY(L->d_map.reserve(n_rows * sizeof(int32_t)));
  GLOC_CUDA_TRY(cudaMemcpy(L->d_map.p, grid_of_row, n_rows * sizeof(int32_t), cudaMemcpyHostToDevice));
  GLOC_CUDA_TRY(cudaStreamSynchronize(0));
  return GLOC_OK;
}

int gloc_loc_localize_device(gloc_localizer* L, const float* d_queries, size_t nq, const float* d_pts,
                             const int64_t* scan_offsets, const double* init_xyyaw,
                             const gloc_loc_params* prm, uint64_t* out_idx, float* out_d2,
                             gloc_csm_result* cand_results, gloc_loc_result* results) {
  if (!L || !prm) return fail(GLOC_ERR_INVALID, "gloc_loc_localize: null argument");
  if (nq == 0) return GLOC_OK;
  if (!d_queries || !d_pts || !scan_offsets || !out_idx || !out_d2 || !results)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize: null buffer");
  DeviceGuard g(L->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_loc_localize: cudaSetDevice failed");
  L->prof_total.begin(L->csm->stream);
  return loc_run(L, d_queries, nq, d_pts, scan_offsets, init_xyyaw, prm, out_idx, out_d2, cand_results, results);
}

int gloc_loc_localize(gloc_localizer* L, const float* queries, size_t nq, const float* pts,
                      const int64_t* scan_offsets, const double* init_xyyaw, const gloc_loc_params* prm,
                      uint64_t* out_idx, float* out_d2, gloc_csm_result* cand_results,
                      gloc_loc_result* results) {
  if (!L || !prm) return fail(GLOC_ERR_INVALID, "gloc_loc_localize: null argument");
  if (nq == 0) return GLOC_OK;
  if (!queries || !pts || !scan_offsets || !out_idx || !out_d2 || !results)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize: null buffer");
  DeviceGuard g(L->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_loc_localize: cudaSetDevice failed");
  const size_t dim = gloc_knn_dim(L->knn);
  const int64_t total_pts = scan_offsets[nq];
  if (total_pts <= 0) return fail(GLOC_ERR_INVALID, "gloc_loc_localize: bad scan offsets");
  cudaStream_t stream = L->csm->stream;
  GLOC_CUDA_TRY(L->d_q.reserve(nq * dim * sizeof(float)));
  GLOC_CUDA_TRY(L->d_pts.reserve((size_t)total_pts * 3 * sizeof(float)));
  L->prof_total.begin(stream);
  GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_q.p, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_pts.p, pts, (size_t)total_pts * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
  return loc_run(L, (const float*)L->d_q.p, nq, (const float*)L->d_pts.p, scan_offsets, init_xyyaw, prm,
                 out_idx, out_d2, cand_results, results);
}

int gloc_loc_localize_sharded(gloc_localizer* L, gloc_comm* comm, const float* queries, size_t nq, const float* pts,
                              const int64_t* scan_offsets, const double* init_xyyaw, const gloc_loc_params* prm,
                              uint64_t* out_idx, float* out_d2, gloc_csm_result* cand_results,
                              gloc_loc_result* results, int buffers_on_device) {
  if (!L || !prm || !comm) return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: null argument");
  if (nq == 0) return GLOC_OK;
  if (!queries || !pts || !scan_offsets || !out_idx || !out_d2 || !results)
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: null buffer");
  if (comm->device != L->device) return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: communicator on another device");
  DeviceGuard g(L->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_loc_localize_sharded: cudaSetDevice failed");
  // every local reason to refuse the call comes BEFORE its first collective (the sliced upload below):
  // a rank that fails alone must not leave its peers waiting inside one
  if (L->shared.on && !(L->shared.size == comm->size && L->shared.local_grids == L->csm->recs.size() &&
                        L->shared.local_rows == gloc_knn_size(L->knn)))
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: the shard changed since gloc_loc_share_grids (share again)");
  if (L->csm->recs.empty()) return fail(GLOC_ERR_NOT_BUILT, "gloc_loc_localize_sharded: the grid store of this shard is empty");
  if (L->h_map.empty() ? gloc_knn_size(L->knn) > L->csm->recs.size() : gloc_knn_size(L->knn) > L->h_map.size())
    return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: shard rows without a map grid");
  cudaStream_t stream = L->csm->stream;
  const float* dq = queries;
  const float* dp = pts;
  L->prof_total.begin(stream);
  if (!buffers_on_device) {
    const size_t dim = gloc_knn_dim(L->knn);
    const int64_t total_pts = scan_offsets[nq];
    if (total_pts <= 0) return fail(GLOC_ERR_INVALID, "gloc_loc_localize_sharded: bad scan offsets");
    GLOC_CUDA_TRY(L->d_q.reserve(nq * dim * sizeof(float)));
    GLOC_CUDA_TRY(L->d_pts.reserve((size_t)total_pts * 3 * sizeof(float)));
    const int N = comm->size;
    if (N > 1 && nq >= (size_t)N && std::getenv("GLOC_LOC_FULL_UPLOAD") == nullptr) {
      // every rank holds the same host batch: each uploads the queries and scans of ITS 1/N of the
      // queries over PCIe, the parts travel to the peers over NVLink (one upload per byte, not N)
      std::vector<size_t> oq((size_t)N + 1), op((size_t)N + 1);
      for (int r = 0; r <= N; ++r) {
        const size_t q = nq * (size_t)r / (size_t)N;
        oq[(size_t)r] = q * dim * sizeof(float);
        op[(size_t)r] = (size_t)scan_offsets[q] * 3 * sizeof(float);
      }
      const int me = comm->rank;
      GLOC_CUDA_TRY(cudaMemcpyAsync((char*)L->d_q.p + oq[(size_t)me], (const char*)queries + oq[(size_t)me],
                                    oq[(size_t)me + 1] - oq[(size_t)me], cudaMemcpyHostToDevice, stream));
      if (op[(size_t)me + 1] > op[(size_t)me])
        GLOC_CUDA_TRY(cudaMemcpyAsync((char*)L->d_pts.p + op[(size_t)me], (const char*)pts + op[(size_t)me],
                                      op[(size_t)me + 1] - op[(size_t)me], cudaMemcpyHostToDevice, stream));
      int rc = comm_all_gather_v(comm, L->d_q.p, oq.data(), stream);
      if (rc == GLOC_OK) rc = comm_all_gather_v(comm, L->d_pts.p, op.data(), stream);
      if (rc != GLOC_OK) return rc;
    } else {
      GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_q.p, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice, stream));
      GLOC_CUDA_TRY(cudaMemcpyAsync(L->d_pts.p, pts, (size_t)total_pts * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    dq = (const float*)L->d_q.p;
    dp = (const float*)L->d_pts.p;
  }
  return loc_run_sharded(L, comm, dq, nq, dp, scan_offsets, init_xyyaw, prm, out_idx, out_d2, cand_results, results);
}

namespace {
struct ShareHeader {
  unsigned long long n_chunks, n_grids, n_rows, row_lo, has_map, n_graded;
  int max_nx, max_ny;
  int ok, pad;   // 0: this rank cannot take part (every rank then fails together, nobody waits)
};
struct ShareGrid {     // a grid record as its owner describes it to its peers
  unsigned long long offset;
  int chunk, nx, ny, enc;
  double resolution, max_x, max_y;
};
}  // namespace

int gloc_loc_assign_pairs(const int32_t* owner, size_t n_pairs, int n_ranks, int32_t* verifier) {
  if (n_pairs && (!owner || !verifier)) return fail(GLOC_ERR_INVALID, "gloc_loc_assign_pairs: null argument");
  if (n_ranks < 1) return fail(GLOC_ERR_INVALID, "gloc_loc_assign_pairs: n_ranks must be positive");
  loc_assign_pairs(owner, n_pairs, n_ranks, verifier);
  return GLOC_OK;
}

int gloc_loc_unshare_grids(gloc_localizer* L, gloc_comm* comm) {
  if (!L || !comm) return fail(GLOC_ERR_INVALID, "gloc_loc_unshare_grids: null argument");
  DeviceGuard g(L->device);
  cudaStreamSynchronize(L->csm->stream);
  for (void* p : L->shared.chunks) comm_unmap_peers(comm, p);
  for (void* p : L->shared.dummies) cudaFree(p);
  L->shared = gloc_localizer::Shared();
  gloc_csm_store* st = L->csm;
  st->foreign.clear();
  st->f_max_nx = st->f_max_ny = 0;
  st->f_graded = 0;
  st->foreign_dirty = true;
  return GLOC_OK;
}

int gloc_loc_share_grids(gloc_localizer* L, gloc_comm* comm) {
  if (!L || !comm) return fail(GLOC_ERR_INVALID, "gloc_loc_share_grids: null argument");
  if (comm->device != L->device) return fail(GLOC_ERR_INVALID, "gloc_loc_share_grids: communicator and localizer live on different devices");
  DeviceGuard g(L->device);
  if (L->shared.on) {
    int rc = gloc_loc_unshare_grids(L, comm);
    if (rc != GLOC_OK) return rc;
  }
  gloc_csm_store* st = L->csm;
  GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
  const int N = comm->size, me = comm->rank;
  ShareHeader mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.n_chunks = st->arena.chunks.size();
  mine.n_grids = st->recs.size();
  mine.n_rows = gloc_knn_size(L->knn);
  mine.row_lo = knn_offset_of(L->knn);
  mine.has_map = L->h_map.empty() ? 0 : 1;
  mine.n_graded = st->n_graded;
  mine.max_nx = st->max_nx;
  mine.max_ny = st->max_ny;
  mine.ok = (mine.has_map ? L->h_map.size() >= mine.n_rows : mine.n_grids >= mine.n_rows) && mine.n_grids > 0;
  std::vector<ShareHeader> hd((size_t)N);
  int rc = comm_host_all_gather(comm, &mine, hd.data(), sizeof(ShareHeader));
  if (rc != GLOC_OK) return rc;
  for (int r = 0; r < N; ++r)
    if (!hd[(size_t)r].ok)
      return fail(GLOC_ERR_INVALID, "gloc_loc_share_grids: rank " + std::to_string(r) +
                                        " has shard rows without a map grid (or an empty grid store)");
  size_t max_chunks = 0, max_grids = 0, max_rows = 0;
  bool any_map = false;
  for (const ShareHeader& h : hd) {
    max_chunks = std::max<size_t>(max_chunks, h.n_chunks);
    max_grids = std::max<size_t>(max_grids, h.n_grids);
    max_rows = std::max<size_t>(max_rows, h.n_rows);
    any_map |= h.has_map != 0;
  }
  auto& SH = L->shared;
  // 1. every arena chunk of every rank, addressable here (ranks with fewer chunks pass stand-ins: the
  //    mapping is a collective per buffer)
  std::vector<std::vector<void*>> chunk_at((size_t)N, std::vector<void*>(max_chunks, nullptr));
  for (size_t i = 0; i < max_chunks; ++i) {
    void* local = nullptr;
    if (i < st->arena.chunks.size()) {
      local = st->arena.chunks[i].p;
    } else {
      GLOC_CUDA_TRY(cudaMalloc(&local, 256));
      SH.dummies.push_back(local);
    }
    void** ptrs = nullptr;
    rc = comm_map_peers(comm, local, &ptrs);
    if (rc != GLOC_OK) {   // collective outcome: every rank fails here together
      for (void* p : SH.chunks) comm_unmap_peers(comm, p);
      for (void* p : SH.dummies) cudaFree(p);
      SH = gloc_localizer::Shared();
      return rc;
    }
    SH.chunks.push_back(local);
    for (int r = 0; r < N; ++r) chunk_at[(size_t)r][i] = ptrs[r];
  }
  // 2. the grid tables
  std::vector<ShareGrid> gs(max_grids), all_g((size_t)N * max_grids);
  std::memset(gs.data(), 0, gs.size() * sizeof(ShareGrid));
  for (size_t j = 0; j < st->recs.size(); ++j) {
    const CsmGridRec& r = st->recs[j];
    int chunk = -1;
    for (size_t c = 0; c < st->arena.chunks.size(); ++c) {
      const unsigned char* b = st->arena.chunks[c].p;
      if ((const unsigned char*)r.data >= b && (const unsigned char*)r.data < b + st->arena.chunks[c].cap) { chunk = (int)c; break; }
    }
    if (chunk < 0) return fail(GLOC_ERR_INVALID, "gloc_loc_share_grids: a grid outside the store's arena");
    gs[j].offset = (unsigned long long)((const unsigned char*)r.data - st->arena.chunks[(size_t)chunk].p);
    gs[j].chunk = chunk; gs[j].nx = r.nx; gs[j].ny = r.ny; gs[j].enc = r.enc;
    gs[j].resolution = r.resolution; gs[j].max_x = r.max_x; gs[j].max_y = r.max_y;
  }
  auto undo = [&](int code) {   // a failed exchange leaves nothing mapped behind
    for (void* p : SH.chunks) comm_unmap_peers(comm, p);
    for (void* p : SH.dummies) cudaFree(p);
    SH = gloc_localizer::Shared();
    return code;
  };
  rc = comm_host_all_gather(comm, gs.data(), all_g.data(), max_grids * sizeof(ShareGrid));
  if (rc != GLOC_OK) return undo(rc);
  // 3. the row -> grid tables (only where some rank has one)
  SH.maps.assign((size_t)N, std::vector<int32_t>());
  if (any_map) {
    std::vector<int32_t> mm(max_rows, 0), all_m((size_t)N * max_rows);
    for (size_t i = 0; i < mine.n_rows; ++i) mm[i] = L->h_map.empty() ? (int32_t)i : L->h_map[i];
    rc = comm_host_all_gather(comm, mm.data(), all_m.data(), max_rows * sizeof(int32_t));
    if (rc != GLOC_OK) return undo(rc);
    for (int r = 0; r < N; ++r)
      SH.maps[(size_t)r].assign(all_m.begin() + (size_t)r * max_rows, all_m.begin() + (size_t)r * max_rows + hd[(size_t)r].n_rows);
  }
  // 4. the peers' grids as records of this store
  st->foreign.clear();
  st->f_max_nx = st->f_max_ny = 0;
  st->f_graded = 0;
  SH.row_lo.resize((size_t)N); SH.row_n.resize((size_t)N); SH.n_grids.resize((size_t)N); SH.foreign_base.assign((size_t)N, 0);
  for (int r = 0; r < N; ++r) {
    SH.row_lo[(size_t)r] = hd[(size_t)r].row_lo;
    SH.row_n[(size_t)r] = hd[(size_t)r].n_rows;
    SH.n_grids[(size_t)r] = hd[(size_t)r].n_grids;
    if (r == me) continue;
    SH.foreign_base[(size_t)r] = st->foreign.size();
    for (size_t j = 0; j < hd[(size_t)r].n_grids; ++j) {
      const ShareGrid& sg = all_g[(size_t)r * max_grids + j];
      CsmGridRec rec;
      rec.data = (const unsigned char*)chunk_at[(size_t)r][(size_t)sg.chunk] + sg.offset;
      rec.nx = sg.nx; rec.ny = sg.ny; rec.enc = sg.enc; rec.pad = 0;
      rec.resolution = sg.resolution; rec.max_x = sg.max_x; rec.max_y = sg.max_y;
      st->foreign.push_back(rec);
    }
    st->f_max_nx = std::max(st->f_max_nx, hd[(size_t)r].max_nx);
    st->f_max_ny = std::max(st->f_max_ny, hd[(size_t)r].max_ny);
    st->f_graded += hd[(size_t)r].n_graded;
  }
  st->foreign_dirty = true;
  SH.size = N;
  SH.rank = me;
  SH.local_grids = st->recs.size();
  SH.local_rows = mine.n_rows;
  SH.on = true;
  return GLOC_OK;
}

int gloc_loc_set_profiling(gloc_localizer* L, int enabled) {
  if (!L) return fail(GLOC_ERR_INVALID, "gloc_loc_set_profiling: null localizer");
  L->prof_total.enabled = L->prof_retrieval.enabled = enabled != 0;
  return GLOC_OK;
}

int gloc_loc_get_profile(gloc_localizer* L, gloc_loc_profile* out) {
  if (!L || !out) return fail(GLOC_ERR_INVALID, "gloc_loc_get_profile: null argument");
  DeviceGuard g(L->device);
  uint64_t n = 0;
  L->prof_total.collect(&out->total_ms, &out->calls);
  L->prof_retrieval.collect(&out->retrieval_ms, &n);
  return GLOC_OK;
}

int gloc_loc_get_stats(const gloc_localizer* L, gloc_loc_stats* out) {
  if (!L || !out) return fail(GLOC_ERR_INVALID, "gloc_loc_get_stats: null argument");
  *out = L->stats;
  return GLOC_OK;
}

}  // extern "C"
