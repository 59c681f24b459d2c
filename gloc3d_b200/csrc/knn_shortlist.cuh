// knn_shortlist.cuh -- host interface of the tensor-core shortlist path
// (K1 norms + K2 tcgen05 GEMM shortlist + K3 FP32 re-rank), defined in knn_shortlist.cu.
#pragma once
#include "common.cuh"

namespace gloc {

struct ShortlistState;  // bf16 copy of the DB, row norms, TMA descriptors, workspaces

struct ShortlistArgs {
  int device = 0;
  const float* d_db = nullptr;  // fp32 rows (device)
  size_t n_rows = 0;            // searchable rows
  size_t n_total = 0;           // rows held by the index (>= n_rows)
  size_t dim = 0;
  const float* d_q = nullptr;
  size_t nq = 0, k = 0;
  uint64_t offset = 0;
  uint64_t* d_idx = nullptr;
  float* d_d2 = nullptr;
  cudaStream_t stream = nullptr;
  EventProfiler* prof = nullptr;  // brackets the GEMM shortlist kernel when enabled
};

// Can the shortlist path run at all for this shape?
bool shortlist_supported(size_t dim, size_t k);
// Is it the faster choice (GLOC_KNN_AUTO)?
bool shortlist_applicable(size_t dim, size_t n_rows, size_t nq, size_t k);
// Derived data of rows >= first_dirty_row must be rebuilt (DB replaced / appended).
void shortlist_invalidate(ShortlistState* s, size_t first_dirty_row);
void shortlist_destroy(ShortlistState* s);
// Answers all nq queries (overflowed shortlists are re-run through the exact scan on the
// device).  Adds to *launches / *fallback / *rows_reranked when the counts are known
// without a device sync (fallback may be left untouched).
int shortlist_query(ShortlistState** s, const ShortlistArgs& a, uint64_t* launches,
                    uint64_t* fallback, uint64_t* rows_reranked);

// Cumulative counters kept on the device (synchronises): rows re-ranked in FP32 and queries
// whose shortlist overflowed (re-run by the exact scan).
int shortlist_counters(ShortlistState* s, uint64_t* rows_reranked, uint64_t* overflowed);

// Resident CTA pairs the experimental cta_group::2 GEMM would run with on the current device;
// 0 unless GLOC_KNN_PAIR=1 is set and the variant can be launched.
int shortlist_pair_workers();

}  // namespace gloc
