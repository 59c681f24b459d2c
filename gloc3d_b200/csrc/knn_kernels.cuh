// knn_kernels.cuh -- launchers of the stage-1 kernels (definitions in knn_exact.cu,
// knn_shortlist.cu).
#pragma once
#include "common.cuh"

namespace gloc {

// ---- exact scan (knn_exact.cu)
int exact_scan_tile_q(int nq);
int exact_scan_tile_n(int nq);
// partial: [nq][n_ranges][k] packed keys, ascending per (query, range), kEmptyKey padded.
cudaError_t launch_knn_exact_scan(const float* db, long long n_rows, int dim, const float* q,
                                  int nq, int k, int n_ranges, long long rows_per_range,
                                  uint64_t* partial, cudaStream_t stream,
                                  const int* qmap = nullptr, const int* nq_dev = nullptr,
                                  bool force_small = false);
cudaError_t launch_knn_finalize(const uint64_t* partial, int nq, int n_lists, int k,
                                uint64_t idx_offset, uint64_t* out_idx, float* out_d2,
                                cudaStream_t stream, const int* qmap = nullptr,
                                const int* nq_dev = nullptr);
// K4 fused with the exchange: lists pulled from the peers' buffers over NVLink (see knn_exact.cu)
cudaError_t launch_knn_p2p_gather_merge(void* const* d_fbufs, void* const* d_bufs, int n_ranks, int rank, unsigned epoch, size_t idx_off,
                                        size_t d2_off, size_t block, int nq, int k, uint64_t* out_idx,
                                        float* out_d2, int* err, cudaStream_t stream);
cudaError_t launch_knn_merge_pairs(const uint64_t* idx, const float* d2, int g, int nq, int k,
                                   uint64_t* out_idx, float* out_d2, cudaStream_t stream);

// ---- streaming scan for 1..4 queries per call (knn_stream.cu)
bool stream_applicable(size_t dim, size_t nq, size_t k);
int stream_grid(int device, size_t dim);
cudaError_t launch_knn_stream(const float* db, long long n_rows, int dim, const float* q, int nq,
                              int k, int grid, uint64_t* partial, uint64_t idx_offset,
                              uint64_t* out_idx, float* out_d2, int* overflow,
                              EventProfiler* prof, cudaStream_t stream);

}  // namespace gloc
