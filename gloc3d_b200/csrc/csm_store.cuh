// csm_store.cuh -- host-side model of the stage-2 grid store (internal to libgloc3d).
//
// What a store keeps per map grid is ONLY the width-1 precomputation grid, in a pooled device
// arena: bit-packed (one bit per cell, ~49 KB for a KITTI-sized 781 x 504 BEV grid) when every
// cell is 0 or 255 -- the BEV grids of this code base -- else the uint8 cells.  Everything the
// matcher derives from it (PrecomputationGridStack2D, fast_..._2d.cpp:192-215: the coarser
// levels; plus this library's bit planes / phase-major copy) is rebuilt on the device for the
// DISTINCT grids of each batch, into a working set that belongs to the store and is reused from
// batch to batch: a million-frame database costs 49 GB of grids, not terabytes of stacks, and a
// batch never waits on a per-grid allocation or a host round trip.
#pragma once
#include <vector>

#include "csm_kernels.cuh"

namespace gloc {

// Device buffer that only grows (contents are NOT kept across a growth).
struct CsmBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    const size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      e = cudaMalloc(&p, need);
      if (e == cudaSuccess) bytes = need;
    } else {
      bytes = want;
    }
    if (e != cudaSuccess) p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

// Bump allocator over large device chunks: grids are added, never freed individually.
struct CsmArena {
  struct Chunk { unsigned char* p; size_t cap, used; };
  std::vector<Chunk> chunks;
  size_t total = 0;
  cudaError_t alloc(size_t bytes, void** out) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (chunks.empty() || chunks.back().used + bytes > chunks.back().cap) {
      // 64 MB, doubling up to 1 GB: a few dozen cudaMallocs for a million grids
      size_t cap = chunks.empty() ? ((size_t)64 << 20) : std::min(chunks.back().cap * 2, (size_t)1 << 30);
      cap = std::max(cap, bytes);
      unsigned char* p = nullptr;
      cudaError_t e = cudaMalloc((void**)&p, cap);
      if (e != cudaSuccess) return e;
      chunks.push_back(Chunk{p, cap, 0});
    }
    Chunk& c = chunks.back();
    *out = c.p + c.used;
    c.used += bytes;
    total += bytes;
    return cudaSuccess;
  }
  // give back the most recent allocation (a grid that turned out not to be binary)
  void rollback(void* p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (!chunks.empty() && chunks.back().p + chunks.back().used - bytes == p) {
      chunks.back().used -= bytes;
      total -= bytes;
    }
  }
  void release() {
    for (auto& c : chunks) cudaFree(c.p);
    chunks.clear();
    total = 0;
  }
};

// Everything one call of the matcher needs to know about its batch (decided on the host from
// grid dimensions only -- never from device data).
struct CsmBatchPlan {
  CsmPlan dev{};
  bool use_bits = false, bits_paired = false, exp_bits = false, pm_kernel = false;
  size_t bits_smem = 0, exp_smem = 0;
};

}  // namespace gloc

struct gloc_csm_store {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint8_t* d_lut = nullptr;  // uint16 cost value -> uint8 width-1 cell
  gloc::CsmArena arena;
  std::vector<gloc::CsmGridRec> recs;   // host mirror; rec.data points into the arena
  // other ranks' grids this store can read through peer memory (gloc_loc_share_grids): records
  // [recs.size(), recs.size() + foreign.size()) of the device table
  std::vector<gloc::CsmGridRec> foreign;
  bool foreign_dirty = false;
  int f_max_nx = 0, f_max_ny = 0;
  size_t f_graded = 0;
  gloc::CsmBuf d_recs;
  size_t recs_on_device = 0;
  int max_nx = 0, max_ny = 0;
  size_t n_graded = 0;                  // grids kept as uint8 (some cell neither 0 nor 255)
  gloc::CsmBuf pts, pairs, slots, slot_gid, hkeys, hslot, ws, rot, bounds, coarse, top, best,
      survivors, nsurv, nodes, misc, cells16, stage_u8, disc;
  gloc_csm_stats stats{};
  gloc::EventProfiler prof;
};

namespace gloc {

// the device copy of the grid records is current
int csm_sync_recs(gloc_csm_store* st);
// plan for a batch whose grids are bounded by (max_nx, max_ny); all_binary: no graded grid in it
int csm_make_plan(int max_nx, int max_ny, bool all_binary, int n_lin, int depth, CsmBatchPlan* out);
CsmParams csm_make_params(int n_lin, int n_ang, int depth, float min_score);
// Matches pairs [0, n_pairs) already on the device (gid, pt_begin, n_pts, w0, z0, tx, ty filled;
// slots are assigned here); points and per-angle quaternions on the device; keys of the best
// candidates (score bits << 32 | ~rank, 0 = none) to the host.  Enqueues on st->stream and
// synchronises once per sub-batch.
int csm_match_core(gloc_csm_store* st, const CsmBatchPlan& plan, const float* d_pts, CsmPairDev* d_pairs,
                   int n_pairs, const CsmParams& prm, const float2* d_rot, unsigned long long* h_best);
// per-angle quaternions from the host libm (GenerateRotatedScans, correlative_..._2d.cpp:99-107)
void csm_host_rotations(int n_ang, double ang_step, std::vector<float2>* rot);
// Candidate2D + pose of one result key (correlative_..._2d.h:74-87, fast_..._2d.cpp:311-318)
void csm_decode(unsigned long long key, const CsmParams& prm, double ang_step, double resolution,
                const double* init_xyyaw, float min_score, gloc_csm_result* r);

}  // namespace gloc
