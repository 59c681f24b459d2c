// csm_kernels.cuh -- device-side data model + launchers of the stage-2 kernels
// (definitions in csm.cu).
#pragma once
#include "common.cuh"

namespace gloc {

constexpr int kCsmMaxDepth = 8;  // precomputation widths 1..128

// Words (uint32) per row of a bit-packed level of `wide_nx` cells: at least one zero word after
// the last cell (lookups clamp out-of-range columns onto it) and an even count, so that rows are
// 8-byte aligned.
__host__ __device__ inline int csm_bit_stride(int wide_nx) { return (((wide_nx + 31) >> 5) + 2) & ~1; }

// One map grid as a batch sees it (a "slot" of the batch's working set; built on the device by
// csm_prepare_slots_kernel from the store's per-grid records).
//   binary == 0: uint8 precomputation stack, level i = width 2^i, (nx+w-1) x (ny+w-1) cells,
//     stride nx+w-1, offset (-w+1,-w+1) -- the layout of PrecomputationGrid2D::cells_
//     (2d/fast_correlative_scan_matcher_2d.h:113-125) -- plus the padded phase-major copy of
//     the coarsest level for the bulk scorer.
//   binary == 1 (every cell 0 or 255, the BEV grids of this code base): every level
//     bit-packed, bit x of row y of level l at word lvl[l][y * lvs[l] + x / 32], plus the
//     coarsest level as bit planes for the bit-sliced scorer.
struct CsmGridDev {
  int nx, ny;
  double resolution, max_x, max_y;
  int binary;
  const uint8_t* stack;
  long long off[kCsmMaxDepth];
  // Coarsest level re-laid for the bulk scorer: zero border of `pm_pad` cells, then split
  // into w*w phase planes (w = coarsest width = lattice step) of PH x PW bytes each, plane
  // (ry, rx) holding the cells with (y % w, x % w) == (ry, rx).  Lattice neighbours of a
  // point are adjacent bytes, and no lookup of an in-grid point needs a bounds check.
  const uint8_t* pm;
  int pm_pad, pm_pw, pm_ph, pm_log2w;
  const unsigned* lvl[kCsmMaxDepth];
  int lvs[kCsmMaxDepth];
  // The coarsest level as bit planes (csm.cu "bit-sliced").  Plane (ry, rx), row r = one
  // 64-bit word, bit c = cell (w c + rx - pmb_px, w r + ry - pmb_py) of the level's own
  // (wide) frame; pmb_rows rows per plane, zero outside the grid.
  // Paired layout (pmb_b1 >= 0; grids whose data columns fit two overlapping 32-column halves
  // starting at columns pmb_b0 and pmb_b1): plane (ry, rx), half h, row pair p = one 64-bit word,
  // low 32 bits = columns [b_h, b_h + 32) of row 2p, high 32 bits = the same columns of row 2p + 1;
  // word index (2 plane + h) * (pmb_rows / 2) + p.  One load then serves two candidate rows.
  const unsigned long long* pmb;
  int pmb_rows, pmb_px, pmb_py, pmb_log2w;
  int pmb_b0, pmb_b1;
};

// A grid as the store keeps it: the width-1 precomputation grid only (everything else is
// derived per batch).  enc 1: bit-packed rows, csm_bit_stride(nx) words each; enc 0: nx*ny bytes.
struct CsmGridRec {
  const void* data;
  int nx, ny, enc, pad;
  double resolution, max_x, max_y;
};

// How one batch lays out its working set (uniform slots sized for the largest grid).
struct CsmPlan {
  int bits;                 // 1: bit levels + bit planes (all grids binary), 0: uint8 stacks
  int depth, n_lin;
  int use_pm;               // uint8 path: build the phase-major copy
  size_t slot_bytes;
  size_t lvl_off[kCsmMaxDepth];   // bits: byte offset of level l inside a slot (l >= 1)
  size_t pmb_off;                 // bits: bit planes
  int pmb_b0, pmb_b1;             // bits: first columns of the two halves of the paired plane layout (b1 < 0: 64-bit rows)
  size_t pm_off;                  // uint8: phase-major copy (stack at offset 0)
  int max_nx, max_ny;
};

// One (grid, scan) pair.
struct CsmPairDev {
  int grid;            // slot of the batch's working set (filled on the device)
  int gid;             // grid id in the store
  long long pt_begin;  // first point of the scan in the concatenated xyz array
  int n_pts;
  float w0, z0;        // quaternion (w, z) of the float initial yaw (host libm)
  float tx, ty;        // float initial translation
};

struct CsmParams {
  int n_lin, n_ang, S, depth;
  int step;       // 1 << (depth-1): coarse lattice step = coarsest width
  int max_side;   // max coarse candidates per axis per scan
  int maxc;       // max_side^2: slots per scan in the coarse arrays
  unsigned W;     // 2*n_lin+1 (rank radix)
  float min_score;
  float min_s;    // PrecomputationGrid2D::min_score_  (1 - kMaxCorrespondenceCost)
  float coef;     // (max_score_ - min_score_) / 255.f  (ToScore)
};

struct CsmBounds {
  int min_x, max_x, min_y, max_y;
};

// A branch-and-bound node: candidate (scan s, offset xo, yo) at tree depth d (its score is
// the bound of the 2^d x 2^d block of fine offsets it covers).
struct CsmNode {
  int pi, s, xo, yo, d;
  float score;
};

// K5
cudaError_t launch_csm_level1_from_cells(const uint16_t* cells, const uint8_t* lut, size_t n,
                                         uint8_t* out, cudaStream_t stream);
// ---- per-batch working set (all batched over the slots of the batch; no host sync)
// distinct grids of pairs [0, n_pairs) -> slots; pairs[i].grid = slot of pairs[i].gid
cudaError_t launch_csm_assign_slots(CsmPairDev* pairs, int n_pairs, int* hash_keys, int* hash_slot,
                                    int hash_size, int* slot_gid, int* n_slots, cudaStream_t stream);
cudaError_t launch_csm_prepare_slots(const CsmGridRec* recs, const int* slot_gid, const int* n_slots,
                                     int max_slots, CsmPlan plan, unsigned char* ws, CsmGridDev* out,
                                     cudaStream_t stream);
// every derived structure of every slot (levels 1..depth-1 and planes, or uint8 stack + pm)
cudaError_t launch_csm_build_slots(const CsmGridRec* recs, const int* slot_gid, const CsmGridDev* slots,
                                   const int* n_slots, int max_slots, CsmPlan plan, cudaStream_t stream,
                                   uint64_t* launches);
// host <-> store encodings
cudaError_t launch_csm_pack_bits(const uint8_t* level1, int nx, int ny, unsigned* out, int* not_binary,
                                 cudaStream_t stream);
cudaError_t launch_csm_unpack_bits(const unsigned* bits, int nx, int ny, uint8_t* out, cudaStream_t stream);
// K6 (standalone, for parity tests of GenerateRotatedScans + DiscretizeScans)
cudaError_t launch_csm_discretize(const float* pts, int n_pts, float w0, float z0, float tx,
                                  float ty, const float2* rot, int S, double resolution,
                                  double max_x, double max_y, int* out_cells, cudaStream_t stream);
int csm_pmb_rows(int wide_ny, int n_lin, int log2w, bool paired);
size_t csm_coarse_bits_smem(int log2w, int rows);
cudaError_t launch_csm_coarse_bits(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                                   const float* pts, const float2* rot, CsmParams prm,
                                   CsmBounds* bounds, int* coarse, unsigned long long* top_coarse,
                                   size_t smem, int warps, bool paired, cudaStream_t stream);
// K7 pipeline
cudaError_t launch_csm_coarse(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                              const float* pts, const float2* rot, CsmParams prm,
                              CsmBounds* bounds, int* coarse, unsigned long long* top_coarse,
                              cudaStream_t stream, bool phase_major);
cudaError_t launch_csm_seed(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                            const float* pts, const float2* rot, CsmParams prm,
                            const CsmBounds* bounds, const unsigned long long* top_coarse,
                            unsigned long long* best, cudaStream_t stream);
// survivors: [n_pairs][S * maxc] (entry = scan * maxc + slot), n_survivors: [n_pairs]
cudaError_t launch_csm_filter(const CsmPairDev* pairs, int n_pairs, CsmParams prm,
                              const CsmBounds* bounds, const int* coarse,
                              const unsigned long long* best, unsigned* survivors,
                              unsigned* n_survivors, cudaStream_t stream);
size_t csm_expand_smem(int wide_ny, int stride);
// the four children of every surviving coarse candidate -> nodes (or leaves at depth 2);
// bits == false: the survivors themselves become the nodes
cudaError_t launch_csm_expand(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                              const float* pts, const float2* rot, CsmParams prm,
                              const CsmBounds* bounds, const int* coarse,
                              const unsigned* survivors, const unsigned* n_survivors,
                              unsigned long long* best, CsmNode* nodes, unsigned* n_nodes,
                              unsigned node_cap, unsigned long long* counters, int chunks,
                              bool bits, size_t smem, cudaStream_t stream);
cudaError_t launch_csm_refine(const CsmGridDev* grids, const CsmPairDev* pairs, const float* pts,
                              const float2* rot, CsmParams prm, const CsmBounds* bounds,
                              const CsmNode* nodes, const unsigned* n_nodes, unsigned node_cap,
                              unsigned* cursor, unsigned long long* best,
                              unsigned long long* counters, int n_ctas, cudaStream_t stream);

}  // namespace gloc
