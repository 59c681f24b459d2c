// csm.cu -- stage 2 kernels: precomputation stack (K5), rotate + discretise (K6),
// candidate scoring + parallel branch and bound (K7).
//
// Replaces, from /root/reference/registration/2d:
//   PrecomputationGrid2D ctor            fast_correlative_scan_matcher_2d.cpp:112-182
//   GenerateRotatedScans/DiscretizeScans correlative_scan_matcher_2d.cpp:93-127
//   SearchParameters::ShrinkToFit        correlative_scan_matcher_2d.cpp:73-91
//   ScoreCandidates                      fast_correlative_scan_matcher_2d.cpp:372-391
//   BranchAndBound                       fast_correlative_scan_matcher_2d.cpp:393-438
//
// Exactness contract: candidate sums are integers; the score is
// ToScore(sum / float(P)) in un-fused float32 (fast_..._2d.h:86-88, .cpp:386-387);
// the answer is the maximum-score (scan, x, y) over the whole shrunk window with
// ties resolved to the smallest (scan, x, y) -- independent of the order in which
// the GPU explores the tree (64-bit atomicMax on (score bits, inverted rank)).
// Point rotation/discretisation is bit-exact with the reference's float/double
// arithmetic: the quaternion (w, z) per angle comes from the host's libm, the
// device uses only round-to-nearest intrinsics (no FMA contraction).
#include <algorithm>

#include "csm_kernels.cuh"

namespace gloc {

namespace {

constexpr int kPointChunk = 4096;  // discretised points cached in shared memory per pass
constexpr int kCoarseThreads = 256;
constexpr int kCandPerThread = 4;
constexpr int kBitRowSlack = 17;    // zero plane rows after the last data row (candidate rows read past it)
// Rows of one bit plane.  The plane stride (rows, or rows / 2 words per half in the paired layout) is
// kept odd: lanes of a warp sit in different planes, an even stride would fold them onto half of the
// shared-memory banks (measured: 23 -> 37 ms for the coarse scorer).
__host__ __device__ inline int csm_pmb_rows_of(int wide_ny, int n_lin, int log2w, bool paired) {
  const int need = ((wide_ny + 3 * n_lin - 1) >> log2w) + 1 + kBitRowSlack;
  return paired ? 2 * (((need + 1) >> 1) | 1) : (need | 1);
}

// Eigen Quaternionf(AngleAxisf(theta, UnitZ)) * v  with vec = (0, 0, z):
//   uv = vec x v; uv += uv; v' = v + w*uv + vec x uv      (see oracle/csm_oracle.c)
__device__ __forceinline__ void rot_z(float w, float z, float vx, float vy, float& rx,
                                      float& ry) {
  float ux = -__fmul_rn(z, vy);
  float uy = __fmul_rn(z, vx);
  ux = __fadd_rn(ux, ux);
  uy = __fadd_rn(uy, uy);
  const float cx = -__fmul_rn(z, uy);
  const float cy = __fmul_rn(z, ux);
  rx = __fadd_rn(__fadd_rn(vx, __fmul_rn(w, ux)), cx);
  ry = __fadd_rn(__fadd_rn(vy, __fmul_rn(w, uy)), cy);
}

// fast_..._2d.cpp:278-283 (initial yaw), correlative_..._2d.cpp:104-106 (scan angle),
// :119-121 (translation), map_limits.h:69-76 (GetCellIndex; lround = half away from zero)
__device__ __forceinline__ int2 discretize_point(const float* __restrict__ p, float w0, float z0,
                                                 float ws, float zs, float tx, float ty,
                                                 double res, double max_x, double max_y) {
  float x0, y0, x1, y1;
  rot_z(w0, z0, p[0], p[1], x0, y0);
  rot_z(ws, zs, x0, y0, x1, y1);
  const float wx = __fadd_rn(x1, tx);
  const float wy = __fadd_rn(y1, ty);
  int2 c;
  c.x = (int)round(__dsub_rn(__ddiv_rn(__dsub_rn(max_y, (double)wy), res), 0.5));
  c.y = (int)round(__dsub_rn(__ddiv_rn(__dsub_rn(max_x, (double)wx), res), 0.5));
  return c;
}

struct LevelView {
  const uint8_t* cells;     // uint8 level (binary == 0)
  const unsigned* bits;     // bit-packed level (binary == 1)
  int stride;               // words per bit-packed row
  int wide_nx, wide_ny, wm1;
};

__device__ __forceinline__ LevelView level_view(const CsmGridDev& g, int level) {
  LevelView v;
  const int w = 1 << level;
  v.cells = g.binary ? nullptr : g.stack + g.off[level];
  v.bits = g.binary ? g.lvl[level] : nullptr;
  v.stride = g.lvs[level];
  v.wide_nx = g.nx + w - 1;
  v.wide_ny = g.ny + w - 1;
  v.wm1 = w - 1;
  return v;
}

// PrecomputationGrid2D::GetValue, fast_..._2d.h:68-83 (a binary grid holds 0 / 255 only)
__device__ __forceinline__ int level_val(const LevelView& v, int x, int y) {
  const unsigned lx = (unsigned)(x + v.wm1), ly = (unsigned)(y + v.wm1);
  if (lx >= (unsigned)v.wide_nx || ly >= (unsigned)v.wide_ny) return 0;
  if (v.bits) return (int)((__ldg(v.bits + (size_t)ly * v.stride + (lx >> 5)) >> (lx & 31)) & 1u) * 255;
  return __ldg(v.cells + (size_t)ly * v.wide_nx + lx);
}

// ToScore(sum / static_cast<float>(P)), fast_..._2d.cpp:386-387 + .h:86-88
__device__ __forceinline__ float score_of(int sum, int n_pts, const CsmParams& prm) {
  const float v = __fdiv_rn((float)sum, (float)n_pts);
  return __fadd_rn(prm.min_s, __fmul_rn(v, prm.coef));
}

__device__ __forceinline__ unsigned rank_of(const CsmParams& prm, int s, int xo, int yo) {
  return ((unsigned)s * prm.W + (unsigned)(xo + prm.n_lin)) * prm.W + (unsigned)(yo + prm.n_lin);
}

// Larger key = better: higher score, then smaller (scan, x, y).
__device__ __forceinline__ unsigned long long key_of(float score, unsigned rank) {
  return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - rank);
}

__device__ __forceinline__ unsigned long long ld_best(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------- K5

__global__ void csm_level1_from_cells_kernel(const uint16_t* __restrict__ cells,
                                             const uint8_t* __restrict__ lut, size_t n,
                                             uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = lut[cells[i]];
}

// ---- store encodings

// uint8 width-1 grid -> bit-packed rows (bit x of row y at word y * stride + x / 32, stride =
// csm_bit_stride(nx), zero beyond nx); *not_binary is set when a cell is neither 0 nor 255.
__global__ void csm_pack_bits_kernel(const uint8_t* __restrict__ level1, int nx, int ny, int stride,
                                     unsigned* __restrict__ out, int* __restrict__ not_binary) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ny * stride) return;
  const int y = idx / stride, wx = idx % stride;
  unsigned bits = 0;
  bool odd = false;
  for (int i = 0; i < 32; ++i) {
    const int x = wx * 32 + i;
    if (x < nx) {
      const unsigned v = level1[(size_t)y * nx + x];
      if (v) bits |= 1u << i;
      odd |= (v != 0u && v != 255u);
    }
  }
  out[idx] = bits;
  if (odd) atomicOr(not_binary, 1);
}

__global__ void csm_unpack_bits_kernel(const unsigned* __restrict__ bits, int nx, int ny, int stride,
                                       uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nx * ny) return;
  const int y = (int)(i / nx), x = (int)(i % nx);
  out[i] = ((bits[(size_t)y * stride + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}

// ---- the working set of a batch: one slot per distinct grid, everything derived from the
// stored width-1 grid on the device, batched over the slots (blockIdx.y = slot)

__device__ __forceinline__ unsigned hash_gid(int gid) {
  unsigned h = (unsigned)gid * 0x9E3779B1u;
  return h ^ (h >> 15);
}

__global__ void csm_slot_insert_kernel(const CsmPairDev* __restrict__ pairs, int n_pairs,
                                       int* __restrict__ keys, int* __restrict__ hslot, int hmask,
                                       int* __restrict__ slot_gid, int* __restrict__ n_slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const int gid = pairs[i].gid;
  unsigned h = hash_gid(gid) & (unsigned)hmask;
  for (;;) {
    const int prev = atomicCAS(keys + h, -1, gid);
    if (prev == -1) {
      const int slot = atomicAdd(n_slots, 1);
      hslot[h] = slot;
      slot_gid[slot] = gid;
      return;
    }
    if (prev == gid) return;
    h = (h + 1) & (unsigned)hmask;
  }
}

__global__ void csm_slot_lookup_kernel(CsmPairDev* __restrict__ pairs, int n_pairs,
                                       const int* __restrict__ keys, const int* __restrict__ hslot,
                                       int hmask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const int gid = pairs[i].gid;
  unsigned h = hash_gid(gid) & (unsigned)hmask;
  while (keys[h] != gid) h = (h + 1) & (unsigned)hmask;
  pairs[i].grid = hslot[h];
}


__global__ void csm_prepare_slots_kernel(const CsmGridRec* __restrict__ recs,
                                         const int* __restrict__ slot_gid,
                                         const int* __restrict__ n_slots, CsmPlan plan,
                                         unsigned char* __restrict__ ws, CsmGridDev* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *n_slots) return;
  const CsmGridRec r = recs[slot_gid[j]];
  unsigned char* base = ws + (size_t)j * plan.slot_bytes;
  CsmGridDev g;
  g.nx = r.nx; g.ny = r.ny; g.resolution = r.resolution; g.max_x = r.max_x; g.max_y = r.max_y;
  g.binary = plan.bits;
  g.stack = nullptr; g.pm = nullptr; g.pmb = nullptr;
  g.pm_pad = g.pm_pw = g.pm_ph = g.pm_log2w = 0;
  g.pmb_rows = g.pmb_px = g.pmb_py = g.pmb_log2w = 0;
  g.pmb_b0 = 0; g.pmb_b1 = -1;
  for (int l = 0; l < kCsmMaxDepth; ++l) { g.off[l] = 0; g.lvl[l] = nullptr; g.lvs[l] = 0; }
  const int top = plan.depth - 1, w = 1 << top;
  if (plan.bits) {
    g.lvl[0] = reinterpret_cast<const unsigned*>(r.data);
    g.lvs[0] = csm_bit_stride(r.nx);
    for (int l = 1; l < plan.depth; ++l) {
      g.lvl[l] = reinterpret_cast<const unsigned*>(base + plan.lvl_off[l]);
      g.lvs[l] = csm_bit_stride(r.nx + (1 << l) - 1);
    }
    g.pmb = reinterpret_cast<const unsigned long long*>(base + plan.pmb_off);
    g.pmb_rows = csm_pmb_rows_of(r.ny + w - 1, plan.n_lin, top, plan.pmb_b1 >= 0);
    g.pmb_px = plan.n_lin; g.pmb_py = 2 * plan.n_lin; g.pmb_log2w = top;
    g.pmb_b0 = plan.pmb_b0; g.pmb_b1 = plan.pmb_b1;
  } else {
    g.stack = base;
    long long total = 0;
    for (int l = 0; l < plan.depth; ++l) {
      const int wl = 1 << l;
      g.off[l] = total;
      total += (long long)(r.nx + wl - 1) * (long long)(r.ny + wl - 1);
      total = (total + 15) & ~15ll;
    }
    if (plan.use_pm) {
      const int wide_nx = r.nx + w - 1, wide_ny = r.ny + w - 1, pad = plan.n_lin;
      g.pm = base + plan.pm_off;
      g.pm_pad = pad;
      g.pm_pw = (wide_nx + 2 * pad + w - 1) / w;
      g.pm_ph = (wide_ny + 2 * pad + w - 1) / w;
      g.pm_log2w = top;
    }
  }
  out[j] = g;
}

// One 64-bit row of a bit plane: bit c = bit (w c + rx - px) of row ly of the coarsest bit level
// (zero outside the level).
__device__ __forceinline__ unsigned long long pmb_row_bits(const unsigned* level, int stride, int wide_nx,
                                                           int wide_ny, int log2w, int ly, int rx, int px) {
  unsigned long long bits = 0;
  if ((unsigned)ly >= (unsigned)wide_ny) return bits;
  if (log2w == 4) {
    // every 64-bit word of the level row (rows are 8-byte aligned, zero beyond wide_nx) holds four of
    // the wanted bits, 16 apart; one multiply gathers the four into a nibble
    const unsigned long long* row = reinterpret_cast<const unsigned long long*>(level + (size_t)ly * stride);
    const int q = rx - px, q16 = q & 15, fq = (q - q16) >> 4;      // column of bit c = 16 (c + fq) + q16
    unsigned long long gathered = 0;                                // bit c' = level bit 16 c' + q16
    const int n64 = min(stride >> 1, 16);
    for (int j = 0; j < n64; ++j) {
      const unsigned long long t = (row[j] >> q16) & 0x0001000100010001ull;
      gathered |= (((t * 0x0001000200040008ull) >> 48) & 0xFull) << (4 * j);
    }
    bits = fq <= 0 ? (-fq < 64 ? gathered << (-fq) : 0ull) : (fq < 64 ? gathered >> fq : 0ull);
  } else {
    const unsigned* row = level + (size_t)ly * stride;
    for (int c = 0; c < 64; ++c) {
      const int lx = (c << log2w) + rx - px;
      if ((unsigned)lx < (unsigned)wide_nx) bits |= (unsigned long long)((row[lx >> 5] >> (lx & 31)) & 1u) << c;
    }
  }
  return bits;
}

// The bit planes of slot g from its coarsest bit level (global or shared memory), by the threads
// first, first + step, ...: either layout of CsmGridDev::pmb.  In GLOBAL memory the words are stored
// row-major -- word (r, plane) at r w^2 + plane, paired: (p, plane, h) at (p w^2 + plane) 2 + h -- so that
// neighbouring threads take neighbouring planes: they read the same level row (a broadcast; threads on
// neighbouring plane ROWS would sit 16 w level rows apart, on the same banks) and write neighbouring
// words.  The scorer transposes to its plane-major shared-memory layout while staging.
__device__ __forceinline__ void pmb_write_planes(const CsmGridDev& g, const unsigned* level, int first, int step) {
  const int log2w = g.pmb_log2w, w = 1 << log2w, rows = g.pmb_rows, px = g.pmb_px, py = g.pmb_py;
  const int wide_nx = g.nx + w - 1, wide_ny = g.ny + w - 1, stride = g.lvs[log2w];
  unsigned long long* out = const_cast<unsigned long long*>(g.pmb);
  const int pm = w * w - 1, lp = 2 * log2w;
  if (g.pmb_b1 < 0) {
    const int n = rows << lp;
    for (int idx = first; idx < n; idx += step) {
      const int plane = idx & pm, r = idx >> lp;
      const int ry = plane >> log2w, rx = plane & (w - 1);
      out[idx] = pmb_row_bits(level, stride, wide_nx, wide_ny, log2w, (r << log2w) + ry - py, rx, px);
    }
  } else {
    const int n = (rows >> 1) << lp, b0 = g.pmb_b0, b1 = g.pmb_b1;
    ulonglong2* out2 = reinterpret_cast<ulonglong2*>(out);
    for (int idx = first; idx < n; idx += step) {
      const int plane = idx & pm, rp = idx >> lp;
      const int ry = plane >> log2w, rx = plane & (w - 1);
      const int ly = ((2 * rp) << log2w) + ry - py;
      const unsigned long long a = pmb_row_bits(level, stride, wide_nx, wide_ny, log2w, ly, rx, px);
      const unsigned long long b = pmb_row_bits(level, stride, wide_nx, wide_ny, log2w, ly + w, rx, px);
      out2[idx] = make_ulonglong2((unsigned long long)(uint32_t)(a >> b0) | ((unsigned long long)(uint32_t)(b >> b0) << 32),
                                  (unsigned long long)(uint32_t)(a >> b1) | ((unsigned long long)(uint32_t)(b >> b1) << 32));
    }
  }
}

// bit level l from bit level l-1: the max over a w x w window is the OR of four w/2 x w/2
// windows, i.e. out = A | (A << h) with A = row(y) | row(y - h) of the previous level.
__global__ void csm_build_lvl_bits_kernel(const CsmGridDev* __restrict__ slots,
                                          const int* __restrict__ n_slots, int l) {
  if ((int)blockIdx.y >= *n_slots) return;
  const CsmGridDev& g = slots[blockIdx.y];
  const int w = 1 << l, h = w >> 1;
  const int wny = g.ny + w - 1, pny = g.ny + h - 1;
  const int stride = g.lvs[l], ps = g.lvs[l - 1];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wny * stride) return;
  const int ly = idx / stride, j = idx % stride;
  const unsigned* prev = g.lvl[l - 1];
  auto A = [&](int jj) -> unsigned {
    if (jj < 0 || jj >= ps) return 0u;
    unsigned v = 0u;
    if (ly < pny) v |= __ldg(prev + (size_t)ly * ps + jj);
    if (ly - h >= 0 && ly - h < pny) v |= __ldg(prev + (size_t)(ly - h) * ps + jj);
    return v;
  };
  const int hs = h >> 5, hb = h & 31;
  const unsigned sh = hb ? __funnelshift_l(A(j - hs - 1), A(j - hs), hb) : A(j - hs);
  const_cast<unsigned*>(g.lvl[l])[idx] = A(j) | sh;
}

// The whole binary working set of one slot in ONE kernel (a CTA per slot): the stored width-1 grid
// goes to shared memory once, every coarser level is derived there from the previous one (two
// buffers alternate), written out for the expand / refine stages, and the coarsest level leaves
// shared memory as bit planes only.  Replaces depth-1 level launches + the plane launch and their
// round trips through L2 (1.8 -> 0.5 ms per 1400 slots); grids too large for two levels in shared
// memory keep the per-level kernels.
__global__ void __launch_bounds__(512)
csm_build_slot_fused_kernel(const CsmGridDev* __restrict__ slots, const int* __restrict__ n_slots, int depth,
                            int buf_words) {
  if ((int)blockIdx.x >= *n_slots) return;
  extern __shared__ __align__(16) unsigned char csm_smem[];
  unsigned* buf[2] = {reinterpret_cast<unsigned*>(csm_smem), reinterpret_cast<unsigned*>(csm_smem) + buf_words};
  const CsmGridDev& g = slots[blockIdx.x];
  const int tid = threadIdx.x, top = depth - 1;
  {   // level 0 as stored
    const int n0 = g.ny * g.lvs[0];
    for (int i = tid; i < n0; i += 512) buf[0][i] = __ldg(g.lvl[0] + i);
  }
  __syncthreads();
  for (int l = 1; l <= top; ++l) {
    const unsigned* prev = buf[(l - 1) & 1];
    unsigned* cur = buf[l & 1];
    const int w = 1 << l, h = w >> 1;
    const int wny = g.ny + w - 1, pny = g.ny + h - 1;
    const int stride = g.lvs[l], ps = g.lvs[l - 1];
    const int hs = h >> 5, hb = h & 31;
    unsigned* out = l < top ? const_cast<unsigned*>(g.lvl[l]) : nullptr;   // the coarsest level is only needed as planes
    // a warp per row: lane = word of the row (strides beyond 32 words take several rounds)
    for (int ly = tid >> 5; ly < wny; ly += 16)
    for (int j = tid & 31; j < stride; j += 32) {
      const int idx = ly * stride + j;
      auto A = [&](int jj) -> unsigned {
        if (jj < 0 || jj >= ps) return 0u;
        unsigned v = 0u;
        if (ly < pny) v |= prev[ly * ps + jj];
        if (ly - h >= 0 && ly - h < pny) v |= prev[(ly - h) * ps + jj];
        return v;
      };
      const unsigned sh = hb ? __funnelshift_l(A(j - hs - 1), A(j - hs), hb) : A(j - hs);
      const unsigned v = A(j) | sh;
      cur[idx] = v;
      if (out) out[idx] = v;
    }
    __syncthreads();
  }
  // bit planes of the coarsest level (see csm_build_pmb_kernel)
  pmb_write_planes(g, buf[top & 1], tid, 512);
}

// uint8 path: width-1 grid of a slot from the store's record (bits or bytes)
__global__ void csm_slot_level0_u8_kernel(const CsmGridRec* __restrict__ recs,
                                          const int* __restrict__ slot_gid,
                                          const CsmGridDev* __restrict__ slots,
                                          const int* __restrict__ n_slots) {
  if ((int)blockIdx.y >= *n_slots) return;
  const CsmGridDev& g = slots[blockIdx.y];
  const CsmGridRec r = recs[slot_gid[blockIdx.y]];
  const size_t n = (size_t)g.nx * g.ny;
  uint8_t* out = const_cast<uint8_t*>(g.stack);
  const int stride = csm_bit_stride(g.nx);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (r.enc == 1) {
      const int y = (int)(i / g.nx), x = (int)(i % g.nx);
      const unsigned wd = reinterpret_cast<const unsigned*>(r.data)[(size_t)y * stride + (x >> 5)];
      out[i] = ((wd >> (x & 31)) & 1u) ? 255 : 0;
    } else {
      out[i] = reinterpret_cast<const uint8_t*>(r.data)[i];
    }
  }
}

// Width-w grid from the width-w/2 grid: the max over a w x w window is the max of
// four w/2 x w/2 windows.  Out-of-stack reads are windows outside the map: 0.
__global__ void csm_build_level_kernel(const CsmGridDev* __restrict__ slots,
                                       const int* __restrict__ n_slots, int l) {
  if ((int)blockIdx.y >= *n_slots) return;
  const CsmGridDev& g = slots[blockIdx.y];
  const int w = 1 << l, h = w >> 1, nx = g.nx, ny = g.ny;
  const int wide_nx = nx + w - 1, wide_ny = ny + w - 1;
  const int pnx = nx + h - 1, pny = ny + h - 1;
  const uint8_t* prev = g.stack + g.off[l - 1];
  uint8_t* out = const_cast<uint8_t*>(g.stack) + g.off[l];
  const size_t n = (size_t)wide_nx * wide_ny;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ly = (int)(i / wide_nx), lx = (int)(i % wide_nx);
    // window origin in map cells, then position in the previous level's local frame
    const int px = lx - (w - 1) + (h - 1), py = ly - (w - 1) + (h - 1);
    int m = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int x = px + dx * h, y = py + dy * h;
        if ((unsigned)x < (unsigned)pnx && (unsigned)y < (unsigned)pny)
          m = max(m, (int)prev[(size_t)y * pnx + x]);
      }
    out[i] = (uint8_t)m;
  }
}

// ------------------------------------------------------------------------- K6

__global__ void csm_discretize_kernel(const float* __restrict__ pts, int n_pts, float w0, float z0,
                                      float tx, float ty, const float2* __restrict__ rot, int S,
                                      double res, double max_x, double max_y,
                                      int* __restrict__ out_cells) {
  const int s = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S || p >= n_pts) return;
  const float2 r = rot[s];
  const int2 c = discretize_point(pts + 3 * (size_t)p, w0, z0, r.x, r.y, tx, ty, res, max_x, max_y);
  out_cells[2 * ((size_t)s * n_pts + p)] = c.x;
  out_cells[2 * ((size_t)s * n_pts + p) + 1] = c.y;
}

// ------------------------------------------------------------------- K7 coarse

// One CTA per (scan, pair): discretise the scan once into shared memory, shrink the
// window (ShrinkToFit), score every lattice candidate of this rotation on the coarsest
// grid, and reduce the best (score, rank) of the rotation bin with warp shuffles.
__global__ void __launch_bounds__(kCoarseThreads)
csm_coarse_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                  const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                  CsmBounds* __restrict__ bounds, int* __restrict__ coarse,
                  unsigned long long* __restrict__ top_coarse) {
  __shared__ int2 cells[kPointChunk];
  __shared__ int red[4][kCoarseThreads / 32];
  __shared__ unsigned long long redk[kCoarseThreads / 32];
  __shared__ CsmBounds sb;
  const int s = blockIdx.x, pi = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CsmPairDev pr = pairs[pi];
  const CsmGridDev g = grids[pr.grid];
  const float2 r = rot[s];
  const int P = pr.n_pts;
  const float* sp = pts + 3 * (size_t)pr.pt_begin;

  // ShrinkToFit, correlative_scan_matcher_2d.cpp:77-90
  int mnx = 0, mny = 0, mxx = 0, mxy = 0;
  for (int p = tid; p < P; p += kCoarseThreads) {
    const int2 c = discretize_point(sp + 3 * (size_t)p, pr.w0, pr.z0, r.x, r.y, pr.tx, pr.ty,
                                    g.resolution, g.max_x, g.max_y);
    if (p < kPointChunk) cells[p] = c;
    mnx = min(mnx, -c.x);
    mny = min(mny, -c.y);
    mxx = max(mxx, g.nx - 1 - c.x);
    mxy = max(mxy, g.ny - 1 - c.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) {
    red[0][warp] = mnx;
    red[1][warp] = mny;
    red[2][warp] = mxx;
    red[3][warp] = mxy;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < kCoarseThreads / 32; ++w2) {
      mnx = min(mnx, red[0][w2]);
      mny = min(mny, red[1][w2]);
      mxx = max(mxx, red[2][w2]);
      mxy = max(mxy, red[3][w2]);
    }
    CsmBounds b;
    b.min_x = max(-prm.n_lin, mnx);
    b.max_x = min(prm.n_lin, mxx);
    b.min_y = max(-prm.n_lin, mny);
    b.max_y = min(prm.n_lin, mxy);
    sb = b;
    bounds[(size_t)pi * prm.S + s] = b;
  }
  __syncthreads();
  const CsmBounds b = sb;
  // GenerateLowestResolutionCandidates, fast_..._2d.cpp:334-370
  const int ncx = (b.max_x - b.min_x + prm.step) / prm.step;
  const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;
  const int ncand = ncx * ncy;
  const LevelView lv = level_view(g, prm.depth - 1);
  int* out = coarse + ((size_t)pi * prm.S + s) * prm.maxc;
  unsigned long long best_key = 0;

  for (int cb = 0; cb < ncand; cb += kCoarseThreads * kCandPerThread) {
    int sum[kCandPerThread], xo[kCandPerThread], yo[kCandPerThread];
    bool ok[kCandPerThread];
#pragma unroll
    for (int j = 0; j < kCandPerThread; ++j) {
      const int c = cb + j * kCoarseThreads + tid;  // x fastest: adjacent lanes, adjacent bytes
      ok[j] = c < ncand;
      const int iy = ok[j] ? c / ncx : 0, ix = ok[j] ? c % ncx : 0;
      xo[j] = b.min_x + ix * prm.step;
      yo[j] = b.min_y + iy * prm.step;
      sum[j] = 0;
    }
    for (int p0 = 0; p0 < P; p0 += kPointChunk) {
      const int n = min(kPointChunk, P - p0);
      if (P > kPointChunk) {
        __syncthreads();
        for (int p = tid; p < n; p += kCoarseThreads)
          cells[p] = discretize_point(sp + 3 * (size_t)(p0 + p), pr.w0, pr.z0, r.x, r.y, pr.tx,
                                      pr.ty, g.resolution, g.max_x, g.max_y);
        __syncthreads();
      }
#pragma unroll 4
      for (int p = 0; p < n; ++p) {
        const int2 c = cells[p];
#pragma unroll
        for (int j = 0; j < kCandPerThread; ++j)
          if (ok[j]) sum[j] += level_val(lv, c.x + xo[j], c.y + yo[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kCandPerThread; ++j) {
      if (!ok[j]) continue;
      const int ix = (xo[j] - b.min_x) / prm.step, iy = (yo[j] - b.min_y) / prm.step;
      out[ix * ncy + iy] = sum[j];  // reference enumeration order: x outer, y inner
      const unsigned long long key =
          key_of(score_of(sum[j], P, prm), rank_of(prm, s, xo[j], yo[j]));
      best_key = max(best_key, key);
    }
  }
  // best candidate of this rotation bin -> best coarse candidate of the pair
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    best_key = max(best_key, __shfl_xor_sync(0xffffffffu, best_key, o));
  if (lane == 0) redk[warp] = best_key;
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < kCoarseThreads / 32; ++w2) best_key = max(best_key, redk[w2]);
    atomicMax(top_coarse + pi, best_key);
  }
}

// ------------------------------------------------------------- K7 coarse (phase-major)

__global__ void csm_build_pm_kernel(const CsmGridDev* __restrict__ slots,
                                    const int* __restrict__ n_slots) {
  if ((int)blockIdx.y >= *n_slots) return;
  const CsmGridDev& g = slots[blockIdx.y];
  const int log2w = g.pm_log2w, w = 1 << log2w, pad = g.pm_pad, pw = g.pm_pw, ph = g.pm_ph;
  const int wide_nx = g.nx + w - 1, wide_ny = g.ny + w - 1;
  const uint8_t* level = g.stack + g.off[log2w];
  uint8_t* out = const_cast<uint8_t*>(g.pm);
  const size_t n = (size_t)pw * w * ph * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % ((size_t)pw * w)), Y = (int)(i / ((size_t)pw * w));   // padded coordinates
    const int lx = X - pad, ly = Y - pad;
    uint8_t v = 0;
    if ((unsigned)lx < (unsigned)wide_nx && (unsigned)ly < (unsigned)wide_ny)
      v = level[(size_t)ly * wide_nx + lx];
    const size_t plane = (size_t)((Y & (w - 1)) * w + (X & (w - 1)));
    out[(plane * ph + (Y >> log2w)) * pw + (X >> log2w)] = v;
  }
}

constexpr int kPmThreads = 512;

// One CTA per (scan, pair).  Phase 1: discretise the scan into shared memory and shrink the
// window (ShrinkToFit).  Phase 2: points inside the padded grid become one int32 base offset
// into the phase-major coarse level (compacted; border points keep their cell for a checked
// path).  Phase 3: thread = (lattice candidate, point slice); the inner loop is
// base + candidate offset -> one byte load -> add, no bounds checks, adjacent lanes read
// adjacent bytes.  Slices are combined with shared-memory atomics; the best (score, rank) of
// the rotation bin is reduced with warp shuffles.
__global__ void __launch_bounds__(kPmThreads)
csm_coarse_pm_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                     const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                     CsmBounds* __restrict__ bounds, int* __restrict__ coarse,
                     unsigned long long* __restrict__ top_coarse) {
  extern __shared__ __align__(16) unsigned char csm_smem[];
  int2* cells = reinterpret_cast<int2*>(csm_smem);                        // [kPointChunk]
  int* bases = reinterpret_cast<int*>(cells + kPointChunk);               // [kPointChunk]
  int2* border = reinterpret_cast<int2*>(bases + kPointChunk);            // [kPointChunk]
  int* sums = reinterpret_cast<int*>(border + kPointChunk);               // [prm.maxc]
  __shared__ int red[4][kPmThreads / 32];
  __shared__ unsigned long long redk[kPmThreads / 32];
  __shared__ CsmBounds sb;
  __shared__ int n_in, n_bd;
  const int s = blockIdx.x, pi = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CsmPairDev pr = pairs[pi];
  const CsmGridDev g = grids[pr.grid];
  const float2 r = rot[s];
  const int P = pr.n_pts;
  const float* sp = pts + 3 * (size_t)pr.pt_begin;

  int mnx = 0, mny = 0, mxx = 0, mxy = 0;
  for (int p = tid; p < P; p += kPmThreads) {
    const int2 c = discretize_point(sp + 3 * (size_t)p, pr.w0, pr.z0, r.x, r.y, pr.tx, pr.ty,
                                    g.resolution, g.max_x, g.max_y);
    if (p < kPointChunk) cells[p] = c;
    mnx = min(mnx, -c.x);
    mny = min(mny, -c.y);
    mxx = max(mxx, g.nx - 1 - c.x);
    mxy = max(mxy, g.ny - 1 - c.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) {
    red[0][warp] = mnx;
    red[1][warp] = mny;
    red[2][warp] = mxx;
    red[3][warp] = mxy;
  }
  for (int i = tid; i < prm.maxc; i += kPmThreads) sums[i] = 0;
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < kPmThreads / 32; ++w2) {
      mnx = min(mnx, red[0][w2]);
      mny = min(mny, red[1][w2]);
      mxx = max(mxx, red[2][w2]);
      mxy = max(mxy, red[3][w2]);
    }
    CsmBounds b;
    b.min_x = max(-prm.n_lin, mnx);
    b.max_x = min(prm.n_lin, mxx);
    b.min_y = max(-prm.n_lin, mny);
    b.max_y = min(prm.n_lin, mxy);
    sb = b;
    bounds[(size_t)pi * prm.S + s] = b;
  }
  __syncthreads();
  const CsmBounds b = sb;
  const int ncx = (b.max_x - b.min_x + prm.step) / prm.step;
  const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;
  const int ncand = ncx * ncy;
  const int wm1 = prm.step - 1, log2w = g.pm_log2w, pad = g.pm_pad;
  const int wide_nx = g.nx + wm1, wide_ny = g.ny + wm1;
  const int plane_sz = g.pm_ph * g.pm_pw;
  const LevelView lv = level_view(g, prm.depth - 1);

  // thread -> (candidate, point slice)
  const int slots = ((ncand + 31) / 32) * 32;             // whole warps share a slice
  const int n_slices = max(1, kPmThreads / slots);
  const int my_slice = tid / slots, my_c = tid % slots;
  const bool active = my_slice < n_slices && (slots <= kPmThreads);
  int* out = coarse + ((size_t)pi * prm.S + s) * prm.maxc;

  for (int cb = 0; cb < ncand; cb += kPmThreads) {          // > 512 candidates: extra rounds
    const int c = (slots <= kPmThreads) ? my_c : cb + tid;
    const bool c_ok = c < ncand && (slots <= kPmThreads ? active : true);
    const int iy = c_ok ? c / ncx : 0, ix = c_ok ? c % ncx : 0;   // x fastest across lanes
    const int coff = iy * g.pm_pw + ix;
    const int xo = b.min_x + ix * prm.step, yo = b.min_y + iy * prm.step;
    int sum = 0;
    for (int p0 = 0; p0 < P; p0 += kPointChunk) {
      const int n = min(kPointChunk, P - p0);
      __syncthreads();
      if (tid == 0) { n_in = 0; n_bd = 0; }
      if (P > kPointChunk) {
        for (int p = tid; p < n; p += kPmThreads)
          cells[p] = discretize_point(sp + 3 * (size_t)(p0 + p), pr.w0, pr.z0, r.x, r.y, pr.tx,
                                      pr.ty, g.resolution, g.max_x, g.max_y);
      }
      __syncthreads();
      for (int p = tid; p < n; p += kPmThreads) {
        const int lx = cells[p].x + wm1, ly = cells[p].y + wm1;     // wide-grid coordinates
        if ((unsigned)lx < (unsigned)wide_nx && (unsigned)ly < (unsigned)wide_ny) {
          const int X = lx + b.min_x + pad, Y = ly + b.min_y + pad;  // >= 0: |min| <= n_lin <= pad
          const int plane = (Y & wm1) * prm.step + (X & wm1);
          bases[atomicAdd(&n_in, 1)] = plane * plane_sz + (Y >> log2w) * g.pm_pw + (X >> log2w);
        } else if (lx >= -prm.n_lin && ly >= -prm.n_lin && lx < wide_nx + prm.n_lin &&
                   ly < wide_ny + prm.n_lin) {
          border[atomicAdd(&n_bd, 1)] = cells[p];                    // may reach the grid: checked path
        }                                                            // else: never lands on the grid
      }
      __syncthreads();
      if (c_ok) {
        const uint8_t* L = g.pm + coff;
        const int ni = n_in, stride = (slots <= kPmThreads) ? n_slices : 1;
        int p = (slots <= kPmThreads) ? my_slice : 0;
        for (; p + 3 * stride < ni; p += 4 * stride) {
          const int b0 = bases[p], b1 = bases[p + stride], b2 = bases[p + 2 * stride],
                    b3 = bases[p + 3 * stride];
          sum += (int)__ldg(L + b0) + (int)__ldg(L + b1) + (int)__ldg(L + b2) + (int)__ldg(L + b3);
        }
        for (; p < ni; p += stride) sum += (int)__ldg(L + bases[p]);
        const int nb = n_bd;
        for (int q = (slots <= kPmThreads) ? my_slice : 0; q < nb; q += stride)
          sum += level_val(lv, border[q].x + xo, border[q].y + yo);
      }
    }
    if (c_ok) {
      if (slots <= kPmThreads && n_slices > 1) atomicAdd(&sums[c], sum);
      else sums[c] = sum;
    }
  }
  __syncthreads();
  unsigned long long best_key = 0;
  for (int c = tid; c < ncand; c += kPmThreads) {
    const int iy = c / ncx, ix = c % ncx;
    const int xo = b.min_x + ix * prm.step, yo = b.min_y + iy * prm.step;
    const int sm = sums[c];
    out[ix * ncy + iy] = sm;                                // reference order: x outer, y inner
    best_key = max(best_key, key_of(score_of(sm, P, prm), rank_of(prm, s, xo, yo)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    best_key = max(best_key, __shfl_xor_sync(0xffffffffu, best_key, o));
  if (lane == 0) redk[warp] = best_key;
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < kPmThreads / 32; ++w2) best_key = max(best_key, redk[w2]);
    atomicMax(top_coarse + pi, best_key);
  }
}

// ------------------------------------------------- K7 coarse (binary grids, bit-sliced)
//
// BEV grids of this code base are binary (a pixel is occupied or free: SURVEY F5,
// 3d/submap_3d.cpp:312-324, :412-424), hence every precomputation level holds only
// {0, 255} and a candidate's sum is 255 x (number of scan points on occupied cells).
// For such grids the coarsest level is kept as bit planes: plane (ry, rx) row r is one
// 64-bit word whose bit c is the cell (16 c + rx - px, 16 r + ry - py) (w = 16 as example).
// The whole structure (~170 KB for an 800 x 800 map) lives in shared memory.
//
// Thread = one rotation (scan), looping over all points: a point contributes, per candidate
// row, the `ncx` adjacent bits of one plane row -- one 8-byte shared-memory load and a
// shift -- and two rows are packed into one 32-bit word whose 16-bit halves are column
// masks.  The per-candidate counts are bit-sliced counters across those words: 16 points
// are folded by a carry-save adder tree (15 full adders = 30 LOP3 per word) into one
// weight-16 carry that ripples into 8 high planes, i.e. ~3 logic instructions per point per
// 26 candidates instead of 26 loads + 26 adds.  No cross-thread reduction, no atomics.
//
// Exactness: the cell of a point is the reference's double-precision GetCellIndex; a float
// evaluation is used only when it is provably on the same side of every rounding boundary
// (fractional part further than `delta` from 0 and 1, delta bounding the float error), else
// the double path runs.  ShrinkToFit needs min/max cell indices only, and the cell index is
// a monotone function of the world coordinate, so the bounds come from the exact cells of
// the four extreme coordinates.

constexpr int kBitChunk = 4080;     // points per pass: 16 x 255, so 8 high planes cannot overflow
constexpr int kBitThreads = 384;    // 192 rotations per CTA, two lanes each

__global__ void csm_build_pmb_kernel(const CsmGridDev* __restrict__ slots,
                                     const int* __restrict__ n_slots) {
  if ((int)blockIdx.y >= *n_slots) return;
  const CsmGridDev& g = slots[blockIdx.y];
  pmb_write_planes(g, g.lvl[g.pmb_log2w], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// map_limits.h:69-76 on an already transformed point (double arithmetic, lround)
__device__ __forceinline__ int cell_exact(float wv, double res, double mx) {
  return (int)round(__dsub_rn(__ddiv_rn(__dsub_rn(mx, (double)wv), res), 0.5));
}

__device__ __noinline__ int2 cells_exact(float wx, float wy, double res, double max_x, double max_y) {
  return make_int2(cell_exact(wy, res, max_y), cell_exact(wx, res, max_x));
}

// Two lanes share one rotation: lane parity h takes candidate rows [8h, 8h + 8) (NP packed
// words of two rows each), and the two lanes split the discretisation work -- the even lane
// discretises the even points, the odd lane the odd ones, the (row address, shift, mask) of a
// point travels to the partner lane by shuffle.  Half the counters per thread: twice the
// warps per SM for the same shared-memory tile.
//
// PR (paired plane layout, CsmGridDev::pmb_b1 >= 0): one word holds the 32 relevant columns of TWO
// plane rows, so a lane's 2 NP candidate rows cost NP loads instead of 2 NP -- this kernel is bound
// by shared-memory wavefronts.  The row pair a point starts in has either parity; lane 0 takes the
// candidate rows from 0, lane 1 from 2 NP - 1 (odd), so exactly one of the two lanes reads its rows
// shifted by one against the stored pairs and re-pairs them with one byte permute per word.  The
// last counter row of either lane is not used (4 NP - 2 candidate rows per rotation).
template <int NP, bool PR>
__global__ void __launch_bounds__(kBitThreads, 1)
csm_coarse_bits_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                       const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                       CsmBounds* __restrict__ bounds, int* __restrict__ coarse,
                       unsigned long long* __restrict__ top_coarse) {
  extern __shared__ __align__(16) unsigned char csm_smem[];
  const int pi = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int half = tid & 1;
  const CsmPairDev pr = pairs[pi];
  const CsmGridDev g = grids[pr.grid];
  const int log2w = g.pmb_log2w, w = 1 << log2w, wm = w - 1;
  const int rows = g.pmb_rows;
  const int n_words = (w * w) * rows;
  unsigned long long* bits = reinterpret_cast<unsigned long long*>(csm_smem);
  float2* P0 = reinterpret_cast<float2*>(bits + n_words);
  // global row-major (see pmb_write_planes) -> plane-major: a lane walks ONE plane (or plane half) for a
  // point, and the odd plane stride spreads the lanes of a warp over the banks
  {
    const int np2 = PR ? 2 * w * w : w * w, lp2 = 2 * log2w + (PR ? 1 : 0), per_plane = PR ? rows >> 1 : rows;
    for (int i = tid; i < n_words; i += blockDim.x) bits[(i & (np2 - 1)) * per_plane + (i >> lp2)] = __ldg(g.pmb + i);
  }

  const int P = pr.n_pts;
  const float* sp = pts + 3 * (size_t)pr.pt_begin;
  const int s = (blockIdx.x * blockDim.x + tid) >> 1;
  const bool s_ok = s < prm.S;
  const float2 r = rot[s_ok ? s : 0];
  // points after the initial-yaw rotation (fast_..._2d.cpp:278-283), shared by all rotations
  // A chunk is padded to a multiple of 16 points with a point that no rotation and no offset of the
  // window brings near the grid: one of its coordinates stays more than `far_pt` - |t| from the origin.
  const float far_pt = -4.f * (fabsf((float)g.max_x) + fabsf((float)g.max_y) + fabsf(pr.tx) + fabsf(pr.ty) +
                               (float)(max(g.nx, g.ny) + w + 2 * prm.n_lin + 32) * (float)g.resolution) - 1000.f;
  auto stage = [&](int p0, int n) {
    __syncthreads();
    for (int p = tid; p < n; p += blockDim.x) {
      float x0, y0;
      rot_z(pr.w0, pr.z0, sp[3 * (size_t)(p0 + p)], sp[3 * (size_t)(p0 + p) + 1], x0, y0);
      P0[p] = make_float2(x0, y0);
    }
    if (tid < 16) P0[n + tid] = make_float2(far_pt, far_pt);
    __syncthreads();
  };
  // ---- pass 1: ShrinkToFit bounds of this rotation (correlative_scan_matcher_2d.cpp:73-91).
  // Shortcut: ONE point whose cell lies at least n_lin cells inside the grid on every side makes
  // all four bounds the full window: min_x = max(-n_lin, min(0, min_p(-cx_p))) and min_p(-cx_p) <=
  // -cx_w <= -n_lin, likewise for the other three.  The point nearest to the sensor is such a
  // witness for (nearly) every rotation of a scan taken inside the map; only where it is not do
  // the threads walk all points for the extreme coordinates.
  // (kept in the points' staging area, which is filled only afterwards: the kernel's dynamic
  // shared memory is sized to the opt-in limit, there is no room for a static variable)
  unsigned long long& s_witness = *reinterpret_cast<unsigned long long*>(P0);
  if (tid == 0) s_witness = ~0ull;
  __syncthreads();
  {
    unsigned long long key = ~0ull;
    for (int p = tid; p < P; p += blockDim.x) {
      const float x = sp[3 * (size_t)p], y = sp[3 * (size_t)p + 1];
      const float n2 = x * x + y * y;                                  // >= 0: bit pattern orders like the value
      key = min(key, ((unsigned long long)__float_as_uint(n2) << 32) | (unsigned)p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if (lane == 0) atomicMin(&s_witness, key);
  }
  __syncthreads();
  bool full = true;
  if (s_ok) {
    const int w = (int)(unsigned)(s_witness & 0xFFFFFFFFull);
    const int2 c = discretize_point(sp + 3 * (size_t)w, pr.w0, pr.z0, r.x, r.y, pr.tx, pr.ty, g.resolution, g.max_x,
                                    g.max_y);
    full = c.x >= prm.n_lin && c.x <= g.nx - 1 - prm.n_lin && c.y >= prm.n_lin && c.y <= g.ny - 1 - prm.n_lin;
  }
  CsmBounds b;
  if (__syncthreads_and(full)) {
    b.min_x = -prm.n_lin; b.max_x = prm.n_lin; b.min_y = -prm.n_lin; b.max_y = prm.n_lin;
    stage(0, min(kBitChunk, P));        // pass 2 expects the first chunk of points in shared memory
  } else {
    // extreme world coordinates of this rotation
    float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int p0 = 0; p0 < P; p0 += kBitChunk) {
      const int n = min(kBitChunk, P - p0);
      stage(p0, n);
#pragma unroll 4
      for (int p = half; p < n; p += 2) {
        const float2 q = P0[p];
        float x1, y1;
        rot_z(r.x, r.y, q.x, q.y, x1, y1);
        const float wx = __fadd_rn(x1, pr.tx), wy = __fadd_rn(y1, pr.ty);
        mnx = fminf(mnx, wx); mxx = fmaxf(mxx, wx);
        mny = fminf(mny, wy); mxy = fmaxf(mxy, wy);
      }
    }
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, 1));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, 1));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, 1));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, 1));
    // cell.x = f(wy), cell.y = f(wx), both monotone non-increasing
    const int cx_max = cell_exact(mny, g.resolution, g.max_y), cx_min = cell_exact(mxy, g.resolution, g.max_y);
    const int cy_max = cell_exact(mnx, g.resolution, g.max_x), cy_min = cell_exact(mxx, g.resolution, g.max_x);
    // correlative_scan_matcher_2d.cpp:77-90
    b.min_x = max(-prm.n_lin, min(0, -cx_max));
    b.max_x = min(prm.n_lin, max(0, g.nx - 1 - cx_min));
    b.min_y = max(-prm.n_lin, min(0, -cy_max));
    b.max_y = min(prm.n_lin, max(0, g.ny - 1 - cy_min));
  }
  if (s_ok && half == 0) bounds[(size_t)pi * prm.S + s] = b;
  const int ncx = (b.max_x - b.min_x + prm.step) / prm.step;
  const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;

  // ---- pass 2: bit-sliced scoring
  const int wide_nx = g.nx + wm, wide_ny = g.ny + wm;
  const float mx_f = (float)g.max_x, my_f = (float)g.max_y, ir = (float)(1.0 / g.resolution);
  // |float cell coordinate - exact| <= 2^-24 (|max|/res + 3 |u|) for |u| <= U; 2x safety (a point further out
  // than U cannot reach any window, there a cell off by one changes nothing)
  const float U = (float)(max(wide_nx, wide_ny) + 2 * prm.n_lin + 32);
  const float delta = 1.1920929e-07f * (fmaxf(fabsf(mx_f), fabsf(my_f)) * ir + 3.f * U);
  const float hi1 = 1.f - delta;
  const int offx = wm + b.min_x + g.pmb_px, offy = wm + b.min_y + g.pmb_py;
  const unsigned span_x = (unsigned)(wide_nx + 2 * prm.n_lin), span_y = (unsigned)(wide_ny + 2 * prm.n_lin);
  const int hx0 = wm + prm.n_lin, hy0 = wm + prm.n_lin;
  const int row0 = PR ? half * (2 * NP - 1) : half * 2 * NP;   // first candidate row of this lane
  // PR: row pairs per plane half, first columns of the halves, first all-zero row, per-lane word offset and
  // the byte selectors of the re-pairing permute (0x3210 = as stored, 0x5432 = shifted by one row)
  const int rpc = rows >> 1, b0 = g.pmb_b0, b1 = g.pmb_b1, zrow = rows - 16;
  const unsigned lane_words = (unsigned)(half * (NP - 1));
  const uint32_t sel_base = half ? 0x5432u : 0x3210u, sel_step = half ? 0u - 0x2222u : 0x2222u;

  // (row address, shift / mask description) of this lane's four points of an 8-point group
  // (points p + 2 i + half): all float discretisations first -- four independent chains with
  // no branch between them --, then one rarely taken exact fallback, then the addressing.
  auto point_addr4 = [&](int p, int n, int (&addr)[4], unsigned (&meta)[4]) {
    float wxs[4], wys[4];
    int cxs[4], cys[4];
    unsigned need = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 q = P0[PR ? p + 2 * i + half : min(p + 2 * i + half, n - 1)];
      float x1, y1;
      rot_z(r.x, r.y, q.x, q.y, x1, y1);
      wxs[i] = __fadd_rn(x1, pr.tx);
      wys[i] = __fadd_rn(y1, pr.ty);
      const float uy = (my_f - wys[i]) * ir, ux = (mx_f - wxs[i]) * ir;
      const float fy = floorf(uy), fx = floorf(ux);
      const float dy = uy - fy, dx = ux - fx;
      cxs[i] = (int)fy;
      cys[i] = (int)fx;
      // (|u| >= 2^23 or not finite: the fractional part is 0 or NaN, the test fails by itself)
      if (!(dy > delta && dy < hi1 && dx > delta && dx < hi1)) need |= 1u << i;
    }
    if (need) {   // rare: near a rounding boundary
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (need & (1u << i)) {
          const int2 ce = cells_exact(wxs[i], wys[i], g.resolution, g.max_x, g.max_y);
          cxs[i] = ce.x;
          cys[i] = ce.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cx = cxs[i], cy = cys[i];
      if (PR) {
        // No test whether the point can land on the grid: the planes are zero outside it, so any address
        // inside the planes is a correct one.  A row index outside [0, zrow] is clamped onto zrow, the first
        // of 16 all-zero rows (true rows are all zero there as well); the column window [ax, ax + ncx)
        // either ends inside half 0 or starts inside half 1 (csm_make_plan admits the layout only then),
        // a start beyond half 1 reads the zero rows, and the shift count 32 + a in [1, 63] applied to
        // (row << 32) shifts right for a >= 0 and LEFT (zeros in) for a window that starts before half 0.
        const int X = (int)((unsigned)cx + (unsigned)offx), Y = (int)((unsigned)cy + (unsigned)offy);
        const int ax = X >> log2w, ay = Y >> log2w;
        const int plane = ((Y & wm) << log2w) | (X & wm);
        const int a0 = ax - b0;
        const bool second = a0 > 32 - ncx;
        const int a = second ? a0 - (b1 - b0) : a0;
        const int ayc = a > 31 ? zrow : (int)min((unsigned)ay, (unsigned)zrow);
        const int sh = max(a, -31) + 32;
        addr[i] = (((plane << 1) | (second ? 1 : 0)) * rpc + (ayc >> 1)) | ((sh & 63) << 16) | ((ayc & 1) << 24);
        meta[i] = 0u;
        continue;
      }
      // can the point land on the grid for some offset of the window at all?
      const bool hitable = (unsigned)(cx + hx0) < span_x && (unsigned)(cy + hy0) < span_y && s_ok &&
                           p + 2 * i + half < n;
      const int X = cx + offx, Y = cy + offy;           // Y >= 0 for hitable points
      const int ax = X >> log2w, ay = hitable ? (Y >> log2w) : 0;
      const int plane = hitable ? (((Y & wm) << log2w) | (X & wm)) : 0;
      const int shl = hitable ? max(0, -ax) : 0, axc = min(max(ax, 0), 63);
      const int nbits = (hitable && ax < 64) ? max(ncx - shl, 0) : 0;
      addr[i] = plane * rows + ay;
      meta[i] = (unsigned)axc | ((unsigned)shl << 8) | ((unsigned)nbits << 16);
    }
  };
  // the NP packed words (two candidate rows each) a point adds to this lane's counters
  auto point_words = [&](int addr, unsigned meta, uint32_t (&v)[NP]) {
    const int axc = meta & 63u, shl = (meta >> 8) & 31u;
    const uint32_t m1 = (1u << ((meta >> 16) & 31u)) - 1u, m2 = m1 | (m1 << 16);
    if (PR) {
      const unsigned A = (unsigned)addr;
      const unsigned par0 = A >> 24, sh = __byte_perm(A, 0u, 0x4442);
      // lane 0 starts at candidate row 0, lane 1 at row 2 NP - 1: their first rows have opposite parity
      // within the stored pairs; the lane whose first row is the odd one re-pairs (t[j].hi, t[j+1].lo)
      const uint2* rowp = reinterpret_cast<const uint2*>(bits) + ((A & 0xFFFFu) + lane_words + (par0 & (unsigned)half));
      const uint32_t sel = sel_base + par0 * sel_step;
      uint32_t t[NP + 1];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const uint2 L = rowp[j];
        const uint32_t lo = (uint32_t)(((unsigned long long)L.x << 32) >> sh);
        const uint32_t hi = (uint32_t)(((unsigned long long)L.y << 32) >> sh);
        t[j] = __byte_perm(lo, hi, 0x5410);
      }
      t[NP] = 0u;
      // bits ncx..15 of a 16-bit field hold columns beyond the window: counters nobody reads
#pragma unroll
      for (int j = 0; j < NP; ++j) v[j] = __byte_perm(t[j], t[j + 1], sel);
      return;
    }
    const unsigned long long* rowp = bits + addr + row0;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      // candidate rows beyond the lattice are never read back: skip their loads (fewer
      // shared-memory wavefronts: this kernel is LSU-bound)
      const uint32_t w0 = row0 + 2 * j < prm.max_side ? (uint32_t)(rowp[2 * j] >> axc) : 0u;
      const uint32_t w1 = row0 + 2 * j + 1 < prm.max_side ? (uint32_t)(rowp[2 * j + 1] >> axc) : 0u;
      v[j] = (__byte_perm(w0, w1, 0x5410) & m2) << shl;
    }
  };
  auto csa = [](uint32_t (&h)[NP], uint32_t (&l)[NP], const uint32_t (&x)[NP], const uint32_t (&y)[NP]) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const uint32_t u = l[j] ^ x[j];
      h[j] = (l[j] & x[j]) | (u & y[j]);
      l[j] = u ^ y[j];
    }
  };

  int* out = coarse + ((size_t)pi * prm.S + (s_ok ? s : 0)) * prm.maxc;
  unsigned long long best_key = 0;
  for (int p0 = 0; p0 < P; p0 += kBitChunk) {
    const int n = min(kBitChunk, P - p0);
    if (P > kBitChunk) stage(p0, n);
    uint32_t ones[NP], twos[NP], fours[NP], eights[NP], hi[8][NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      ones[j] = twos[j] = fours[j] = eights[j] = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) hi[i][j] = 0u;
    }
    auto pair_words = [&](int addr, unsigned meta, uint32_t (&ta)[NP]) {   // two points -> carry ta, sum into ones
      const int addr_o = __shfl_xor_sync(0xffffffffu, addr, 1);
      const unsigned meta_o = PR ? 0u : __shfl_xor_sync(0xffffffffu, meta, 1);
      uint32_t v0[NP], v1[NP];
      point_words(addr, meta, v0);
      point_words(addr_o, meta_o, v1);
      csa(ta, ones, v0, v1);
    };
    auto oct_words = [&](int p, uint32_t (&ea)[NP]) {   // 8 points -> carry ea
      int addr[4];
      unsigned meta[4];
      point_addr4(p, n, addr, meta);
      uint32_t fa[NP], fb[NP];
      {
        uint32_t ta[NP], tb[NP];
        pair_words(addr[0], meta[0], ta);
        pair_words(addr[1], meta[1], tb);
        csa(fa, twos, ta, tb);
      }
      {
        uint32_t ta[NP], tb[NP];
        pair_words(addr[2], meta[2], ta);
        pair_words(addr[3], meta[3], tb);
        csa(fb, twos, ta, tb);
      }
      csa(ea, fours, fa, fb);
    };
#pragma unroll 1
    for (int p = 0; p < n; p += 16) {
      uint32_t ea[NP], eb[NP], c16[NP];
      oct_words(p, ea);
      oct_words(p + 8, eb);
      csa(c16, eights, ea, eb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const uint32_t t = hi[i][j] & c16[j];
          hi[i][j] ^= c16[j];
          c16[j] = t;
        }
      }
    }
    // counters -> integer sums (255 per hit), accumulated over the passes in `coarse`
    if (s_ok) {
      const bool first = p0 == 0, last = p0 + n >= P;
#pragma unroll
      for (int ly = 0; ly < (PR ? 2 * NP - 1 : 2 * NP); ++ly) {
        const int iy = row0 + ly;
        if (iy < ncy) {
          const int j = ly >> 1, sh0 = (ly & 1) * 16;
          for (int ix = 0; ix < ncx; ++ix) {
            const int bp = sh0 + ix;
            int cnt = (int)((ones[j] >> bp) & 1u) | (int)(((twos[j] >> bp) & 1u) << 1) |
                      (int)(((fours[j] >> bp) & 1u) << 2) | (int)(((eights[j] >> bp) & 1u) << 3);
#pragma unroll
            for (int i = 0; i < 8; ++i) cnt |= (int)(((hi[i][j] >> bp) & 1u) << (4 + i));
            int* o = out + ix * ncy + iy;   // reference enumeration order: x outer, y inner
            const int sum = 255 * cnt + (first ? 0 : *o);
            *o = sum;
            if (last)
              best_key = max(best_key, key_of(score_of(sum, P, prm),
                                              rank_of(prm, s, b.min_x + ix * prm.step, b.min_y + iy * prm.step)));
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    best_key = max(best_key, __shfl_xor_sync(0xffffffffu, best_key, o));
  if (lane == 0 && best_key) atomicMax(top_coarse + pi, best_key);
}

// --------------------------------------------------------------------- K7 seed

// One CTA per pair: descend greedily from the best coarse candidate to a leaf to get
// a first incumbent (a real fine score), exactly what the reference's DFS does first.
__global__ void __launch_bounds__(256)
csm_seed_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                const CsmBounds* __restrict__ bounds,
                const unsigned long long* __restrict__ top_coarse,
                unsigned long long* __restrict__ best) {
  __shared__ int red[4][8];
  __shared__ int s_choice[3];
  const int pi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long top = top_coarse[pi];
  if (top == 0) return;
  const float top_score = __uint_as_float((unsigned)(top >> 32));
  if (!(top_score > prm.min_score)) return;  // nothing can beat min_score
  const unsigned rank = 0xFFFFFFFFu - (unsigned)(top & 0xFFFFFFFFull);
  const int s = (int)(rank / (prm.W * prm.W));
  int xo = (int)((rank / prm.W) % prm.W) - prm.n_lin;
  int yo = (int)(rank % prm.W) - prm.n_lin;
  if (prm.depth == 1) {
    if (tid == 0) atomicMax(best + pi, top);
    return;
  }
  const CsmPairDev pr = pairs[pi];
  const CsmGridDev g = grids[pr.grid];
  const CsmBounds b = bounds[(size_t)pi * prm.S + s];
  const float2 r = rot[s];
  const int P = pr.n_pts;
  const float* sp = pts + 3 * (size_t)pr.pt_begin;
  for (int d = prm.depth - 1; d > 0; --d) {
    const int h = 1 << (d - 1);
    const LevelView lv = level_view(g, d - 1);
    int sum[4] = {0, 0, 0, 0};
    for (int p = tid; p < P; p += 256) {
      const int2 c = discretize_point(sp + 3 * (size_t)p, pr.w0, pr.z0, r.x, r.y, pr.tx, pr.ty,
                                      g.resolution, g.max_x, g.max_y);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        sum[ch] += level_val(lv, c.x + xo + (ch >> 1) * h, c.y + yo + (ch & 1) * h);
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) sum[ch] = warp_sum(sum[ch]);
    if (lane == 0)
      for (int ch = 0; ch < 4; ++ch) red[ch][warp] = sum[ch];
    __syncthreads();
    if (tid == 0) {
      int bi = -1;
      unsigned long long bk = 0;
      for (int ch = 0; ch < 4; ++ch) {
        const int cx = xo + (ch >> 1) * h, cy = yo + (ch & 1) * h;
        if (cx > b.max_x || cy > b.max_y) continue;  // fast_..._2d.cpp:415-423
        int t = 0;
        for (int w2 = 0; w2 < 8; ++w2) t += red[ch][w2];
        const unsigned long long k2 = key_of(score_of(t, P, prm), rank_of(prm, s, cx, cy));
        if (bi < 0 || k2 > bk) {
          bk = k2;
          bi = ch;
        }
      }
      s_choice[0] = xo + (bi >> 1) * h;
      s_choice[1] = yo + (bi & 1) * h;
      if (d == 1) atomicMax(best + pi, bk);
    }
    __syncthreads();
    xo = s_choice[0];
    yo = s_choice[1];
    __syncthreads();
  }
}

// ------------------------------------------------------------------- K7 filter

// Keep the coarse candidates whose bound can still beat the incumbent: one list per pair
// (entries = scan * maxc + slot), so that the next stage can work pair by pair.
__global__ void csm_filter_kernel(const CsmPairDev* __restrict__ pairs, int n_pairs, CsmParams prm,
                                  const CsmBounds* __restrict__ bounds,
                                  const int* __restrict__ coarse,
                                  const unsigned long long* __restrict__ best,
                                  unsigned* __restrict__ survivors,
                                  unsigned* __restrict__ n_survivors) {
  // One warp per (pair, scan): its survivors are compacted with ballots and appended to the pair's
  // list with ONE atomicAdd, so the survivors of a scan sit next to each other -- the grouped
  // expand kernel discretises a scan once for all of them.
  const unsigned per_pair = (unsigned)prm.S * (unsigned)prm.maxc;   // < 2^32 (checked by the caller)
  const unsigned lane = threadIdx.x & 31u;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned total = (unsigned)n_pairs * (unsigned)prm.S;
  for (unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < total; gw += n_warps) {
    const unsigned pi = gw / (unsigned)prm.S, s = gw - pi * (unsigned)prm.S;
    const CsmBounds b = bounds[gw];
    const int ncx = (b.max_x - b.min_x + prm.step) / prm.step;
    const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;
    const int ncand = ncx * ncy;
    const int* c = coarse + (size_t)pi * per_pair + (size_t)s * prm.maxc;
    const unsigned long long incumbent = best[pi];
    const int P = pairs[pi].n_pts;
    unsigned ballots[8];                       // maxc <= 16 x 16 = 256 slots
    unsigned n_s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int slot = 32 * k + (int)lane;
      bool keep = false;
      if (32 * k < ncand && slot < ncand) {
        const int ix = slot / ncy;
        const int xo = b.min_x + ix * prm.step, yo = b.min_y + (slot - ix * ncy) * prm.step;
        keep = key_of(score_of(c[slot], P, prm), rank_of(prm, (int)s, xo, yo)) > incumbent;
      }
      ballots[k] = __ballot_sync(0xffffffffu, keep);
      n_s += __popc(ballots[k]);
    }
    if (n_s == 0) continue;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(n_survivors + pi, n_s);
    base = __shfl_sync(0xffffffffu, base, 0);
    unsigned* out = survivors + (size_t)pi * per_pair + base;
    unsigned before = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ballots[k] & (1u << lane))
        out[before + __popc(ballots[k] & ((1u << lane) - 1u))] = s * (unsigned)prm.maxc + 32u * k + lane;
      before += __popc(ballots[k]);
    }
  }
}

// ------------------------------------------------------------------- K7 expand

constexpr int kExpChunk = 2048;   // points staged per pass of the expand kernel

// Survivors of grids without bit planes go to the depth-first refinement unexpanded.
__global__ void csm_survivors_to_nodes_kernel(const CsmPairDev* __restrict__ pairs, CsmParams prm,
                                              const CsmBounds* __restrict__ bounds,
                                              const int* __restrict__ coarse,
                                              const unsigned* __restrict__ survivors,
                                              const unsigned* __restrict__ n_survivors,
                                              CsmNode* __restrict__ nodes, unsigned* __restrict__ n_nodes,
                                              unsigned node_cap) {
  const int pi = blockIdx.y;
  const unsigned ns = n_survivors[pi];
  const size_t per_pair = (size_t)prm.S * prm.maxc;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
    const unsigned v = survivors[(size_t)pi * per_pair + i];
    const int s = (int)(v / prm.maxc), slot = (int)(v % prm.maxc);
    const CsmBounds b = bounds[(size_t)pi * prm.S + s];
    const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;
    CsmNode nd;
    nd.pi = pi; nd.s = s;
    nd.xo = b.min_x + (slot / ncy) * prm.step;
    nd.yo = b.min_y + (slot % ncy) * prm.step;
    nd.d = prm.depth - 1;
    nd.score = score_of(coarse[(size_t)pi * per_pair + v], pairs[pi].n_pts, prm);
    const unsigned pos = atomicAdd(n_nodes, 1u);
    if (pos < node_cap) nodes[pos] = nd;
  }
}

// First level below the coarse lattice, where nearly every surviving coarse candidate dies
// (its four children's bounds fall under the incumbent).  Thread = one survivor, a CTA works
// on survivors of ONE pair: the pair's points (after the initial-yaw rotation) and the
// bit-packed next finer level of its (binary) grid are staged in shared memory; every lane
// rotates the points by its own scan angle, discretises (float fast path with the exact double
// fallback, see the bit-sliced scorer) and tests its four children's cells.  Children that
// can still beat the incumbent become nodes for the depth-first refinement; at depth 2 they
// are leaves.
__global__ void __launch_bounds__(256)
csm_expand_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                  const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                  const CsmBounds* __restrict__ bounds, const int* __restrict__ coarse,
                  const unsigned* __restrict__ survivors, const unsigned* __restrict__ n_survivors,
                  unsigned long long* __restrict__ best, CsmNode* __restrict__ nodes,
                  unsigned* __restrict__ n_nodes, unsigned node_cap,
                  unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char csm_smem[];
  __shared__ int s_sum[32][4];
  const int pi = blockIdx.y, tid = threadIdx.x, lane = tid & 31, slice = tid >> 5;
  const unsigned ns = n_survivors[pi];
  if ((unsigned)blockIdx.x * 32u >= ns) return;
  const size_t per_pair = (size_t)prm.S * prm.maxc;
  const CsmPairDev pr = pairs[pi];
  const CsmGridDev g = grids[pr.grid];
  const int P = pr.n_pts;
  const float* sp = pts + 3 * (size_t)pr.pt_begin;
  const int d = prm.depth - 2;                   // level of the children
  const int h = 1 << d, wm1 = h - 1;
  const int wide_nx = g.nx + wm1, wide_ny = g.ny + wm1, stride = g.lvs[d];
  const unsigned* lvb = g.lvl[d];
  float2* P0 = reinterpret_cast<float2*>(csm_smem);                       // [kExpChunk]
  // rows 16 k apart (neighbouring coarse candidates of one scan) would share two banks with a
  // plain row stride: one extra word every 16 rows spreads them over all banks
  // one zero row after the level and at least one zero word after every row: a lookup clamps an
  // out-of-level coordinate onto them with one unsigned minimum, no test, no mask
  unsigned* bits = reinterpret_cast<unsigned*>(P0 + kExpChunk + 4);      // [wide_ny + 1][stride] (+ row / 16)
  for (int i = tid; i < (wide_ny + 1) * stride; i += 256) {
    const int row = i / stride;
    bits[i + (row >> 4)] = row < wide_ny ? __ldg(lvb + i) : 0u;
  }
  const unsigned wnx_u = (unsigned)wide_nx, wny_u = (unsigned)wide_ny;
  const float mx_f = (float)g.max_x, my_f = (float)g.max_y, ir = (float)(1.0 / g.resolution);
  const float U = (float)(max(g.nx, g.ny) + 2 * prm.n_lin + 64 + 2 * prm.step);
  const float delta = 1.1920929e-07f * (fmaxf(fabsf(mx_f), fabsf(my_f)) * ir + 3.f * U);
  const float hi1 = 1.f - delta;
  const float far_pt = -4.f * (fabsf(mx_f) + fabsf(my_f) + fabsf(pr.tx) + fabsf(pr.ty) + U * (float)g.resolution) - 1000.f;
  unsigned long long expanded = 0;
  const unsigned first_base = (unsigned)blockIdx.x * 32u;

  // 32 survivors per pass (one per lane); the 8 warps split the points of every pass, so that
  // the serial chain of one thread stays short (the kernel's latency is what matters: there
  // are only a few hundred survivors per pair)
  for (unsigned base = first_base; base < ns; base += gridDim.x * 32u) {
    const unsigned i = base + lane;
    bool active = i < ns;
    int s = 0, xo = 0, yo = 0;
    CsmBounds b = {0, 0, 0, 0};
    if (active) {
      const unsigned v = survivors[(size_t)pi * per_pair + i];
      s = (int)(v / prm.maxc);
      const int slot = (int)(v % prm.maxc);
      b = bounds[(size_t)pi * prm.S + s];
      const int ncy = (b.max_y - b.min_y + prm.step) / prm.step;
      xo = b.min_x + (slot / ncy) * prm.step;
      yo = b.min_y + (slot % ncy) * prm.step;
      // the incumbent may have improved since the filter ran
      const float sc = score_of(coarse[(size_t)pi * per_pair + v], P, prm);
      active = key_of(sc, rank_of(prm, s, xo, yo)) > ld_best(best + pi);
    }
    const float2 r = rot[s];
    const int ax = xo + wm1, ay = yo + wm1;   // cell -> level frame of child (0, 0)
    int sum[4] = {0, 0, 0, 0};
    if (tid < 128) s_sum[tid >> 2][tid & 3] = 0;
    for (int p0 = 0; p0 < P; p0 += kExpChunk) {
      const int n = min(kExpChunk, P - p0);
      __syncthreads();
      if (base == first_base || P > kExpChunk) {
        for (int p = tid; p < n; p += 256) {
          float x0, y0;
          rot_z(pr.w0, pr.z0, sp[3 * (size_t)(p0 + p)], sp[3 * (size_t)(p0 + p) + 1], x0, y0);
          P0[p] = make_float2(x0, y0);
        }
        // the last group of four is padded with a point that no rotation and no offset of the window can
        // bring onto the level: at least one of its coordinates stays `far_pt` - |t| away from the origin
        if (tid < 4) P0[n + tid] = make_float2(far_pt, far_pt);
      }
      __syncthreads();
      if (active) {
        // four points per step: all float discretisations first (independent chains, no
        // branch between them), one rarely taken exact fallback, then the cell tests
        for (int p = 4 * slice; p < n; p += 32) {
          int cxs[4], cys[4];
          float wxs[4], wys[4];
          unsigned need = 0u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 q = P0[p + u];
            float x1, y1;
            rot_z(r.x, r.y, q.x, q.y, x1, y1);
            wxs[u] = __fadd_rn(x1, pr.tx);
            wys[u] = __fadd_rn(y1, pr.ty);
            const float uy = (my_f - wys[u]) * ir, ux = (mx_f - wxs[u]) * ir;
            const float fy = floorf(uy), fx = floorf(ux);
            const float dy = uy - fy, dx = ux - fx;
            cxs[u] = (int)fy;
            cys[u] = (int)fx;
            // (a coordinate beyond U cells is outside every child's level: its lookups clamp onto the zero
            // row / word whatever the float cell is, and |u| >= 2^23 or NaN fails the test by itself)
            if (!(fminf(dy, dx) > delta && fmaxf(dy, dx) < hi1)) need |= 1u << u;
          }
          if (need) {   // rare: near a rounding boundary
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (need & (1u << u)) {
                const int2 ce = cells_exact(wxs[u], wys[u], g.resolution, g.max_x, g.max_y);
                cxs[u] = ce.x;
                cys[u] = ce.y;
              }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // children (x, y), (x, y + h), (x + h, y), (x + h, y + h): coordinates outside the level
            // (negative ones wrap to huge unsigned values) land on the zero row / zero word
            const unsigned lx = (unsigned)(cxs[u] + ax), ly = (unsigned)(cys[u] + ay);
            const unsigned cx0 = min(lx, wnx_u), cx1 = min(lx + (unsigned)h, wnx_u);
            const unsigned ry0 = min(ly, wny_u), ry1 = min(ly + (unsigned)h, wny_u);
            const unsigned* r0 = bits + ry0 * stride + (ry0 >> 4);
            const unsigned* r1 = bits + ry1 * stride + (ry1 >> 4);
            const unsigned w0 = cx0 >> 5, w1 = cx1 >> 5, b0 = cx0 & 31u, b1 = cx1 & 31u;
            sum[0] += (int)((r0[w0] >> b0) & 1u);
            sum[1] += (int)((r1[w0] >> b0) & 1u);
            sum[2] += (int)((r0[w1] >> b1) & 1u);
            sum[3] += (int)((r1[w1] >> b1) & 1u);
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) atomicAdd(&s_sum[lane][ch], sum[ch]);
    }
    __syncthreads();
    if (active && slice == 0) {
      ++expanded;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {   // x outer, y inner (fast_..._2d.cpp:414-429)
        const int cx = xo + (ch >> 1) * h, cy = yo + (ch & 1) * h;
        if (cx > b.max_x || cy > b.max_y) continue;
        const float sc = score_of(255 * s_sum[lane][ch], P, prm);
        const unsigned long long key = key_of(sc, rank_of(prm, s, cx, cy));
        if (!(key > ld_best(best + pi))) continue;
        if (d == 0) {
          atomicMax(best + pi, key);
        } else {
          const unsigned pos = atomicAdd(n_nodes, 1u);
          if (pos < node_cap) {
            CsmNode nd;
            nd.pi = pi; nd.s = s; nd.xo = cx; nd.yo = cy; nd.d = d; nd.score = sc;
            nodes[pos] = nd;
          }
        }
      }
    }
    __syncthreads();   // s_sum is reset by the next pass
  }
  if (expanded) atomicAdd(counters, expanded);
}

struct Node {
  int xo, yo, d;
  float score;
};

constexpr int kRefineThreads = 128;

// Persistent CTAs: each pops a node and runs the reference's DFS (children sorted by score,
// fast_..._2d.cpp:430-436) inside its subtree, pruning against the pair's incumbent shared
// through global memory.  The threads of a CTA split the points of every expansion (a node's
// expansions are sequential, so the chain of dependent lookups per expansion is kept short);
// they all hold the same stack.
__global__ void __launch_bounds__(kRefineThreads)
csm_refine_kernel(const CsmGridDev* __restrict__ grids, const CsmPairDev* __restrict__ pairs,
                  const float* __restrict__ pts, const float2* __restrict__ rot, CsmParams prm,
                  const CsmBounds* __restrict__ bounds, const CsmNode* __restrict__ nodes,
                  const unsigned* __restrict__ n_nodes, unsigned node_cap,
                  unsigned* __restrict__ cursor, unsigned long long* __restrict__ best,
                  unsigned long long* __restrict__ counters) {
  __shared__ unsigned s_next;
  __shared__ int s_sum[2][kRefineThreads / 32][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned total = min(*n_nodes, node_cap);
  unsigned long long expanded = 0;
  int round = 0;
  Node stack[3 * kCsmMaxDepth + 4];
  for (;;) {
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(cursor, 1u);
    __syncthreads();
    const unsigned i = s_next;
    if (i >= total) break;
    const CsmNode root = nodes[i];
    const int s = root.s, pi = root.pi;
    const CsmPairDev pr = pairs[pi];
    const CsmGridDev g = grids[pr.grid];
    const CsmBounds b = bounds[(size_t)pi * prm.S + s];
    const float2 r = rot[s];
    const int P = pr.n_pts;
    const float* sp = pts + 3 * (size_t)pr.pt_begin;
    int top = 0;
    stack[0].xo = root.xo;
    stack[0].yo = root.yo;
    stack[0].d = root.d;
    stack[0].score = root.score;
    top = 1;
    while (top > 0) {
      const Node nd = stack[--top];
      const unsigned long long nk = key_of(nd.score, rank_of(prm, s, nd.xo, nd.yo));
      // one thread reads the incumbent so that the whole CTA takes the same decision
      if (tid == 0) s_next = nk > ld_best(best + pi) ? 1u : 0u;
      __syncthreads();
      const bool go = s_next != 0u;
      __syncthreads();
      if (!go) continue;                         // bound cannot beat the incumbent
      if (nd.d == 0) {                           // only when depth == 1
        if (tid == 0) atomicMax(best + pi, nk);
        continue;
      }
      const int h = 1 << (nd.d - 1);
      const LevelView lv = level_view(g, nd.d - 1);
      int sum[4] = {0, 0, 0, 0};
#pragma unroll 2
      for (int p = tid; p < P; p += kRefineThreads) {
        const int2 c = discretize_point(sp + 3 * (size_t)p, pr.w0, pr.z0, r.x, r.y, pr.tx, pr.ty,
                                        g.resolution, g.max_x, g.max_y);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          sum[ch] += level_val(lv, c.x + nd.xo + (ch >> 1) * h, c.y + nd.yo + (ch & 1) * h);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) sum[ch] = warp_sum(sum[ch]);
      if (lane == 0)
        for (int ch = 0; ch < 4; ++ch) s_sum[round & 1][warp][ch] = sum[ch];
      __syncthreads();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        sum[ch] = 0;
        for (int w2 = 0; w2 < kRefineThreads / 32; ++w2) sum[ch] += s_sum[round & 1][w2][ch];
      }
      ++round;
      ++expanded;
      Node kids[4];
      int nk_n = 0;
      for (int ch = 0; ch < 4; ++ch) {  // x outer, y inner (fast_..._2d.cpp:414-429)
        const int cx = nd.xo + (ch >> 1) * h, cy = nd.yo + (ch & 1) * h;
        if (cx > b.max_x || cy > b.max_y) continue;
        kids[nk_n].xo = cx;
        kids[nk_n].yo = cy;
        kids[nk_n].d = nd.d - 1;
        kids[nk_n].score = score_of(sum[ch], P, prm);
        ++nk_n;
      }
      if (nd.d - 1 == 0) {
        unsigned long long bk = 0;
        for (int c2 = 0; c2 < nk_n; ++c2)
          bk = max(bk, key_of(kids[c2].score, rank_of(prm, s, kids[c2].xo, kids[c2].yo)));
        if (tid == 0 && bk > ld_best(best + pi)) atomicMax(best + pi, bk);
      } else {
        // stable insertion sort, descending score; push worst first so best pops first
        for (int a = 1; a < nk_n; ++a) {
          const Node t = kids[a];
          int c2 = a - 1;
          while (c2 >= 0 && kids[c2].score < t.score) {
            kids[c2 + 1] = kids[c2];
            --c2;
          }
          kids[c2 + 1] = t;
        }
        for (int c2 = nk_n - 1; c2 >= 0; --c2) stack[top++] = kids[c2];
      }
    }
  }
  if (tid == 0 && expanded) atomicAdd(counters, expanded);
}

}  // namespace

// ---------------------------------------------------------------------------

cudaError_t launch_csm_level1_from_cells(const uint16_t* cells, const uint8_t* lut, size_t n,
                                         uint8_t* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  csm_level1_from_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cells, lut, n, out);
  return cudaGetLastError();
}

cudaError_t launch_csm_pack_bits(const uint8_t* level1, int nx, int ny, unsigned* out, int* not_binary,
                                 cudaStream_t stream) {
  const int stride = csm_bit_stride(nx), n = ny * stride;
  csm_pack_bits_kernel<<<(n + 255) / 256, 256, 0, stream>>>(level1, nx, ny, stride, out, not_binary);
  return cudaGetLastError();
}

cudaError_t launch_csm_unpack_bits(const unsigned* bits, int nx, int ny, uint8_t* out, cudaStream_t stream) {
  const size_t n = (size_t)nx * ny;
  csm_unpack_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bits, nx, ny, csm_bit_stride(nx), out);
  return cudaGetLastError();
}

cudaError_t launch_csm_assign_slots(CsmPairDev* pairs, int n_pairs, int* hash_keys, int* hash_slot,
                                    int hash_size, int* slot_gid, int* n_slots, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(hash_keys, 0xFF, (size_t)hash_size * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(n_slots, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  const int blocks = (n_pairs + 255) / 256;
  csm_slot_insert_kernel<<<blocks, 256, 0, stream>>>(pairs, n_pairs, hash_keys, hash_slot, hash_size - 1,
                                                     slot_gid, n_slots);
  csm_slot_lookup_kernel<<<blocks, 256, 0, stream>>>(pairs, n_pairs, hash_keys, hash_slot, hash_size - 1);
  return cudaGetLastError();
}

cudaError_t launch_csm_prepare_slots(const CsmGridRec* recs, const int* slot_gid, const int* n_slots,
                                     int max_slots, CsmPlan plan, unsigned char* ws, CsmGridDev* out,
                                     cudaStream_t stream) {
  csm_prepare_slots_kernel<<<(max_slots + 127) / 128, 128, 0, stream>>>(recs, slot_gid, n_slots, plan, ws, out);
  return cudaGetLastError();
}

cudaError_t launch_csm_build_slots(const CsmGridRec* recs, const int* slot_gid, const CsmGridDev* slots,
                                   const int* n_slots, int max_slots, CsmPlan plan, cudaStream_t stream,
                                   uint64_t* launches) {
  // grid.y = slot (early exit beyond the device-side count); grid.x sized for the largest grid
  auto blocks = [](size_t elems, int per_block) {
    return (unsigned)std::min<size_t>((elems + per_block - 1) / per_block, 1u << 20);
  };
  const int top = plan.depth - 1;
  if (plan.bits) {
    // two buffers of the largest level in shared memory: the whole slot in one kernel
    const size_t buf_words = (size_t)(plan.max_ny + (1 << top) - 1) * (size_t)csm_bit_stride(plan.max_nx + (1 << top) - 1);
    if (2 * buf_words * 4 <= (size_t)200 * 1024 && std::getenv("GLOC_CSM_NO_FUSED_BUILD") == nullptr) {
      static unsigned long long attr_mask = 0;
      if (first_use_on_current_device(attr_mask)) {
        cudaError_t e = cudaFuncSetAttribute(csm_build_slot_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024);
        if (e != cudaSuccess) return e;
      }
      csm_build_slot_fused_kernel<<<max_slots, 512, 2 * buf_words * 4, stream>>>(slots, n_slots, plan.depth, (int)buf_words);
      ++*launches;
      return cudaGetLastError();
    }
    for (int l = 1; l < plan.depth; ++l) {
      const int w = 1 << l;
      const size_t words = (size_t)(plan.max_ny + w - 1) * (size_t)csm_bit_stride(plan.max_nx + w - 1);
      csm_build_lvl_bits_kernel<<<dim3(blocks(words, 256), max_slots), 256, 0, stream>>>(slots, n_slots, l);
      ++*launches;
    }
    const int w = 1 << top;
    const size_t words = (size_t)w * w * (size_t)csm_pmb_rows(plan.max_ny + w - 1, plan.n_lin, top, plan.pmb_b1 >= 0);
    csm_build_pmb_kernel<<<dim3(blocks(words, 128), max_slots), 128, 0, stream>>>(slots, n_slots);
    ++*launches;
  } else {
    const size_t n0 = (size_t)plan.max_nx * plan.max_ny;
    csm_slot_level0_u8_kernel<<<dim3(blocks(n0, 1024), max_slots), 256, 0, stream>>>(recs, slot_gid, slots, n_slots);
    ++*launches;
    for (int l = 1; l < plan.depth; ++l) {
      const int w = 1 << l;
      const size_t n = (size_t)(plan.max_nx + w - 1) * (size_t)(plan.max_ny + w - 1);
      csm_build_level_kernel<<<dim3(blocks(n, 1024), max_slots), 256, 0, stream>>>(slots, n_slots, l);
      ++*launches;
    }
    if (plan.use_pm) {
      const int w = 1 << top;
      const size_t n = (size_t)(plan.max_nx + 2 * w + 2 * plan.n_lin) * (size_t)(plan.max_ny + 2 * w + 2 * plan.n_lin);
      csm_build_pm_kernel<<<dim3(blocks(n, 1024), max_slots), 256, 0, stream>>>(slots, n_slots);
      ++*launches;
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_csm_discretize(const float* pts, int n_pts, float w0, float z0, float tx,
                                  float ty, const float2* rot, int S, double resolution,
                                  double max_x, double max_y, int* out_cells, cudaStream_t stream) {
  if (n_pts == 0 || S == 0) return cudaSuccess;
  dim3 grd((n_pts + 255) / 256, S);
  csm_discretize_kernel<<<grd, 256, 0, stream>>>(pts, n_pts, w0, z0, tx, ty, rot, S, resolution,
                                                 max_x, max_y, out_cells);
  return cudaGetLastError();
}

cudaError_t launch_csm_coarse(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                              const float* pts, const float2* rot, CsmParams prm,
                              CsmBounds* bounds, int* coarse, unsigned long long* top_coarse,
                              cudaStream_t stream, bool phase_major) {
  dim3 grd(prm.S, n_pairs);
  if (phase_major) {
    const size_t smem = (size_t)kPointChunk * (8 + 4 + 8) + (size_t)prm.maxc * 4;
    static unsigned long long attr_mask = 0;
    if (first_use_on_current_device(attr_mask)) {
      cudaError_t e = cudaFuncSetAttribute(csm_coarse_pm_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) return e;
    }
    csm_coarse_pm_kernel<<<grd, kPmThreads, smem, stream>>>(grids, pairs, pts, rot, prm, bounds,
                                                            coarse, top_coarse);
  } else {
    csm_coarse_kernel<<<grd, kCoarseThreads, 0, stream>>>(grids, pairs, pts, rot, prm, bounds,
                                                          coarse, top_coarse);
  }
  return cudaGetLastError();
}

int csm_pmb_rows(int wide_ny, int n_lin, int log2w, bool paired) { return csm_pmb_rows_of(wide_ny, n_lin, log2w, paired); }

size_t csm_coarse_bits_smem(int log2w, int rows) {
  return ((size_t)(1 << (2 * log2w)) * rows + kBitChunk + 16) * 8;   // planes + one chunk of points + padding points
}

namespace {
template <int NP, bool PR>
cudaError_t launch_bits_np(dim3 grd, int threads, size_t smem, cudaStream_t stream,
                           const CsmGridDev* grids, const CsmPairDev* pairs, const float* pts,
                           const float2* rot, CsmParams prm, CsmBounds* bounds, int* coarse,
                           unsigned long long* top_coarse) {
  static unsigned long long attr_mask = 0;
  if (first_use_on_current_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(csm_coarse_bits_kernel<NP, PR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  csm_coarse_bits_kernel<NP, PR><<<grd, threads, smem, stream>>>(grids, pairs, pts, rot, prm, bounds,
                                                             coarse, top_coarse);
  return cudaGetLastError();
}
}  // namespace

// All grids of the batch must carry bit planes built for prm.depth-1 / prm.n_lin and share
// the largest shared-memory footprint `smem`.
cudaError_t launch_csm_coarse_bits(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                                   const float* pts, const float2* rot, CsmParams prm,
                                   CsmBounds* bounds, int* coarse, unsigned long long* top_coarse,
                                   size_t smem, int warps, bool paired, cudaStream_t stream) {
  const int threads = 32 * warps;            // two lanes per rotation
  dim3 grd((2 * prm.S + threads - 1) / threads, n_pairs);
  // packed words (two rows each) per lane: 4 NP candidate rows per rotation, 4 NP - 2 when paired
  const int np = paired ? (prm.max_side + 5) / 4 : (prm.max_side + 3) / 4;
#define GLOC_BITS_CASE(N)                                                                             \
  case N:                                                                                             \
    return paired ? launch_bits_np<N, true>(grd, threads, smem, stream, grids, pairs, pts, rot, prm,  \
                                            bounds, coarse, top_coarse)                               \
                  : launch_bits_np<N, false>(grd, threads, smem, stream, grids, pairs, pts, rot, prm, \
                                             bounds, coarse, top_coarse);
  switch (np) {
    GLOC_BITS_CASE(1) GLOC_BITS_CASE(2) GLOC_BITS_CASE(3) GLOC_BITS_CASE(4)
    default: return cudaErrorInvalidValue;
  }
#undef GLOC_BITS_CASE
}

cudaError_t launch_csm_seed(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                            const float* pts, const float2* rot, CsmParams prm,
                            const CsmBounds* bounds, const unsigned long long* top_coarse,
                            unsigned long long* best, cudaStream_t stream) {
  csm_seed_kernel<<<n_pairs, 256, 0, stream>>>(grids, pairs, pts, rot, prm, bounds, top_coarse,
                                               best);
  return cudaGetLastError();
}

cudaError_t launch_csm_filter(const CsmPairDev* pairs, int n_pairs, CsmParams prm,
                              const CsmBounds* bounds, const int* coarse,
                              const unsigned long long* best, unsigned* survivors,
                              unsigned* n_survivors, cudaStream_t stream) {
  const size_t warps = (size_t)n_pairs * prm.S;
  const unsigned blocks = (unsigned)std::min<size_t>((warps + 7) / 8, 148 * 32);
  csm_filter_kernel<<<blocks, 256, 0, stream>>>(pairs, n_pairs, prm, bounds, coarse, best,
                                                survivors, n_survivors);
  return cudaGetLastError();
}

size_t csm_expand_smem(int wide_ny, int stride) {
  return (size_t)(kExpChunk + 4) * sizeof(float2) +
         ((size_t)(wide_ny + 1) * stride + (size_t)((wide_ny + 1) >> 4) + 1) * 4;
}

// bits == true: every grid of the batch carries bit-packed levels (binary slots), and
// `smem` is the largest footprint; else the survivors become nodes unexpanded.
cudaError_t launch_csm_expand(const CsmGridDev* grids, const CsmPairDev* pairs, int n_pairs,
                              const float* pts, const float2* rot, CsmParams prm,
                              const CsmBounds* bounds, const int* coarse,
                              const unsigned* survivors, const unsigned* n_survivors,
                              unsigned long long* best, CsmNode* nodes, unsigned* n_nodes,
                              unsigned node_cap, unsigned long long* counters, int chunks,
                              bool bits, size_t smem, cudaStream_t stream) {
  dim3 grd(chunks, n_pairs);
  if (!bits) {
    csm_survivors_to_nodes_kernel<<<grd, 256, 0, stream>>>(pairs, prm, bounds, coarse, survivors,
                                                           n_survivors, nodes, n_nodes, node_cap);
    return cudaGetLastError();
  }
  static unsigned long long attr_mask = 0;
  if (first_use_on_current_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(csm_expand_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  csm_expand_kernel<<<grd, 256, smem, stream>>>(grids, pairs, pts, rot, prm, bounds, coarse, survivors,
                                                n_survivors, best, nodes, n_nodes, node_cap, counters);
  return cudaGetLastError();
}

cudaError_t launch_csm_refine(const CsmGridDev* grids, const CsmPairDev* pairs, const float* pts,
                              const float2* rot, CsmParams prm, const CsmBounds* bounds,
                              const CsmNode* nodes, const unsigned* n_nodes, unsigned node_cap,
                              unsigned* cursor, unsigned long long* best,
                              unsigned long long* counters, int n_ctas, cudaStream_t stream) {
  csm_refine_kernel<<<n_ctas, kRefineThreads, 0, stream>>>(grids, pairs, pts, rot, prm, bounds, nodes,
                                                           n_nodes, node_cap, cursor, best, counters);
  return cudaGetLastError();
}

}  // namespace gloc
