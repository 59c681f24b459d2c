// csm_api.cu -- C ABI of stage 2 (scan-match verification).  See include/gloc3d.h
// for the reference interfaces each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "csm_kernels.cuh"

using namespace gloc;

// bev.cu
const uint8_t* gloc_bev_device_image(const gloc_bev_projector* b, gloc_bev_info* info);
int gloc_bev_device_of(const gloc_bev_projector* b);
cudaError_t gloc_bev_launch_level1(const uint8_t* img, size_t n, uint8_t* out, cudaStream_t s);
cudaError_t gloc_bev_launch_level1_aligned(const uint8_t* img, int W, int H, uint8_t* out, cudaStream_t s);

namespace {

// /root/reference/registration/3d/probability_values.h:64-67, float32 on purpose.
const float kMinProbability = 0.1f;
const float kMaxProbability = 1.f - kMinProbability;
const float kMinCorrespondenceCost = 1.f - kMaxProbability;
const float kMaxCorrespondenceCost = 1.f - kMinProbability;

// ValueToCorrespondenceCost table, 3d/probability_values.cpp:27-36,38-52,59-63
float value_to_cost(uint16_t value) {
  const uint16_t v = value & 32767u;  // the table repeats for update-marked values
  if (v == 0) return kMaxCorrespondenceCost;
  const float kScale = (kMaxCorrespondenceCost - kMinCorrespondenceCost) / 32766.f;
  return v * kScale + (kMinCorrespondenceCost - kScale);
}

struct HostGrid {
  int nx = 0, ny = 0;
  double resolution = 0, max_x = 0, max_y = 0;
  uint8_t* d_stack = nullptr;  // levels 0..depth-1 concatenated; level 0 is the width-1 grid
  int depth = 0;               // levels currently built
  size_t bytes = 0;
  long long off[kCsmMaxDepth] = {0};
  // padded phase-major copy of the coarsest level used by the last batches (CsmGridDev::pm)
  uint8_t* d_pm = nullptr;
  int pm_level = -1, pm_pad = 0, pm_pw = 0, pm_ph = 0;
  // bit planes of the coarsest level (binary grids; CsmGridDev::pmb)
  unsigned long long* d_pmb = nullptr;
  int pmb_level = -1, pmb_nlin = -1, pmb_rows = 0;
  int binary = -1;   // -1 unknown, 0 some cell is neither 0 nor 255, 1 binary
  unsigned* d_lvb = nullptr;   // bit-packed level depth-2 (CsmGridDev::lvb)
  int lvb_level = -1, lvb_stride = 0;
};

size_t stack_bytes(int nx, int ny, int depth, long long* off) {
  size_t total = 0;
  for (int i = 0; i < depth; ++i) {
    const int w = 1 << i;
    if (off) off[i] = (long long)total;
    total += (size_t)(nx + w - 1) * (size_t)(ny + w - 1);
    total = (total + 15) & ~(size_t)15;
  }
  return total;
}

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need + need / 4 + 256);
    if (e == cudaSuccess) bytes = need + need / 4 + 256; else p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

struct gloc_csm_store {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<HostGrid> grids;
  uint8_t* d_lut = nullptr;  // uint16 cost value -> uint8 width-1 cell
  Buf pts, pairs, gridtab, rot, bounds, coarse, top, best, survivors, nsurv, nodes, misc, cells16, disc;
  gloc_csm_stats stats{};
  EventProfiler prof;
};

namespace {

// Build the uint16 -> uint8 table exactly as PrecomputationGrid2D does per cell:
// ComputeCellValue(1.f - |cost|), fast_..._2d.cpp:118-119,130-131,184-190.
void build_lut(std::vector<uint8_t>& lut) {
  lut.resize(65536);
  const float min_score = 1.f - kMaxCorrespondenceCost;
  const float max_score = 1.f - kMinCorrespondenceCost;
  for (int v = 0; v < 65536; ++v) {
    const float p = 1.f - std::fabs(value_to_cost((uint16_t)v));
    const long cell = std::lround((p - min_score) * (255.f / (max_score - min_score)));
    lut[v] = (uint8_t)std::min(255l, std::max(0l, cell));
  }
}

int ensure_stack(gloc_csm_store* st, HostGrid& g, int depth) {
  if (g.depth >= depth) return GLOC_OK;
  long long off[kCsmMaxDepth] = {0};
  const size_t bytes = stack_bytes(g.nx, g.ny, depth, off);
  uint8_t* ns = nullptr;
  GLOC_CUDA_TRY(cudaMalloc((void**)&ns, bytes));
  const size_t l1 = (size_t)g.nx * g.ny;
  cudaError_t e = cudaMemcpyAsync(ns, g.d_stack, l1, cudaMemcpyDeviceToDevice, st->stream);
  for (int i = 1; i < depth && e == cudaSuccess; ++i) {
    e = launch_csm_build_level(ns + off[i - 1], g.nx, g.ny, 1 << i, ns + off[i], st->stream);
    st->stats.kernel_launches++;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  if (e != cudaSuccess) {
    cudaFree(ns);
    return fail(GLOC_ERR_CUDA, std::string("precomputation stack: ") + cudaGetErrorString(e));
  }
  cudaFree(g.d_stack);
  g.d_stack = ns;
  g.depth = depth;
  g.bytes = bytes;
  std::memcpy(g.off, off, sizeof(off));
  return GLOC_OK;
}

int add_grid_common(gloc_csm_store* st, int nx, int ny, double resolution, double max_x,
                    double max_y, uint8_t** d_level1) {
  if (nx < 1 || ny < 1 || !(resolution > 0.))  // MapLimits ctor CHECKs, map_limits.h:44-46
    return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid: bad limits");
  if ((long long)nx * ny > (1ll << 30)) return fail(GLOC_ERR_RANGE, "gloc_csm_add_grid: grid too large");
  (void)max_x;
  (void)max_y;
  GLOC_CUDA_TRY(cudaMalloc((void**)d_level1, (size_t)nx * ny));
  (void)st;
  return GLOC_OK;
}

}  // namespace

extern "C" {

int gloc_csm_create(gloc_csm_store** out, int device) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_csm_create: out is null");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_csm_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(GLOC_ERR_INVALID, "gloc_csm_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return fail(GLOC_ERR_CUDA, "gloc_csm_create: device is not sm_100 (kernels are sm_100a only)");
  DeviceGuard g(device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_csm_create: cudaSetDevice failed");
  gloc_csm_store* st = new (std::nothrow) gloc_csm_store;
  if (!st) return fail(GLOC_ERR_NOMEM, "gloc_csm_create: out of host memory");
  st->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_lut, 65536);
  if (e == cudaSuccess) {
    std::vector<uint8_t> lut;
    build_lut(lut);
    e = cudaMemcpy(st->d_lut, lut.data(), 65536, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    gloc_csm_destroy(st);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_create: ") + cudaGetErrorString(e));
  }
  *out = st;
  return GLOC_OK;
}

void gloc_csm_destroy(gloc_csm_store* st) {
  if (!st) return;
  DeviceGuard g(st->device);
  if (st->stream) {
    cudaStreamSynchronize(st->stream);
    cudaStreamDestroy(st->stream);
  }
  for (auto& gr : st->grids) {
    if (gr.d_stack) cudaFree(gr.d_stack);
    if (gr.d_pm) cudaFree(gr.d_pm);
    if (gr.d_pmb) cudaFree(gr.d_pmb);
    if (gr.d_lvb) cudaFree(gr.d_lvb);
  }
  if (st->d_lut) cudaFree(st->d_lut);
  for (Buf* b : {&st->pts, &st->pairs, &st->gridtab, &st->rot, &st->bounds, &st->coarse, &st->top,
                 &st->best, &st->survivors, &st->nsurv, &st->nodes, &st->misc, &st->cells16, &st->disc})
    b->release();
  delete st;
}

int gloc_csm_add_grid_cells(gloc_csm_store* st, const uint16_t* cells, int nx, int ny,
                            double resolution, double max_x, double max_y, int* grid_id) {
  if (!st || !cells) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_cells: null argument");
  DeviceGuard g(st->device);
  uint8_t* d_l1 = nullptr;
  int rc = add_grid_common(st, nx, ny, resolution, max_x, max_y, &d_l1);
  if (rc != GLOC_OK) return rc;
  const size_t n = (size_t)nx * ny;
  cudaError_t e = st->cells16.reserve(n * 2);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(st->cells16.p, cells, n * 2, cudaMemcpyHostToDevice, st->stream);
  if (e == cudaSuccess)
    e = launch_csm_level1_from_cells((const uint16_t*)st->cells16.p, st->d_lut, n, d_l1, st->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  if (e != cudaSuccess) {
    cudaFree(d_l1);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_add_grid_cells: ") + cudaGetErrorString(e));
  }
  st->stats.kernel_launches++;
  HostGrid hg;
  hg.nx = nx; hg.ny = ny; hg.resolution = resolution; hg.max_x = max_x; hg.max_y = max_y;
  hg.d_stack = d_l1; hg.depth = 1; hg.bytes = n;
  st->grids.push_back(hg);
  if (grid_id) *grid_id = (int)st->grids.size() - 1;
  return GLOC_OK;
}

int gloc_csm_add_grid_u8(gloc_csm_store* st, const uint8_t* level1, int nx, int ny,
                         double resolution, double max_x, double max_y, int* grid_id) {
  if (!st || !level1) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_u8: null argument");
  DeviceGuard g(st->device);
  uint8_t* d_l1 = nullptr;
  int rc = add_grid_common(st, nx, ny, resolution, max_x, max_y, &d_l1);
  if (rc != GLOC_OK) return rc;
  cudaError_t e = cudaMemcpy(d_l1, level1, (size_t)nx * ny, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_l1);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_add_grid_u8: ") + cudaGetErrorString(e));
  }
  HostGrid hg;
  hg.nx = nx; hg.ny = ny; hg.resolution = resolution; hg.max_x = max_x; hg.max_y = max_y;
  hg.d_stack = d_l1; hg.depth = 1; hg.bytes = (size_t)nx * ny;
  st->grids.push_back(hg);
  if (grid_id) *grid_id = (int)st->grids.size() - 1;
  return GLOC_OK;
}

int gloc_csm_add_grid_from_bev(gloc_csm_store* st, gloc_bev_projector* bev, int* grid_id) {
  if (!st || !bev) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev: null argument");
  gloc_bev_info I;
  const uint8_t* d_img = gloc_bev_device_image(bev, &I);
  if (!d_img || I.width < 1 || I.height < 1)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_csm_add_grid_from_bev: the projector holds no image");
  if (gloc_bev_device_of(bev) != st->device)
    return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev: projector and store live on different devices");
  DeviceGuard g(st->device);
  // ProjectToGrid, 3d/submap_3d.cpp:376-388: limits from the voxel bounding box
  const double max_x = (I.min_ix + I.width - 1) * I.resolution;
  const double max_y = (I.min_iy + I.height - 1) * I.resolution;
  uint8_t* d_l1 = nullptr;
  int rc = add_grid_common(st, I.width, I.height, I.resolution, max_x, max_y, &d_l1);
  if (rc != GLOC_OK) return rc;
  const size_t n = (size_t)I.width * I.height;
  cudaError_t e = gloc_bev_launch_level1(d_img, n, d_l1, st->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  if (e != cudaSuccess) {
    cudaFree(d_l1);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_add_grid_from_bev: ") + cudaGetErrorString(e));
  }
  st->stats.kernel_launches++;
  HostGrid hg;
  hg.nx = I.width; hg.ny = I.height; hg.resolution = I.resolution; hg.max_x = max_x; hg.max_y = max_y;
  hg.d_stack = d_l1; hg.depth = 1; hg.bytes = n;
  st->grids.push_back(hg);
  if (grid_id) *grid_id = (int)st->grids.size() - 1;
  return GLOC_OK;
}

int gloc_csm_add_grid_from_bev_aligned(gloc_csm_store* st, gloc_bev_projector* bev, int* grid_id) {
  if (!st || !bev) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev_aligned: null argument");
  gloc_bev_info I;
  const uint8_t* d_img = gloc_bev_device_image(bev, &I);
  if (!d_img || I.width < 1 || I.height < 1)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_csm_add_grid_from_bev_aligned: the projector holds no image");
  if (gloc_bev_device_of(bev) != st->device)
    return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev_aligned: projector and store live on different devices");
  DeviceGuard g(st->device);
  // cell x runs along world -y and cell y along world -x (2d/map_limits.h:69-76): the grid has
  // height x width cells and max() half a cell beyond the last voxel centre
  const int nx = I.height, ny = I.width;
  const double max_x = (I.min_ix + I.width - 1 + 0.5) * I.resolution;
  const double max_y = (I.min_iy + I.height - 1 + 0.5) * I.resolution;
  uint8_t* d_l1 = nullptr;
  int rc = add_grid_common(st, nx, ny, I.resolution, max_x, max_y, &d_l1);
  if (rc != GLOC_OK) return rc;
  cudaError_t e = gloc_bev_launch_level1_aligned(d_img, I.width, I.height, d_l1, st->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  if (e != cudaSuccess) {
    cudaFree(d_l1);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_add_grid_from_bev_aligned: ") + cudaGetErrorString(e));
  }
  st->stats.kernel_launches++;
  HostGrid hg;
  hg.nx = nx; hg.ny = ny; hg.resolution = I.resolution; hg.max_x = max_x; hg.max_y = max_y;
  hg.d_stack = d_l1; hg.depth = 1; hg.bytes = (size_t)nx * ny;
  st->grids.push_back(hg);
  if (grid_id) *grid_id = (int)st->grids.size() - 1;
  return GLOC_OK;
}

int gloc_csm_num_grids(const gloc_csm_store* st) { return st ? (int)st->grids.size() : 0; }

int gloc_csm_get_grid_info(const gloc_csm_store* st, int grid_id, gloc_grid_info* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_grid_info: null argument");
  if (grid_id < 0 || grid_id >= (int)st->grids.size())
    return fail(GLOC_ERR_INVALID, "gloc_csm_get_grid_info: bad grid id");
  const HostGrid& hg = st->grids[grid_id];
  out->nx = hg.nx;
  out->ny = hg.ny;
  out->resolution = hg.resolution;
  out->max_x = hg.max_x;
  out->max_y = hg.max_y;
  return GLOC_OK;
}

int gloc_csm_get_precomputation_grid(gloc_csm_store* st, int grid_id, int width, uint8_t* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_precomputation_grid: null argument");
  if (grid_id < 0 || grid_id >= (int)st->grids.size())
    return fail(GLOC_ERR_INVALID, "gloc_csm_get_precomputation_grid: bad grid id");
  int level = 0;
  while ((1 << level) < width) ++level;
  if (width < 1 || (1 << level) != width || level >= kCsmMaxDepth)  // CHECK_GE(width, 1)
    return fail(GLOC_ERR_RANGE, "gloc_csm_get_precomputation_grid: width must be 1,2,4,...,128");
  DeviceGuard g(st->device);
  HostGrid& hg = st->grids[grid_id];
  int rc = ensure_stack(st, hg, level + 1);
  if (rc != GLOC_OK) return rc;
  const size_t n = (size_t)(hg.nx + width - 1) * (size_t)(hg.ny + width - 1);
  GLOC_CUDA_TRY(cudaMemcpy(out, hg.d_stack + hg.off[level], n, cudaMemcpyDeviceToHost));
  return GLOC_OK;
}

int gloc_csm_match_batch(gloc_csm_store* st, const float* pts, const int64_t* scan_offsets,
                         int n_scans, const int* grid_ids, const int* scan_ids,
                         const double* init_xyyaw, int n_pairs, int n_lin, int n_ang,
                         double ang_step, int depth, float min_score, gloc_csm_result* results) {
  if (!st) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: null store");
  if (n_pairs == 0) return GLOC_OK;
  if (!pts || !scan_offsets || !grid_ids || !scan_ids || !init_xyyaw || !results || n_pairs < 0 ||
      n_scans < 1)
    return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: null argument");
  if (depth < 1 || depth > kCsmMaxDepth)  // CHECK_GE(branch_and_bound_depth, 1), fast_..._2d.cpp:195
    return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: depth must be in [1, 8]");
  if (n_lin < 0 || n_ang < 0) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: negative window");
  const long long S = 2ll * n_ang + 1, W = 2ll * n_lin + 1;
  if (S * W * W >= (1ll << 32) || S > 65535)
    return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: search window too large (scans*(2*n_lin+1)^2 must be < 2^32)");
  DeviceGuard guard(st->device);
  if (!guard.ok) return fail(GLOC_ERR_CUDA, "gloc_csm_match_batch: cudaSetDevice failed");

  CsmParams prm;
  prm.n_lin = n_lin;
  prm.n_ang = n_ang;
  prm.S = (int)S;
  prm.depth = depth;
  prm.step = 1 << (depth - 1);
  prm.max_side = (2 * n_lin) / prm.step + 1;
  prm.maxc = prm.max_side * prm.max_side;
  prm.W = (unsigned)W;
  prm.min_score = min_score;
  prm.min_s = 1.f - kMaxCorrespondenceCost;
  prm.coef = ((1.f - kMinCorrespondenceCost) - (1.f - kMaxCorrespondenceCost)) / 255.f;

  // per-angle quaternions from the host libm, theta accumulated in double exactly as
  // GenerateRotatedScans does (correlative_scan_matcher_2d.cpp:99-107)
  std::vector<float2> rot((size_t)S);
  {
    double delta_theta = -n_ang * ang_step;
    for (long long s = 0; s < S; ++s, delta_theta += ang_step) {
      const float ha = 0.5f * (float)delta_theta;
      rot[(size_t)s] = make_float2(std::cos(ha), std::sin(ha));
    }
  }
  const int64_t total_pts = scan_offsets[n_scans];
  if (total_pts < 0) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad scan offsets");

  // validate + device tables
  std::vector<CsmPairDev> hp((size_t)n_pairs);
  bool use_pm = std::getenv("GLOC_CSM_NO_PM") == nullptr;
  // bit-sliced coarse scorer: binary grids whose coarsest level fits 64 columns per plane
  // row and shared memory, at most 16 lattice candidates per axis
  bool use_bits = std::getenv("GLOC_CSM_NO_BITS") == nullptr && prm.max_side <= 16;
  size_t bits_smem = 0, exp_smem = 0;
  bool exp_bits = use_bits && depth >= 2;   // expand stage on bit-packed level depth-2
  std::map<int, int> grid_slot;
  std::vector<CsmGridDev> hg;
  for (int i = 0; i < n_pairs; ++i) {
    const int gi = grid_ids[i], si = scan_ids[i];
    if (gi < 0 || gi >= (int)st->grids.size() || si < 0 || si >= n_scans)
      return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad grid or scan id");
    const int64_t b = scan_offsets[si], e = scan_offsets[si + 1];
    if (b < 0 || e < b || e > total_pts || e - b > INT32_MAX)
      return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad scan offsets");
    if (e == b) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: empty scan");
    auto it = grid_slot.find(gi);
    if (it == grid_slot.end()) {
      HostGrid& g = st->grids[gi];
      int rc = ensure_stack(st, g, depth);
      if (rc != GLOC_OK) return rc;
      CsmGridDev d;
      d.stack = g.d_stack;
      std::memcpy(d.off, g.off, sizeof(d.off));
      d.nx = g.nx; d.ny = g.ny; d.resolution = g.resolution; d.max_x = g.max_x; d.max_y = g.max_y;
      {  // padded phase-major coarsest level (rebuilt when depth or window change)
        const int level = depth - 1, w = 1 << level, pad = n_lin;
        const int wide_nx = g.nx + w - 1, wide_ny = g.ny + w - 1;
        const int pw = (wide_nx + 2 * pad + w - 1) / w, ph = (wide_ny + 2 * pad + w - 1) / w;
        const size_t cells = (size_t)pw * ph * w * w;
        if (cells > ((size_t)1 << 27)) use_pm = false;  // absurd windows: bounds-checked kernel
        if (use_pm && (g.pm_level != level || g.pm_pad != pad)) {
          if (g.d_pm) cudaFree(g.d_pm);
          g.d_pm = nullptr;
          g.pm_level = -1;
          GLOC_CUDA_TRY(cudaMalloc((void**)&g.d_pm, cells));
          GLOC_CUDA_TRY(launch_csm_build_pm(g.d_stack + g.off[level], wide_nx, wide_ny, pad, level,
                                            pw, ph, g.d_pm, st->stream));
          st->stats.kernel_launches++;
          g.pm_level = level; g.pm_pad = pad; g.pm_pw = pw; g.pm_ph = ph;
        }
        d.pm = g.d_pm; d.pm_pad = g.pm_pad; d.pm_pw = g.pm_pw; d.pm_ph = g.pm_ph; d.pm_log2w = level;
      }
      d.pmb = nullptr; d.pmb_rows = 0; d.pmb_px = 0; d.pmb_py = 0; d.pmb_log2w = 0;
      d.lvb = nullptr; d.lvb_stride = 0;
      if (use_bits && g.binary != 0) {
        const int level = depth - 1, w = 1 << level;
        const int wide_nx = g.nx + w - 1, wide_ny = g.ny + w - 1;
        const int rows = csm_pmb_rows(wide_ny, n_lin, level);
        const size_t smem = csm_coarse_bits_smem(level, rows);
        if (((wide_nx + n_lin - 1) >> level) >= 64 || smem > (size_t)225 * 1024) {
          use_bits = false;
        } else {
          if (g.pmb_level != level || g.pmb_nlin != n_lin) {
            if (g.d_pmb) cudaFree(g.d_pmb);
            g.d_pmb = nullptr;
            g.pmb_level = -1;
            GLOC_CUDA_TRY(cudaMalloc((void**)&g.d_pmb, (size_t)w * w * rows * 8));
            GLOC_CUDA_TRY(st->misc.reserve(64));
            GLOC_CUDA_TRY(cudaMemsetAsync(st->misc.p, 0, 4, st->stream));
            GLOC_CUDA_TRY(launch_csm_build_pmb(g.d_stack + g.off[level], wide_nx, wide_ny, n_lin,
                                               2 * n_lin, level, rows, g.d_pmb, (int*)st->misc.p,
                                               st->stream));
            int not_binary = 0;
            GLOC_CUDA_TRY(cudaMemcpyAsync(&not_binary, st->misc.p, 4, cudaMemcpyDeviceToHost, st->stream));
            GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
            st->stats.kernel_launches++;
            g.binary = not_binary ? 0 : 1;
            g.pmb_level = level; g.pmb_nlin = n_lin; g.pmb_rows = rows;
          }
          if (g.binary == 1) {
            d.pmb = g.d_pmb; d.pmb_rows = g.pmb_rows; d.pmb_px = n_lin; d.pmb_py = 2 * n_lin;
            d.pmb_log2w = level;
            bits_smem = std::max(bits_smem, smem);
          } else {
            use_bits = false;
          }
          d.lvb = nullptr; d.lvb_stride = 0;
          if (g.binary == 1 && depth >= 2) {   // the level below, bit-packed, for the expand stage
            const int l2 = depth - 2, w2 = 1 << l2;
            const int wnx = g.nx + w2 - 1, wny = g.ny + w2 - 1, stride = (wnx + 31) / 32 + 1;
            if (csm_expand_smem(wny, stride) > (size_t)112 * 1024) {
              exp_bits = false;
            } else {
              if (g.lvb_level != l2) {
                if (g.d_lvb) cudaFree(g.d_lvb);
                g.d_lvb = nullptr;
                g.lvb_level = -1;
                GLOC_CUDA_TRY(cudaMalloc((void**)&g.d_lvb, (size_t)wny * stride * 4));
                GLOC_CUDA_TRY(launch_csm_build_lvb(g.d_stack + g.off[l2], wnx, wny, stride, g.d_lvb,
                                                   st->stream));
                st->stats.kernel_launches++;
                g.lvb_level = l2; g.lvb_stride = stride;
              }
              d.lvb = g.d_lvb; d.lvb_stride = g.lvb_stride;
              exp_smem = std::max(exp_smem, csm_expand_smem(wny, stride));
            }
          }
        }
      } else {
        use_bits = false;
      }
      if (!use_bits) exp_bits = false;
      it = grid_slot.emplace(gi, (int)hg.size()).first;
      hg.push_back(d);
    }
    CsmPairDev& p = hp[(size_t)i];
    p.grid = it->second;
    p.pt_begin = b;
    p.n_pts = (int)(e - b);
    // initial_rotation.cast<float>().angle() -> Quaternionf(AngleAxisf) (fast_..._2d.cpp:278-283)
    const float ha = 0.5f * (float)init_xyyaw[3 * i + 2];
    p.w0 = std::cos(ha);
    p.z0 = std::sin(ha);
    p.tx = (float)init_xyyaw[3 * i];      // Eigen::Translation2f(double, double), :287-288
    p.ty = (float)init_xyyaw[3 * i + 1];
  }

  cudaStream_t stream = st->stream;
  GLOC_CUDA_TRY(st->pts.reserve((size_t)total_pts * 3 * sizeof(float)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(st->gridtab.reserve(hg.size() * sizeof(CsmGridDev)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->pts.p, pts, (size_t)total_pts * 3 * sizeof(float),
                                cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2),
                                cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->gridtab.p, hg.data(), hg.size() * sizeof(CsmGridDev),
                                cudaMemcpyHostToDevice, stream));

  // sub-batches bound the coarse-score workspace and keep candidate ids in 32 bits
  const long long per_pair = S * prm.maxc;
  long long sub = std::min<long long>(n_pairs, std::max<long long>(1, (1ll << 28) / per_pair));
  sub = std::min<long long>(sub, 65535);
  unsigned long long init_key = 0ull;  // incumbent starts at (min_score, worst rank)
  if (min_score > 0.f) {
    uint32_t u;
    std::memcpy(&u, &min_score, 4);
    init_key = (unsigned long long)u << 32;
  }
  std::vector<unsigned long long> hbest((size_t)n_pairs);
  const int n_ctas = sm_count(st->device) * 8;   // persistent refinement CTAs (128 threads)
  int bits_warps = 12;   // rotations per CTA = 16 x warps (two lanes per rotation)
  if (const char* e = std::getenv("GLOC_CSM_BITS_WARPS")) bits_warps = std::min(12, std::max(1, std::atoi(e)));
  for (long long p0 = 0; p0 < n_pairs; p0 += sub) {
    const int np = (int)std::min<long long>(sub, n_pairs - p0);
    GLOC_CUDA_TRY(st->pairs.reserve((size_t)np * sizeof(CsmPairDev)));
    GLOC_CUDA_TRY(st->bounds.reserve((size_t)np * S * sizeof(CsmBounds)));
    GLOC_CUDA_TRY(st->coarse.reserve((size_t)np * per_pair * sizeof(int)));
    GLOC_CUDA_TRY(st->survivors.reserve((size_t)np * per_pair * sizeof(unsigned)));
    GLOC_CUDA_TRY(st->top.reserve((size_t)np * 8));
    GLOC_CUDA_TRY(st->best.reserve((size_t)np * 8));
    GLOC_CUDA_TRY(st->nsurv.reserve((size_t)np * 4));
    // nodes handed from the expand stage to the depth-first refinement (24 B each)
    const unsigned node_cap = (unsigned)std::min<long long>(4ll * np * per_pair, 16ll << 20);
    GLOC_CUDA_TRY(st->nodes.reserve((size_t)node_cap * sizeof(CsmNode)));
    GLOC_CUDA_TRY(st->misc.reserve(64));
    GLOC_CUDA_TRY(cudaMemcpyAsync(st->pairs.p, hp.data() + p0, (size_t)np * sizeof(CsmPairDev),
                                  cudaMemcpyHostToDevice, stream));
    GLOC_CUDA_TRY(cudaMemsetAsync(st->top.p, 0, (size_t)np * 8, stream));
    GLOC_CUDA_TRY(cudaMemsetAsync(st->nsurv.p, 0, (size_t)np * 4, stream));
    GLOC_CUDA_TRY(cudaMemsetAsync(st->misc.p, 0, 64, stream));
    std::vector<unsigned long long> init((size_t)np, init_key);
    GLOC_CUDA_TRY(cudaMemcpyAsync(st->best.p, init.data(), (size_t)np * 8, cudaMemcpyHostToDevice,
                                  stream));
    unsigned* n_surv = (unsigned*)st->nsurv.p;
    unsigned* n_nodes = (unsigned*)st->misc.p;
    unsigned* cursor = n_nodes + 1;
    unsigned long long* counters = (unsigned long long*)((char*)st->misc.p + 16);
    const CsmGridDev* dg = (const CsmGridDev*)st->gridtab.p;
    const CsmPairDev* dp = (const CsmPairDev*)st->pairs.p;
    // tuning aid: GLOC_CSM_TIMING=1 prints the duration of every stage of this sub-batch
    const bool timing = std::getenv("GLOC_CSM_TIMING") != nullptr;
    cudaEvent_t tev[6];
    if (timing) {
      for (auto& e : tev) cudaEventCreate(&e);
      cudaEventRecord(tev[0], stream);
    }
    st->prof.begin(stream);
    cudaError_t ce;
    if (use_bits)
      ce = launch_csm_coarse_bits(dg, dp, np, (const float*)st->pts.p, (const float2*)st->rot.p, prm,
                                  (CsmBounds*)st->bounds.p, (int*)st->coarse.p,
                                  (unsigned long long*)st->top.p, bits_smem, bits_warps, stream);
    else
      ce = launch_csm_coarse(dg, dp, np, (const float*)st->pts.p,
                             (const float2*)st->rot.p, prm, (CsmBounds*)st->bounds.p,
                             (int*)st->coarse.p, (unsigned long long*)st->top.p, stream,
                             use_pm && (size_t)prm.maxc * 4 + 20 * 4096 <= 150 * 1024);
    st->prof.end(stream);
    GLOC_CUDA_TRY(ce);
    if (timing) cudaEventRecord(tev[1], stream);
    GLOC_CUDA_TRY(launch_csm_seed(dg, dp, np, (const float*)st->pts.p, (const float2*)st->rot.p,
                                  prm, (const CsmBounds*)st->bounds.p,
                                  (const unsigned long long*)st->top.p,
                                  (unsigned long long*)st->best.p, stream));
    if (timing) cudaEventRecord(tev[2], stream);
    GLOC_CUDA_TRY(launch_csm_filter(dp, np, prm, (const CsmBounds*)st->bounds.p,
                                    (const int*)st->coarse.p, (const unsigned long long*)st->best.p,
                                    (unsigned*)st->survivors.p, n_surv, stream));
    if (timing) cudaEventRecord(tev[3], stream);
    if (depth >= 2) {
      GLOC_CUDA_TRY(launch_csm_expand(dg, dp, np, (const float*)st->pts.p, (const float2*)st->rot.p,
                                      prm, (const CsmBounds*)st->bounds.p, (const int*)st->coarse.p,
                                      (const unsigned*)st->survivors.p, n_surv,
                                      (unsigned long long*)st->best.p, (CsmNode*)st->nodes.p, n_nodes,
                                      node_cap, counters, 128, exp_bits && use_bits, exp_smem, stream));
      if (timing) cudaEventRecord(tev[5], stream);
      GLOC_CUDA_TRY(launch_csm_refine(dg, dp, (const float*)st->pts.p, (const float2*)st->rot.p, prm,
                                      (const CsmBounds*)st->bounds.p, (const CsmNode*)st->nodes.p,
                                      n_nodes, node_cap, cursor, (unsigned long long*)st->best.p,
                                      counters, n_ctas, stream));
    }
    if (timing) cudaEventRecord(tev[4], stream);
    GLOC_CUDA_TRY(cudaMemcpyAsync(hbest.data() + p0, st->best.p, (size_t)np * 8,
                                  cudaMemcpyDeviceToHost, stream));
    unsigned long long hc = 0;
    unsigned hn = 0;
    GLOC_CUDA_TRY(cudaMemcpyAsync(&hc, counters, 8, cudaMemcpyDeviceToHost, stream));
    GLOC_CUDA_TRY(cudaMemcpyAsync(&hn, n_nodes, 4, cudaMemcpyDeviceToHost, stream));
    GLOC_CUDA_TRY(cudaStreamSynchronize(stream));
    if (hn > node_cap)
      return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: branch-and-bound node list overflowed "
                                  "(more than 16M open nodes in one sub-batch); match fewer pairs per call");
    if (timing) {
      float t[4] = {0, 0, 0, 0};
      std::vector<unsigned> hs((size_t)np);
      cudaMemcpy(hs.data(), n_surv, (size_t)np * 4, cudaMemcpyDeviceToHost);
      unsigned long long ns = 0;
      for (unsigned v : hs) ns += v;
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
      float t_exp = 0.f;
      if (depth >= 2) cudaEventElapsedTime(&t_exp, tev[3], tev[5]);
      fprintf(stderr, "[csm] pairs=%d coarse(%s)=%.3f ms seed=%.3f filter=%.3f expand+refine=%.3f (expand %.3f) | "
                      "survivors=%llu nodes=%u expanded=%llu\n", np, use_bits ? "bits" : "u8", t[0], t[1],
              t[2], t[3], t_exp, ns, hn, hc);
      unsigned mx = 0;
      for (unsigned v : hs) mx = std::max(mx, v);
      fprintf(stderr, "[csm] max survivors in one pair = %u\n", mx);
      for (auto& e : tev) cudaEventDestroy(e);
    }
    st->stats.kernel_launches += depth >= 2 ? 5 : 3;
    st->stats.refined_nodes += hc;
    st->stats.coarse_candidates += (uint64_t)np * (uint64_t)per_pair;  // upper bound (slots)
  }
  st->stats.matches += (uint64_t)n_pairs;

  // decode: Candidate2D (correlative_scan_matcher_2d.h:74-87) + pose (fast_..._2d.cpp:311-318)
  for (int i = 0; i < n_pairs; ++i) {
    gloc_csm_result& r = results[i];
    std::memset(&r, 0, sizeof(r));
    const unsigned long long key = hbest[(size_t)i];
    uint32_t sb = (uint32_t)(key >> 32);
    float score;
    std::memcpy(&score, &sb, 4);
    r.score = min_score;
    if (key != 0 && score > min_score) {
      const uint32_t rank = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
      const int s = (int)(rank / (prm.W * prm.W));
      const int xo = (int)((rank / prm.W) % prm.W) - n_lin;
      const int yo = (int)(rank % prm.W) - n_lin;
      const double resolution = st->grids[grid_ids[i]].resolution;
      const double cx = -yo * resolution;
      const double cy = -xo * resolution;
      const double orientation = (s - n_ang) * ang_step;
      r.found = 1;
      r.score = score;
      r.scan_index = s;
      r.x_offset = xo;
      r.y_offset = yo;
      r.pose_x = init_xyyaw[3 * i] + cx;
      r.pose_y = init_xyyaw[3 * i + 1] + cy;
      r.pose_yaw = init_xyyaw[3 * i + 2] + orientation;
    }
  }
  return GLOC_OK;
}

int gloc_csm_discretize(gloc_csm_store* st, const float* pts, int n_pts, double init_x,
                        double init_y, double init_yaw, int n_ang, double ang_step,
                        double resolution, double max_x, double max_y, int32_t* out_cells) {
  if (!st || !pts || !out_cells || n_pts < 0 || n_ang < 0)
    return fail(GLOC_ERR_INVALID, "gloc_csm_discretize: bad argument");
  if (n_pts == 0) return GLOC_OK;
  DeviceGuard guard(st->device);
  const int S = 2 * n_ang + 1;
  std::vector<float2> rot((size_t)S);
  double delta_theta = -n_ang * ang_step;
  for (int s = 0; s < S; ++s, delta_theta += ang_step) {
    const float ha = 0.5f * (float)delta_theta;
    rot[(size_t)s] = make_float2(std::cos(ha), std::sin(ha));
  }
  const float ha0 = 0.5f * (float)init_yaw;
  GLOC_CUDA_TRY(st->pts.reserve((size_t)n_pts * 3 * sizeof(float)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(st->disc.reserve((size_t)S * n_pts * 2 * sizeof(int)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->pts.p, pts, (size_t)n_pts * 3 * sizeof(float),
                                cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2),
                                cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(launch_csm_discretize((const float*)st->pts.p, n_pts, std::cos(ha0), std::sin(ha0),
                                      (float)init_x, (float)init_y, (const float2*)st->rot.p, S,
                                      resolution, max_x, max_y, (int*)st->disc.p, st->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_cells, st->disc.p, (size_t)S * n_pts * 2 * sizeof(int),
                                cudaMemcpyDeviceToHost, st->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
  st->stats.kernel_launches++;
  return GLOC_OK;
}

// SearchParameters production ctor, correlative_scan_matcher_2d.cpp:27-55 (host-side, O(P))
int gloc_csm_search_params(double linear_window, double angular_window, const float* pts,
                           int n_pts, double resolution, int* n_lin, int* n_ang,
                           double* ang_step) {
  if ((n_pts > 0 && !pts) || !n_lin || !n_ang || !ang_step || !(resolution > 0.))
    return fail(GLOC_ERR_INVALID, "gloc_csm_search_params: bad argument");
  float max_scan_range = 3.f * resolution;
  for (int i = 0; i < n_pts; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1];
    const float range = std::sqrt(x * x + y * y);
    max_scan_range = std::max(range, max_scan_range);
  }
  const double kSafetyMargin = 1. - 1e-3;
  const float r2 = max_scan_range * max_scan_range;
  *ang_step = kSafetyMargin * std::acos(1. - (resolution * resolution) / (2. * r2));
  *n_ang = (int)std::ceil(angular_window / *ang_step);
  *n_lin = (int)std::ceil(linear_window / resolution);
  return GLOC_OK;
}

// GridToVirtualPointCloud, fast_..._2d.cpp:78-95 (host-side, O(cells))
int gloc_csm_grid_to_points(const uint16_t* cells, int nx, int ny, double resolution, double ox,
                            double oy, float* pts, int capacity, int* n_out) {
  if (!cells || nx < 1 || ny < 1 || !n_out)
    return fail(GLOC_ERR_INVALID, "gloc_csm_grid_to_points: bad argument");
  int n = 0;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const float cost = value_to_cost(cells[(size_t)nx * j + i]);
      if (cost < 0.11) {
        if (pts && n < capacity) {
          pts[3 * n] = (float)(ox + i * resolution);
          pts[3 * n + 1] = (float)(oy + j * resolution);
          pts[3 * n + 2] = 0.f;
        }
        ++n;
      }
    }
  *n_out = n;
  return GLOC_OK;
}

int gloc_csm_set_profiling(gloc_csm_store* st, int enabled) {
  if (!st) return fail(GLOC_ERR_INVALID, "gloc_csm_set_profiling: null store");
  st->prof.enabled = enabled != 0;
  return GLOC_OK;
}

int gloc_csm_get_profile(gloc_csm_store* st, gloc_profile* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_profile: null argument");
  DeviceGuard g(st->device);
  st->prof.collect(&out->dominant_ms, &out->dominant_launches);
  return GLOC_OK;
}

int gloc_csm_get_stats(const gloc_csm_store* st, gloc_csm_stats* stats) {
  if (!st || !stats) return fail(GLOC_ERR_INVALID, "gloc_csm_get_stats: null argument");
  *stats = st->stats;
  return GLOC_OK;
}

}  // extern "C"
