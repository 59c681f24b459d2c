// bev.cu -- BEV projection of one LiDAR scan on the GPU (the producer of
// both stages' inputs) + its C ABI.
//
// Replaces RpyPCLoopDetector::get_projected_grid (/root/reference/registration/
// loop_detector.cpp:122-135), i.e. for a fresh Submap3D with the identity pose:
//   point_cloud_to_range_data            loop_detector.cpp:108-120   (range <= 100 m -> return)
//   Submap3D::InsertRangeData            3d/submap_3d.cpp:162-177, FilterRangeDataByMaxRange :43-52
//   RangeDataInserter3D::Insert          3d/range_data_inserter_3d.cpp:63-77 (hit voxel -> p = 0.55)
//   HybridGrid::GetCellIndex             3d/hybrid_grid.h:429-434   (lround(p / resolution), float)
//   ProjectToCvMat / ProjectToGrid       3d/submap_3d.cpp:238-326, :328-429
//   crop_pad_occupancy                   loop_detector.cpp:83-106
// What survives of that pipeline for ONE scan (derivation in oracle/bev_oracle.c): the set of
// hit voxels; a pixel (ix, iy) is occupied iff its column holds >= 2 distinct hit voxels
// (2 x 0.55 > kMaxProbability >= 0.55); the image spans the bounding box of all hit voxels.
// "At least two distinct z" == (max z > min z), so the HybridGrid is replaced by two dense
// per-column arrays updated with atomicMin / atomicMax -- no hashing, no sorting, exact.
#include <algorithm>
#include <climits>
#include <cmath>
#include <new>

#include "common.cuh"

using namespace gloc;

struct gloc_bev_projector {
  int device = 0;
  cudaStream_t stream = nullptr;
  float resolution = 0.2f, max_range = 100.f;
  int R = 0, side = 0;            // dense column array covers voxel indices [-R, R]^2
  int* d_zmin = nullptr;          // [side*side]
  int* d_zmax = nullptr;
  int* d_box = nullptr;           // min_ix, max_ix, min_iy, max_iy, n_occupied, n_in_range
  float* d_pts = nullptr;
  size_t pts_cap = 0;
  uint8_t* d_img = nullptr;       // [h][w], 0 = occupied, 255 = free (the reference's cv::Mat)
  size_t img_cap = 0;
  gloc_bev_info info{};
  bool valid = false;
  uint64_t launches = 0;
};

namespace {

__global__ void bev_reset_kernel(int* __restrict__ zmin, int* __restrict__ zmax, int n,
                                 int* __restrict__ box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    zmin[i] = INT_MAX;
    zmax[i] = INT_MIN;
  }
  if (i == 0) {
    box[0] = INT_MAX; box[1] = INT_MIN; box[2] = INT_MAX; box[3] = INT_MIN; box[4] = 0; box[5] = 0;
  }
}

// loop_detector.cpp:112 + submap_3d.cpp:47 (range test), hybrid_grid.h:429-434 (voxel index)
__global__ void bev_voxelize_kernel(const float* __restrict__ pts, size_t n, int stride,
                                    float resolution, float max_range, int R, int side,
                                    int* __restrict__ zmin, int* __restrict__ zmax,
                                    int* __restrict__ box) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ix = 0, iy = 0;
  bool hit = false;
  if (i < n) {
    const float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    if (__fsqrt_rn(r2) <= max_range) {
      ix = (int)lroundf(__fdiv_rn(x, resolution));
      iy = (int)lroundf(__fdiv_rn(y, resolution));
      const int iz = (int)lroundf(__fdiv_rn(z, resolution));
      if (abs(ix) <= R && abs(iy) <= R) {   // always true: |x|, |y| <= range <= max_range
        hit = true;
        const int c = (iy + R) * side + (ix + R);
        atomicMin(zmin + c, iz);
        atomicMax(zmax + c, iz);
      }
    }
  }
  // bounding box of the hit voxels, one atomic per warp and bound
  int mnx = hit ? ix : INT_MAX, mxx = hit ? ix : INT_MIN, mny = hit ? iy : INT_MAX, mxy = hit ? iy : INT_MIN;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  const unsigned any = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && any) {
    atomicMin(box + 0, mnx);
    atomicMax(box + 1, mxx);
    atomicMin(box + 2, mny);
    atomicMax(box + 3, mxy);
    atomicAdd(box + 5, __popc(any));
  }
}

// submap_3d.cpp:294-325: pixel = 0 (occupied) iff the column's probability sum > 0.9
__global__ void bev_image_kernel(const int* __restrict__ zmin, const int* __restrict__ zmax,
                                 int R, int side, int* __restrict__ box, uint8_t* __restrict__ img,
                                 int w, int h, int min_ix, int min_iy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  bool occ = false;
  if (x < w && y < h) {
    const int c = (y + min_iy + R) * side + (x + min_ix + R);
    occ = zmax[c] > zmin[c];
    img[(size_t)y * w + x] = occ ? 0 : 255;
  }
  const unsigned m = __ballot_sync(0xffffffffu, occ);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(box + 4, __popc(m));
}

// crop_pad_occupancy (loop_detector.cpp:83-106), one channel
__global__ void bev_crop_pad_kernel(const uint8_t* __restrict__ src, int sw, int sh, int width,
                                    int height, int cw, int ch, int sx, int sy, int dx, int dy,
                                    uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width || y >= height) return;
  uint8_t v = 255;
  const int rx = x - dx, ry = y - dy;
  if (rx >= 0 && rx < cw && ry >= 0 && ry < ch) v = src[(size_t)(sy + ry) * sw + (sx + rx)];
  dst[(size_t)y * width + x] = v;
}

// ProjectToGrid (submap_3d.cpp:328-429) -> the matcher's width-1 precomputation grid:
// occupied pixel -> probability 0.9 -> cost 0.1 -> 255; free -> probability 0.1 -> 0.
__global__ void bev_level1_kernel(const uint8_t* __restrict__ img, size_t n, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = img[i] == 0 ? 255 : 0;
}

// MapLimits-consistent layout (see gloc_csm_add_grid_from_bev_aligned): cell (cx, cy) holds
// pixel (ix, iy) = (W - 1 - cy, H - 1 - cx); flat index H * cy + cx (num_x_cells = H).
__global__ void bev_level1_aligned_kernel(const uint8_t* __restrict__ img, int W, int H,
                                          uint8_t* __restrict__ out) {
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
  if (cx >= H || cy >= W) return;
  out[(size_t)H * cy + cx] = img[(size_t)(H - 1 - cx) * W + (W - 1 - cy)] == 0 ? 255 : 0;
}

// GridToVirtualPointCloud (2d/fast_correlative_scan_matcher_2d.cpp:78-95): i outer, j inner
__global__ void bev_points_kernel(const uint8_t* __restrict__ img, int w, int h, double ox, double oy,
                                  double res, const int* __restrict__ col_prefix,
                                  float* __restrict__ pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // column (x cell)
  if (i >= w) return;
  int at = col_prefix[i];
  for (int j = 0; j < h; ++j) {
    if (img[(size_t)j * w + i] == 0) {
      pts[3 * (size_t)at] = (float)(ox + i * res);
      pts[3 * (size_t)at + 1] = (float)(oy + j * res);
      pts[3 * (size_t)at + 2] = 0.f;
      ++at;
    }
  }
}

__global__ void bev_col_count_kernel(const uint8_t* __restrict__ img, int w, int h, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w) return;
  int c = 0;
  for (int j = 0; j < h; ++j) c += img[(size_t)j * w + i] == 0;
  cnt[i] = c;
}

}  // namespace

// used by csm_api.cu (gloc_csm_add_grid_from_bev)
const uint8_t* gloc_bev_device_image(const gloc_bev_projector* b, gloc_bev_info* info) {
  if (!b || !b->valid) return nullptr;
  *info = b->info;
  return b->d_img;
}
int gloc_bev_device_of(const gloc_bev_projector* b) { return b ? b->device : -1; }
cudaError_t gloc_bev_launch_level1(const uint8_t* img, size_t n, uint8_t* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  bev_level1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, n, out);
  return cudaGetLastError();
}

cudaError_t gloc_bev_launch_level1_aligned(const uint8_t* img, int W, int H, uint8_t* out,
                                           cudaStream_t s) {
  dim3 grd((H + 127) / 128, W);
  bev_level1_aligned_kernel<<<grd, 128, 0, s>>>(img, W, H, out);
  return cudaGetLastError();
}

extern "C" {

int gloc_bev_create(gloc_bev_projector** out, int device, float resolution, float max_range) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_bev_create: out is null");
  *out = nullptr;
  if (!(resolution > 0.f) || !(max_range > 0.f) || !(max_range / resolution < 8000.f))
    return fail(GLOC_ERR_RANGE, "gloc_bev_create: need resolution > 0 and max_range / resolution < 8000");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_bev_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(GLOC_ERR_INVALID, "gloc_bev_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(GLOC_ERR_CUDA, "gloc_bev_create: device is not sm_100 (kernels are sm_100a only)");
  DeviceGuard g(device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_bev_create: cudaSetDevice failed");
  gloc_bev_projector* b = new (std::nothrow) gloc_bev_projector;
  if (!b) return fail(GLOC_ERR_NOMEM, "gloc_bev_create: out of host memory");
  b->device = device;
  b->resolution = resolution;
  b->max_range = max_range;
  b->R = (int)std::lround(max_range / resolution) + 2;
  b->side = 2 * b->R + 1;
  const size_t n = (size_t)b->side * b->side;
  cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_zmin, n * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_zmax, n * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_box, 8 * sizeof(int));
  if (e != cudaSuccess) {
    gloc_bev_destroy(b);
    return fail(GLOC_ERR_CUDA, std::string("gloc_bev_create: ") + cudaGetErrorString(e));
  }
  *out = b;
  return GLOC_OK;
}

void gloc_bev_destroy(gloc_bev_projector* b) {
  if (!b) return;
  DeviceGuard g(b->device);
  if (b->stream) {
    cudaStreamSynchronize(b->stream);
    cudaStreamDestroy(b->stream);
  }
  for (void* p : {(void*)b->d_zmin, (void*)b->d_zmax, (void*)b->d_box, (void*)b->d_pts, (void*)b->d_img})
    if (p) cudaFree(p);
  delete b;
}

int gloc_bev_project(gloc_bev_projector* b, const float* pts, size_t n_pts, int stride,
                     gloc_bev_info* info) {
  if (!b || (n_pts > 0 && !pts) || stride < 3)
    return fail(GLOC_ERR_INVALID, "gloc_bev_project: bad argument");
  DeviceGuard g(b->device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_bev_project: cudaSetDevice failed");
  b->valid = false;
  cudaStream_t s = b->stream;
  const size_t need = n_pts * (size_t)stride;
  if (need > b->pts_cap) {
    if (b->d_pts) cudaFree(b->d_pts);
    b->d_pts = nullptr;
    b->pts_cap = 0;
    GLOC_CUDA_TRY(cudaMalloc((void**)&b->d_pts, (need + need / 4 + 64) * sizeof(float)));
    b->pts_cap = need + need / 4 + 64;
  }
  if (need) GLOC_CUDA_TRY(cudaMemcpyAsync(b->d_pts, pts, need * sizeof(float), cudaMemcpyHostToDevice, s));
  const int ncol = b->side * b->side;
  bev_reset_kernel<<<(ncol + 255) / 256, 256, 0, s>>>(b->d_zmin, b->d_zmax, ncol, b->d_box);
  GLOC_CUDA_TRY(cudaGetLastError());
  if (n_pts) {
    bev_voxelize_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, s>>>(
        b->d_pts, n_pts, stride, b->resolution, b->max_range, b->R, b->side, b->d_zmin, b->d_zmax, b->d_box);
    GLOC_CUDA_TRY(cudaGetLastError());
  }
  int box[6];
  GLOC_CUDA_TRY(cudaMemcpyAsync(box, b->d_box, sizeof(box), cudaMemcpyDeviceToHost, s));
  GLOC_CUDA_TRY(cudaStreamSynchronize(s));
  b->launches += 2;
  gloc_bev_info& I = b->info;
  I = gloc_bev_info{};
  I.resolution = (double)b->resolution;
  I.n_points_in_range = (uint64_t)box[5];
  if (box[5] > 0) {
    I.width = box[1] - box[0] + 1;
    I.height = box[3] - box[2] + 1;
    I.min_ix = box[0];
    I.min_iy = box[2];
    I.ox = box[0] * (double)b->resolution;   // submap_3d.cpp:271-272
    I.oy = box[2] * (double)b->resolution;
    const size_t cells = (size_t)I.width * I.height;
    if (cells > b->img_cap) {
      if (b->d_img) cudaFree(b->d_img);
      b->d_img = nullptr;
      b->img_cap = 0;
      GLOC_CUDA_TRY(cudaMalloc((void**)&b->d_img, cells + cells / 4 + 64));
      b->img_cap = cells + cells / 4 + 64;
    }
    dim3 grd((I.width + 127) / 128, I.height);
    bev_image_kernel<<<grd, 128, 0, s>>>(b->d_zmin, b->d_zmax, b->R, b->side, b->d_box, b->d_img,
                                         I.width, I.height, I.min_ix, I.min_iy);
    GLOC_CUDA_TRY(cudaGetLastError());
    GLOC_CUDA_TRY(cudaMemcpyAsync(box, b->d_box, sizeof(box), cudaMemcpyDeviceToHost, s));
    GLOC_CUDA_TRY(cudaStreamSynchronize(s));
    I.n_occupied = (uint64_t)box[4];
    b->launches += 1;
  }
  b->valid = true;
  if (info) *info = I;
  return GLOC_OK;
}

int gloc_bev_get_image(gloc_bev_projector* b, uint8_t* img, size_t capacity) {
  if (!b || !b->valid) return fail(GLOC_ERR_NOT_BUILT, "gloc_bev_get_image: no projection yet");
  const size_t cells = (size_t)b->info.width * b->info.height;
  if (cells == 0) return GLOC_OK;
  if (!img || capacity < cells) return fail(GLOC_ERR_INVALID, "gloc_bev_get_image: buffer too small");
  DeviceGuard g(b->device);
  GLOC_CUDA_TRY(cudaMemcpy(img, b->d_img, cells, cudaMemcpyDeviceToHost));
  return GLOC_OK;
}

int gloc_bev_get_cnn_input_roi(gloc_bev_projector* b, int width, int height, uint8_t* out, int32_t roi[4]) {
  if (!roi) return fail(GLOC_ERR_INVALID, "gloc_bev_get_cnn_input_roi: roi is null");
  const int rc = gloc_bev_get_cnn_input(b, width, height, out);
  if (rc != GLOC_OK) return rc;
  // roi_dst of crop_pad_occupancy, loop_detector.cpp:99-102
  const int sw = b->info.width, sh = b->info.height;
  const int cw = sw >= width ? width : sw, ch = sh >= height ? height : sh;
  roi[0] = (int)std::floor((width - cw) / 2.);
  roi[1] = (int)std::floor((height - ch) / 2.);
  roi[2] = cw;
  roi[3] = ch;
  return GLOC_OK;
}

int gloc_bev_get_cnn_input(gloc_bev_projector* b, int width, int height, uint8_t* out) {
  if (!b || !b->valid) return fail(GLOC_ERR_NOT_BUILT, "gloc_bev_get_cnn_input: no projection yet");
  if (!out || width < 1 || height < 1) return fail(GLOC_ERR_INVALID, "gloc_bev_get_cnn_input: bad argument");
  DeviceGuard g(b->device);
  const int sw = b->info.width, sh = b->info.height;
  const int cw = sw >= width ? width : sw, ch = sh >= height ? height : sh;
  const int sx = (int)std::floor((sw - cw) / 2.), sy = (int)std::floor((sh - ch) / 2.);
  const int dx = (int)std::floor((width - cw) / 2.), dy = (int)std::floor((height - ch) / 2.);
  uint8_t* d_out = nullptr;
  GLOC_CUDA_TRY(cudaMalloc((void**)&d_out, (size_t)width * height));
  dim3 grd((width + 127) / 128, height);
  bev_crop_pad_kernel<<<grd, 128, 0, b->stream>>>(b->d_img, sw, sh, width, height, cw, ch, sx, sy, dx,
                                                  dy, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, d_out, (size_t)width * height, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(d_out);
  b->launches += 1;
  if (e != cudaSuccess) return fail(GLOC_ERR_CUDA, std::string("gloc_bev_get_cnn_input: ") + cudaGetErrorString(e));
  return GLOC_OK;
}

int gloc_bev_get_occupied_points(gloc_bev_projector* b, float* pts, size_t capacity, size_t* n_out) {
  if (!b || !b->valid) return fail(GLOC_ERR_NOT_BUILT, "gloc_bev_get_occupied_points: no projection yet");
  if (!n_out) return fail(GLOC_ERR_INVALID, "gloc_bev_get_occupied_points: n_out is null");
  *n_out = (size_t)b->info.n_occupied;
  if (!pts || b->info.n_occupied == 0) return GLOC_OK;
  if (capacity < b->info.n_occupied) return fail(GLOC_ERR_INVALID, "gloc_bev_get_occupied_points: buffer too small");
  DeviceGuard g(b->device);
  const int w = b->info.width, h = b->info.height;
  int* d_cnt = nullptr;
  float* d_pts = nullptr;
  GLOC_CUDA_TRY(cudaMalloc((void**)&d_cnt, (size_t)w * sizeof(int)));
  cudaError_t e = cudaMalloc((void**)&d_pts, (size_t)b->info.n_occupied * 3 * sizeof(float));
  std::vector<int> cnt((size_t)w);
  if (e == cudaSuccess) {
    bev_col_count_kernel<<<(w + 127) / 128, 128, 0, b->stream>>>(b->d_img, w, h, d_cnt);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)w * sizeof(int), cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (e == cudaSuccess) {   // exclusive prefix over at most ~1000 columns: host
    int run = 0;
    for (int i = 0; i < w; ++i) { const int c = cnt[(size_t)i]; cnt[(size_t)i] = run; run += c; }
    e = cudaMemcpyAsync(d_cnt, cnt.data(), (size_t)w * sizeof(int), cudaMemcpyHostToDevice, b->stream);
  }
  if (e == cudaSuccess) {
    bev_points_kernel<<<(w + 127) / 128, 128, 0, b->stream>>>(b->d_img, w, h, b->info.ox, b->info.oy,
                                                              b->info.resolution, d_cnt, d_pts);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(pts, d_pts, (size_t)b->info.n_occupied * 3 * sizeof(float), cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(d_cnt);
  if (d_pts) cudaFree(d_pts);
  b->launches += 2;
  if (e != cudaSuccess) return fail(GLOC_ERR_CUDA, std::string("gloc_bev_get_occupied_points: ") + cudaGetErrorString(e));
  return GLOC_OK;
}

uint64_t gloc_bev_kernel_launches(const gloc_bev_projector* b) { return b ? b->launches : 0; }

}  // extern "C"
