// vlad.cu -- the NetVLAD_fc pooling head (the last step of descriptor
// extraction): encoder feature maps [B][C][S] on the device -> 512-d place descriptors on the
// device, batched, ready for gloc_knn_query_device without a host round trip.
//
// Replaces NetVLAD.forward of /root/reference/model/netvlad_fc.py:73-109 (vladv2 = False, no
// gating) as traced into the TorchScript module that RpyPCLoopDetector::get_place_feature runs
// one frame at a time (loop_detector.cpp:137-172):
//   x^ = x / max(||x[:, s]||, 1e-12);  a = softmax_k(W x^ (+ b));  V[k] = sum_s a[k,s] (x^[:,s] - c_k)
//   V[k] /= max(||V[k]||, 1e-12);  v = vec(V) / max(||vec(V)||, 1e-12);  out = v^T H
// FP32 throughout (SIMT; the head is 0.3 GFLOP per frame against the encoder's 360), sums in a
// fixed order (no atomics): results do not depend on the batch a frame travels in.  The
// encoder (VGG16 convolutions) is not part of this file.
//
// STATUS: written without a GPU at hand (round 1 ran out of GPU minutes): compiles for
// sm_100a, checked by reading only.  tests/test_vlad_gpu.py is the parity test against the
// oracle (which is pinned to the reference's own module); it is opt-in until it has run once.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/gloc3d.h"
#include "common.cuh"

namespace gloc {

namespace {

// [kernels-begin] (tests/cpp/vlad_emu_test.cpp compiles the text up to [kernels-end] for the host)
constexpr int kVladMaxK = 64;        // clusters (accumulators per thread in the assignment kernel)
constexpr int kAssignThreads = 128;  // locations per CTA
constexpr int kFcRows = 256;         // rows of the hidden matrix per CTA
constexpr int kFcCols = 128;         // output columns per CTA (one per thread)
constexpr int kFcBatch = 8;          // frames per pass over the hidden matrix
constexpr float kNormEps = 1e-12f;   // F.normalize's eps

// K1: per location, L2 norm over the channels, the 1x1 convolution onto the K clusters and the
// softmax over them.  thread = location (coalesced reads of x along s), W staged in shared
// memory and read as a broadcast.  a: [B][K][S], inv: [B][S].
__global__ void __launch_bounds__(kAssignThreads)
vlad_assign_kernel(const float* __restrict__ x, const float* __restrict__ conv_w,
                   const float* __restrict__ conv_b, int C, int S, int K, float* __restrict__ a,
                   float* __restrict__ inv) {
  extern __shared__ float w_s[];   // [K][C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < K * C; i += kAssignThreads) w_s[i] = conv_w[i];
  __syncthreads();
  const int s = blockIdx.x * kAssignThreads + threadIdx.x;
  if (s >= S) return;
  const float* xb = x + (size_t)b * C * S + s;
  float acc[kVladMaxK];
#pragma unroll
  for (int k = 0; k < kVladMaxK; ++k) acc[k] = 0.f;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) {
    const float xv = __ldg(xb + (size_t)c * S);
    ss = fmaf(xv, xv, ss);
#pragma unroll
    for (int k = 0; k < kVladMaxK; ++k)
      if (k < K) acc[k] = fmaf(w_s[k * C + c], xv, acc[k]);
  }
  const float r = 1.f / fmaxf(sqrtf(ss), kNormEps);
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kVladMaxK; ++k)
    if (k < K) {
      acc[k] = acc[k] * r + (conv_b ? conv_b[k] : 0.f);
      mx = fmaxf(mx, acc[k]);
    }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kVladMaxK; ++k)
    if (k < K) {
      acc[k] = expf(acc[k] - mx);
      sum += acc[k];
    }
  const float rs = 1.f / sum;
  float* ab = a + (size_t)b * K * S + s;
#pragma unroll
  for (int k = 0; k < kVladMaxK; ++k)
    if (k < K) ab[(size_t)k * S] = acc[k] * rs;
  inv[(size_t)b * S + s] = r;
}

// K2: V[b][k][c] = sum_s a[k][s] x^[c][s] - (sum_s a[k][s]) cent[k][c] for a tile of 32 channels
// and all clusters.  256 threads: c = t & 31, eight clusters (t >> 5) * 8 .. + 7 each; tiles of
// 32 locations go through shared memory (x^ tile padded to 33, a tile read as a broadcast).
__global__ void __launch_bounds__(256)
vlad_aggregate_kernel(const float* __restrict__ x, const float* __restrict__ a,
                      const float* __restrict__ inv, const float* __restrict__ cent, int C, int S,
                      int K, float* __restrict__ V) {
  __shared__ float xs[32][33];
  __shared__ float as[kVladMaxK][32];
  const int b = blockIdx.y, c0 = blockIdx.x * 32;
  const int t = threadIdx.x, ci = t & 31, kg = t >> 5;
  const float* xb = x + (size_t)b * C * S;
  const float* ab = a + (size_t)b * K * S;
  const float* ib = inv + (size_t)b * S;
  float acc[8], asum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = asum[j] = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int si = t & 31, s = s0 + si;
    const float r = s < S ? ib[s] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {          // x^ tile: rows (t >> 5) + 8 j of the 32 channels
      const int cc = (t >> 5) + 8 * j;
      xs[cc][si] = (s < S && c0 + cc < C) ? __ldg(xb + (size_t)(c0 + cc) * S + s) * r : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {          // a tile: rows (t >> 5) + 8 j of the clusters
      const int k = (t >> 5) + 8 * j;
      as[k][si] = (s < S && k < K) ? __ldg(ab + (size_t)k * S + s) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const float xv = xs[ci][i];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float av = as[kg * 8 + j][i];
        acc[j] = fmaf(av, xv, acc[j]);
        asum[j] += av;
      }
    }
    __syncthreads();
  }
  const int c = c0 + ci;
  if (c < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      if (k < K) V[((size_t)b * K + k) * C + c] = acc[j] - asum[j] * __ldg(cent + (size_t)k * C + c);
    }
  }
}

// K3: intra-normalisation of every cluster row, then L2 normalisation of the whole K*C vector,
// in place.  One CTA of 256 threads per frame; warp w owns clusters w, w + 8, ...
__global__ void __launch_bounds__(256)
vlad_normalize_kernel(float* __restrict__ V, int C, int K) {
  __shared__ float f_s[kVladMaxK], r_s[kVladMaxK];
  __shared__ float g_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Vb = V + (size_t)b * K * C;
  for (int k = warp; k < K; k += 8) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = Vb[(size_t)k * C + c];
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
      const float n = sqrtf(ss);
      const float f = 1.f / fmaxf(n, kNormEps);
      f_s[k] = f;
      r_s[k] = (n * f) * (n * f);   // squared norm of the row after its own normalisation
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < K; ++k) tot += r_s[k];
    g_s = 1.f / fmaxf(sqrtf(tot), kNormEps);
  }
  __syncthreads();
  const float g = g_s;
  for (int i = threadIdx.x; i < K * C; i += 256) Vb[i] *= f_s[i / C] * g;
}

// K4: partial[chunk][b][j] = sum over the chunk's kFcRows rows i of v[b][i] H[i][j].  The hidden
// matrix (K*C x D, 64 MB at the reference's sizes) is read once per pass of up to kFcBatch
// frames, coalesced along j; the frames' slices of v sit in shared memory.
__global__ void __launch_bounds__(kFcCols)
vlad_fc_kernel(const float* __restrict__ v, const float* __restrict__ H, int I, int D, int b0,
               int nb, float* __restrict__ partial, int B) {
  __shared__ float v_s[kFcBatch][kFcRows];
  const int j = blockIdx.x * kFcCols + threadIdx.x;
  const int chunk = blockIdx.y, i0 = chunk * kFcRows;
  for (int e = threadIdx.x; e < kFcBatch * kFcRows; e += kFcCols) {
    const int bb = e / kFcRows, ii = e % kFcRows;
    v_s[bb][ii] = (bb < nb && i0 + ii < I) ? v[(size_t)(b0 + bb) * I + i0 + ii] : 0.f;
  }
  __syncthreads();
  if (j >= D) return;
  float acc[kFcBatch];
#pragma unroll
  for (int bb = 0; bb < kFcBatch; ++bb) acc[bb] = 0.f;
  const int rows = min(kFcRows, I - i0);
  for (int ii = 0; ii < rows; ++ii) {
    const float w = __ldg(H + (size_t)(i0 + ii) * D + j);
#pragma unroll
    for (int bb = 0; bb < kFcBatch; ++bb) acc[bb] = fmaf(v_s[bb][ii], w, acc[bb]);
  }
  for (int bb = 0; bb < nb; ++bb) partial[((size_t)chunk * B + b0 + bb) * D + j] = acc[bb];
}

// K5: out[b][j] = sum over the chunks, in chunk order.
__global__ void vlad_fc_reduce_kernel(const float* __restrict__ partial, int n_chunks, int B, int D,
                                      float* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)B * D) return;
  float s = 0.f;
  for (int ch = 0; ch < n_chunks; ++ch) s += partial[(size_t)ch * B * D + e];
  out[e] = s;
}

// [kernels-end]

struct DevBuf {
  float* p = nullptr;
  size_t n = 0;
  cudaError_t reserve(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(float));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

}  // namespace gloc

struct gloc_vlad_head {
  int device = 0, C = 0, K = 0, D = 0;
  bool has_bias = false;
  float *d_w = nullptr, *d_b = nullptr, *d_cent = nullptr, *d_hidden = nullptr;
  gloc::DevBuf a, inv, V, partial, feat, out;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
};

using gloc::fail;

namespace {

int forward_device(gloc_vlad_head* h, const float* d_feat, int B, int S, float* d_out) {
  using namespace gloc;
  const int C = h->C, K = h->K, D = h->D, I = K * C;
  const int n_chunks = (I + kFcRows - 1) / kFcRows;
  GLOC_CUDA_TRY(h->a.reserve((size_t)B * K * S));
  GLOC_CUDA_TRY(h->inv.reserve((size_t)B * S));
  GLOC_CUDA_TRY(h->V.reserve((size_t)B * I));
  GLOC_CUDA_TRY(h->partial.reserve((size_t)n_chunks * B * D));
  cudaStream_t st = h->stream;
  const size_t smem = (size_t)K * C * sizeof(float);
  static unsigned long long attr_mask = 0;
  if (first_use_on_current_device(attr_mask))
    GLOC_CUDA_TRY(cudaFuncSetAttribute(vlad_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  vlad_assign_kernel<<<dim3((S + kAssignThreads - 1) / kAssignThreads, B), kAssignThreads, smem, st>>>(
      d_feat, h->d_w, h->has_bias ? h->d_b : nullptr, C, S, K, h->a.p, h->inv.p);
  GLOC_CUDA_TRY(cudaGetLastError());
  vlad_aggregate_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(d_feat, h->a.p, h->inv.p, h->d_cent, C, S, K,
                                                                h->V.p);
  GLOC_CUDA_TRY(cudaGetLastError());
  vlad_normalize_kernel<<<B, 256, 0, st>>>(h->V.p, C, K);
  GLOC_CUDA_TRY(cudaGetLastError());
  h->launches += 3;
  for (int b0 = 0; b0 < B; b0 += kFcBatch) {
    const int nb = std::min(kFcBatch, B - b0);
    vlad_fc_kernel<<<dim3((D + kFcCols - 1) / kFcCols, n_chunks), kFcCols, 0, st>>>(h->V.p, h->d_hidden, I, D, b0,
                                                                                   nb, h->partial.p, B);
    GLOC_CUDA_TRY(cudaGetLastError());
    ++h->launches;
  }
  const size_t n_out = (size_t)B * D;
  vlad_fc_reduce_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(h->partial.p, n_chunks, B, D, d_out);
  GLOC_CUDA_TRY(cudaGetLastError());
  ++h->launches;
  return GLOC_OK;
}

}  // namespace

extern "C" {

int gloc_vlad_create(gloc_vlad_head** out, int device, int dim, int clusters, int out_dim,
                     const float* conv_w, const float* conv_b, const float* centroids,
                     const float* hidden_w) {
  if (!out || !conv_w || !centroids || !hidden_w)
    return fail(GLOC_ERR_INVALID, "gloc_vlad_create: null argument");
  *out = nullptr;
  if (dim < 32 || dim % 32 != 0 || clusters < 1 || clusters > gloc::kVladMaxK || out_dim < 1 ||
      (size_t)clusters * dim * sizeof(float) > (size_t)200 * 1024)
    return fail(GLOC_ERR_RANGE, "gloc_vlad_create: needs dim % 32 == 0, 1 <= clusters <= 64, clusters * dim <= 51200");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_vlad_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= n_dev) return fail(GLOC_ERR_INVALID, "gloc_vlad_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(GLOC_ERR_CUDA, "gloc_vlad_create: device is not sm_100 (kernels are sm_100a only)");
  gloc::DeviceGuard scope(device);
  gloc_vlad_head* h = new gloc_vlad_head;
  h->device = device;
  h->C = dim;
  h->K = clusters;
  h->D = out_dim;
  h->has_bias = conv_b != nullptr;
  const size_t kc = (size_t)clusters * dim;
  cudaError_t e = cudaStreamCreate(&h->stream);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_w, kc * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_cent, kc * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_hidden, kc * (size_t)out_dim * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_b, (size_t)clusters * 4);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_w, conv_w, kc * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_cent, centroids, kc * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_hidden, hidden_w, kc * (size_t)out_dim * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && conv_b) e = cudaMemcpy(h->d_b, conv_b, (size_t)clusters * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    const std::string msg = std::string("gloc_vlad_create: ") + cudaGetErrorString(e);
    gloc_vlad_destroy(h);
    return fail(GLOC_ERR_CUDA, msg);
  }
  *out = h;
  return GLOC_OK;
}

void gloc_vlad_destroy(gloc_vlad_head* h) {
  if (!h) return;
  gloc::DeviceGuard scope(h->device);
  for (float* p : {h->d_w, h->d_b, h->d_cent, h->d_hidden})
    if (p) cudaFree(p);
  for (gloc::DevBuf* b : {&h->a, &h->inv, &h->V, &h->partial, &h->feat, &h->out}) b->release();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int gloc_vlad_forward_device(gloc_vlad_head* h, const float* d_feat, int batch, int n_loc, float* d_out) {
  if (!h || !d_feat || !d_out) return fail(GLOC_ERR_INVALID, "gloc_vlad_forward_device: null argument");
  if (batch < 0 || n_loc < 1) return fail(GLOC_ERR_INVALID, "gloc_vlad_forward_device: batch >= 0 and n_loc >= 1 required");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(h->device);
  const int rc = forward_device(h, d_feat, batch, n_loc, d_out);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GLOC_OK;
}

int gloc_vlad_forward(gloc_vlad_head* h, const float* feat, int batch, int n_loc, float* out) {
  if (!h || !feat || !out) return fail(GLOC_ERR_INVALID, "gloc_vlad_forward: null argument");
  if (batch < 0 || n_loc < 1) return fail(GLOC_ERR_INVALID, "gloc_vlad_forward: batch >= 0 and n_loc >= 1 required");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(h->device);
  const size_t n_in = (size_t)batch * h->C * n_loc, n_out = (size_t)batch * h->D;
  GLOC_CUDA_TRY(h->feat.reserve(n_in));
  GLOC_CUDA_TRY(h->out.reserve(n_out));
  GLOC_CUDA_TRY(cudaMemcpyAsync(h->feat.p, feat, n_in * 4, cudaMemcpyHostToDevice, h->stream));
  const int rc = forward_device(h, h->feat.p, batch, n_loc, h->out.p);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpyAsync(out, h->out.p, n_out * 4, cudaMemcpyDeviceToHost, h->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GLOC_OK;
}

uint64_t gloc_vlad_kernel_launches(const gloc_vlad_head* h) { return h ? h->launches : 0; }

}  // extern "C"
