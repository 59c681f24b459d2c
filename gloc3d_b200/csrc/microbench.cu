// microbench.cu -- measured ceilings for kernels whose binding unit is neither HBM nor the tensor
// pipe.  SURVEY 8d asks for the stage-2 scorer's gather rate "against the real ceiling, measured
// with a gather micro-benchmark": the coarse scorer (csm_coarse_bits_kernel) lives on random
// 8-byte shared-memory loads, so the ceiling is the chip-wide rate of exactly those -- random
// LDS.64 with the bank conflicts random addresses bring.
#include <algorithm>
#include <string>

#include "../../include/gloc3d.h"
#include "common.cuh"

namespace gloc {
namespace {

constexpr int kGatherWords = 16384;     // 128 KB table of 8-byte words per CTA (the scorer's bit planes: ~170 KB)
constexpr int kGatherThreads = 1024;

__global__ void __launch_bounds__(kGatherThreads)
smem_gather_kernel(unsigned long long* __restrict__ sink, int iters) {
  extern __shared__ unsigned long long gather_tab[];
  for (int i = threadIdx.x; i < kGatherWords; i += kGatherThreads)
    gather_tab[i] = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; it += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x = x * 1664525u + 1013904223u;                 // independent of the loaded data: loads overlap
      acc ^= gather_tab[(x >> 9) & (kGatherWords - 1)];
    }
  }
  if (acc == 0x1234567ull) sink[0] = acc;             // keeps the loads alive
}

}  // namespace
}  // namespace gloc

using gloc::fail;

extern "C" int gloc_bench_smem_gather(int device, double* loads_per_s) {
  using namespace gloc;
  if (!loads_per_s) return fail(GLOC_ERR_INVALID, "gloc_bench_smem_gather: null argument");
  *loads_per_s = 0.0;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_bench_smem_gather: no CUDA device");
  }
  if (device < 0 || device >= n_dev) return fail(GLOC_ERR_INVALID, "gloc_bench_smem_gather: bad device");
  DeviceGuard scope(device);
  const size_t smem = (size_t)kGatherWords * 8;
  GLOC_CUDA_TRY(cudaFuncSetAttribute(smem_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* sink = nullptr;
  GLOC_CUDA_TRY(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sm_count(device), iters = 1 << 14;
  smem_gather_kernel<<<grid, kGatherThreads, smem>>>(sink, 1 << 10);        // warm-up
  cudaEventRecord(e0);
  smem_gather_kernel<<<grid, kGatherThreads, smem>>>(sink, iters);
  cudaEventRecord(e1);
  cudaError_t ce = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (ce != cudaSuccess || !(ms > 0.f))
    return fail(GLOC_ERR_CUDA, std::string("gloc_bench_smem_gather: ") + cudaGetErrorString(ce));
  *loads_per_s = (double)grid * kGatherThreads * iters / (ms * 1e-3);
  return GLOC_OK;
}
