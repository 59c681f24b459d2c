// cuda_emu.hpp -- a tiny CUDA execution emulator for tests: runs a __global__ function's source
// on the host, one OS thread per CUDA thread of a block (blocks one after the other), with
// __syncthreads() and warp shuffles as real barriers.  Enough for kernels that use only global
// and shared memory, __syncthreads, __shfl_xor_sync, __ldg and libm -- it checks indexing,
// barrier placement and arithmetic of such kernels without a GPU.  (It says nothing about
// performance, memory coalescing or anything tcgen05/TMA.)
//
// Usage: #include this, then the kernel text with `extern __shared__` rewritten to `extern`,
// define the dynamic shared array as a global, and call emu::launch(grid, block, kernel, args...).
#ifndef CUDA_EMU_HPP_
#define CUDA_EMU_HPP_

#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
inline thread_local dim3 t_thread, t_block;
inline dim3 g_grid, g_blockdim;
inline std::barrier<>* g_block_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_warp_barriers;
inline std::vector<std::vector<uint32_t>> g_warp_xchg;   // [warp][lane]
inline thread_local int t_linear = 0;

template <typename T>
inline T shfl_xor(T v, int lane_mask) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  const int warp = t_linear >> 5, lane = t_linear & 31;
  uint32_t bits;
  __builtin_memcpy(&bits, &v, 4);
  g_warp_xchg[warp][lane] = bits;
  g_warp_barriers[warp]->arrive_and_wait();
  const uint32_t got = g_warp_xchg[warp][lane ^ lane_mask];
  g_warp_barriers[warp]->arrive_and_wait();
  T out;
  __builtin_memcpy(&out, &got, 4);
  return out;
}

template <typename F, typename... Args>
void launch(dim3 grid, dim3 block, F kernel, Args... args) {
  g_grid = grid;
  g_blockdim = block;
  const int n = (int)(block.x * block.y * block.z);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(n);
        g_block_barrier = &bar;
        g_warp_barriers.clear();
        g_warp_xchg.assign((n + 31) / 32, std::vector<uint32_t>(32, 0));
        for (int w = 0; w < (n + 31) / 32; ++w)
          g_warp_barriers.emplace_back(new std::barrier<>(std::min(32, n - 32 * w)));
        std::vector<std::thread> ts;
        for (int t = 0; t < n; ++t)
          ts.emplace_back([&, t] {
            t_linear = t;
            t_thread = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            t_block = dim3(bx, by, bz);
            kernel(args...);
            // a thread that returned early must not hold up later barriers of its block
            bar.arrive_and_drop();
          });
        for (auto& th : ts) th.join();
      }
}
}  // namespace emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define threadIdx emu::t_thread
#define blockIdx emu::t_block
#define blockDim emu::g_blockdim
#define gridDim emu::g_grid
#define __syncthreads() emu::g_block_barrier->arrive_and_wait()
#define __shfl_xor_sync(mask, v, o) emu::shfl_xor((v), (o))
template <typename T>
inline T __ldg(const T* p) { return *p; }
using std::min;
using std::max;

#endif  // CUDA_EMU_HPP_
