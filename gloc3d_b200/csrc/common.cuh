// common.cuh -- shared helpers for libgloc3d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gloc3d.h"

namespace gloc {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define GLOC_CUDA_TRY(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      return ::gloc::fail(GLOC_ERR_CUDA, std::string(#expr) + ": " +               \
                                             cudaGetErrorString(_e));              \
    }                                                                              \
  } while (0)

// RAII device switch: every entry point runs on its handle's device and restores
// the caller's current device on exit.
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// (d2, idx) packed so that unsigned 64-bit order == lexicographic (d2, idx) order.
// d2 is a sum of squares: >= +0, so its IEEE bit pattern orders like the value.
__host__ __device__ inline uint64_t pack_key(float d2, uint32_t idx) {
#ifdef __CUDA_ARCH__
  return ((uint64_t)__float_as_uint(d2) << 32) | idx;
#else
  union { float f; uint32_t u; } c;
  c.f = d2;
  return ((uint64_t)c.u << 32) | idx;
#endif
}
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count(int device);

// Per-device one-time setup (cudaFuncSetAttribute applies to the current device only): true the
// first time it is called on the current device for the given mask, false afterwards.
inline bool first_use_on_current_device(unsigned long long& mask) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;   // unknown: just do the setup
  const unsigned long long bit = 1ull << d;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// Times one kernel class with CUDA events recorded on the launching stream (the bench's
// roofline line needs the dominant kernel's average launch duration, measured live).
struct EventProfiler {
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
  void begin(cudaStream_t s) {
    if (!enabled) return;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
    cudaEventRecord(a, s);
    spans.emplace_back(a, b);
    open_ = true;
  }
  void end(cudaStream_t s) {
    if (!enabled || !open_) return;
    cudaEventRecord(spans.back().second, s);
    open_ = false;
  }
  // Waits for the recorded spans; returns total ms and the number of spans; resets.
  void collect(double* ms, uint64_t* n) {
    double t = 0;
    uint64_t c = 0;
    for (auto& sp : spans) {
      float f = 0.f;
      if (cudaEventSynchronize(sp.second) == cudaSuccess &&
          cudaEventElapsedTime(&f, sp.first, sp.second) == cudaSuccess) {
        t += f;
        ++c;
      }
      cudaEventDestroy(sp.first);
      cudaEventDestroy(sp.second);
    }
    spans.clear();
    *ms = t;
    *n = c;
  }
 private:
  bool open_ = false;
};

}  // namespace gloc
