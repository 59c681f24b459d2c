// tc_emu.hpp -- functional model of the sm_100a asynchronous machinery for tests: see the
// header of gemm_emu_test.cpp.  Included by gemm_emu_test.cpp and encoder_emu_test.cpp, which
// add the emulated PTX wrappers under the names their kernels use.
#ifndef TC_EMU_HPP_
#define TC_EMU_HPP_
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ execution model
struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
};
struct float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct CUtensorMap {   // what the emulated TMA needs to know: a 16-bit tensor of rank <= 4
  const uint16_t* base;
  int rank;
  uint64_t dim[4];      // innermost first
  uint64_t stride[4];   // in elements; stride[0] == 1
  uint32_t box[4];      // box[0] == 64 elements = one 128-byte swizzle row
};
static inline CUtensorMap emu_map_2d(const uint16_t* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return CUtensorMap{base, 2, {cols, rows, 1, 1}, {1, cols, 0, 0}, {64, box_rows, 1, 1}};
}

namespace emu {
constexpr int kMaxThreads = 384;
constexpr size_t kSmemBuf = 256 * 1024;

struct Cta {
  unsigned char* smem = nullptr;              // 1024-aligned, kSmemBuf bytes
  std::vector<float> tmem;                    // [128 lanes][512 columns]
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<std::vector<uint32_t>> warp_x;  // [warp][lane]
};
Cta g_cta[2];
int g_n_cta = 1;                              // CTAs running concurrently (cluster size)
std::unique_ptr<std::barrier<>> g_cluster_bar;
Dim3 g_grid, g_blockdim;
thread_local Dim3 t_thread, t_block;
thread_local int t_rank = 0, t_tid = 0;
std::mutex g_bar_mu;
std::atomic<uint64_t> g_progress{0};
std::atomic<bool> g_done{false};

struct Waiting {   // for the watchdog's report
  std::atomic<uint32_t> addr{0};
  std::atomic<int> parity{-1};
};
Waiting g_waiting[2][kMaxThreads];

struct BarState {   // lives in the 8 bytes of the mbarrier object in shared memory
  int32_t tx;
  uint16_t pending;
  uint16_t count_phase;   // count << 1 | phase
};
static_assert(sizeof(BarState) == 8, "mbarrier state");

inline uint32_t addr_of(const void* p) {   // shared::cta window address, CTA rank in bit 24
  return ((uint32_t)t_rank << 24) | (uint32_t)((const unsigned char*)p - g_cta[t_rank].smem);
}
inline BarState* bar_at(uint32_t cluster_addr) {
  return reinterpret_cast<BarState*>(g_cta[(cluster_addr >> 24) & 1].smem + (cluster_addr & 0xFFFFFF));
}
inline void settle(BarState* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->count_phase ^= 1;
    b->pending = b->count_phase >> 1;
  }
}
std::vector<uint32_t> g_bars;   // every initialised barrier of the running cluster
inline void bar_init(uint32_t a, uint32_t count) {
  std::lock_guard<std::mutex> l(g_bar_mu);
  g_bars.push_back(a);
  BarState* b = bar_at(a);
  b->tx = 0;
  b->pending = (uint16_t)count;
  b->count_phase = (uint16_t)(count << 1);
}
inline void bar_arrive(uint32_t a, int32_t expect_bytes) {
  std::lock_guard<std::mutex> l(g_bar_mu);
  BarState* b = bar_at(a);
  if (b->pending == 0) {
    std::fprintf(stderr, "emu: arrive on a barrier with no pending arrivals (addr %08x)\n", a);
    std::abort();
  }
  b->tx += expect_bytes;
  b->pending--;
  settle(b);
  g_progress++;
}
inline void bar_complete_tx(uint32_t a, int32_t bytes) {
  std::lock_guard<std::mutex> l(g_bar_mu);
  BarState* b = bar_at(a);
  b->tx -= bytes;
  settle(b);
  g_progress++;
}
inline bool bar_try_wait(uint32_t a, uint32_t parity) {
  std::lock_guard<std::mutex> l(g_bar_mu);
  return (uint32_t)(bar_at(a)->count_phase & 1) != (parity & 1);
}

template <typename T>
inline T warp_exchange_reduce(T v, T (*op)(T, T)) {   // all 32 lanes call; everyone gets the reduction
  static_assert(sizeof(T) == 4, "32-bit values");
  Cta& c = g_cta[t_rank];
  const int warp = t_tid >> 5, lane = t_tid & 31;
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  c.warp_x[warp][lane] = bits;
  c.warp_bar[warp]->arrive_and_wait();
  T acc;
  std::memcpy(&acc, &c.warp_x[warp][0], 4);
  for (int i = 1; i < 32; ++i) {
    T o;
    std::memcpy(&o, &c.warp_x[warp][i], 4);
    acc = op(acc, o);
  }
  c.warp_bar[warp]->arrive_and_wait();
  return acc;
}
inline int op_or(int a, int b) { return a | b; }
inline int op_max(int a, int b) { return a > b ? a : b; }

inline uint32_t swz(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }   // 128B swizzle on the address

inline float half_to_float(uint16_t h) {
  const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31, m = h & 1023;
  uint32_t bits;
  if (e == 0) {
    if (m == 0) {
      bits = s;
    } else {   // subnormal
      int sh = 0;
      uint32_t mm = m;
      while (!(mm & 1024)) { mm <<= 1; ++sh; }
      bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023) << 13);
    }
  } else if (e == 31) {
    bits = s | 0x7F800000u | (m << 13);
  } else {
    bits = s | ((e + 112) << 23) | (m << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

inline uint32_t tma_box_bytes(const CUtensorMap* m) { return m->box[0] * m->box[1] * m->box[2] * m->box[3] * 2; }
// box -> shared memory, innermost dimension fastest, 128-byte rows, 128B swizzle; elements whose
// coordinates fall outside the tensor (also below zero) are written as zeros
inline void tma_copy(void* smem_dst, const CUtensorMap* map, const int (&c)[4]) {
  Cta& cta = g_cta[t_rank];
  const uint32_t dst = (uint32_t)((unsigned char*)smem_dst - cta.smem);
  if ((dst & 1023u) || map->box[0] != 64) {
    std::fprintf(stderr, "emu: TMA destination %x not 1024-byte aligned or box[0] != 64\n", dst);
    std::abort();
  }
  uint32_t row = 0;
  for (uint32_t i3 = 0; i3 < map->box[3]; ++i3)
    for (uint32_t i2 = 0; i2 < map->box[2]; ++i2)
      for (uint32_t i1 = 0; i1 < map->box[1]; ++i1, ++row)
        for (uint32_t i0 = 0; i0 < 64; ++i0) {
          const long long x[4] = {(long long)c[0] + i0, (long long)c[1] + i1, (long long)c[2] + i2, (long long)c[3] + i3};
          bool in = true;
          uint64_t off = 0;
          for (int d = 0; d < 4; ++d) {
            in = in && x[d] >= 0 && (uint64_t)x[d] < map->dim[d];
            off += (uint64_t)(x[d] < 0 ? 0 : x[d]) * map->stride[d];
          }
          const uint16_t v = in ? map->base[off] : (uint16_t)0;
          std::memcpy(cta.smem + swz(dst + row * 128 + i0 * 2), &v, 2);
        }
}

inline float operand(const Cta& c, uint32_t start, int row, int k) {   // K-major SW128 tile, SBO = 1024
  uint16_t v;
  std::memcpy(&v, c.smem + swz(start + (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 + (uint32_t)k * 2), 2);
  return half_to_float(v);
}

// D[lanes][cols] (+)= A[M x 16] * B[N x 16]^T for one UMMA of kind::f16
inline void umma(bool pair, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t a0 = (uint32_t)(desc_a & 0x3FFF) << 4, b0 = (uint32_t)(desc_b & 0x3FFF) << 4;
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  const int col0 = (int)(tmem_d & 0xFFFF);
  if ((tmem_d >> 16) != 0 || M != (pair ? 256 : 128) || col0 + N > 512) {
    std::fprintf(stderr, "emu: unexpected MMA shape/address M=%d N=%d tmem=%08x\n", M, N, tmem_d);
    std::abort();
  }
  const int n_per_cta = pair ? N / 2 : N;
  std::vector<float> B((size_t)N * 16);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k)
      B[(size_t)n * 16 + k] = operand(g_cta[pair ? n / n_per_cta : t_rank], b0, n % n_per_cta, k);
  for (int m = 0; m < M; ++m) {
    Cta& c = g_cta[pair ? m / 128 : t_rank];
    float arow[16];
    for (int k = 0; k < 16; ++k) arow[k] = operand(c, a0, m % 128, k);
    float* d = c.tmem.data() + (size_t)(m % 128) * 512 + col0;
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < 16; ++k) s += arow[k] * B[(size_t)n * 16 + k];
      d[n] = accumulate ? d[n] + s : s;
    }
  }
  g_progress++;
}
// ---- asynchronous engine (GLOC_EMU_ASYNC=<seed>): TMA loads complete in any order and late,
// MMAs and commits execute in issue order but detached from the issuing thread, as on hardware.
// Off: every asynchronous operation completes at issue (one legal timing).
struct Engine {
  bool on = false;
  std::mutex mu;
  std::deque<std::function<void()>> mma;      // in order
  std::vector<std::function<void()>> tma;     // any order
  std::mt19937 rng{1};
  std::thread th;
  std::atomic<bool> stop{false};
  std::atomic<int> in_flight{0};
  void submit(bool is_mma, std::function<void()> op) {
    if (!on) {
      op();
      return;
    }
    std::lock_guard<std::mutex> l(mu);
    ++in_flight;
    if (is_mma) mma.push_back(std::move(op));
    else tma.push_back(std::move(op));
  }
  void loop() {
    while (!stop.load()) {
      std::function<void()> op;
      {
        std::lock_guard<std::mutex> l(mu);
        const bool take_mma = !mma.empty() && (tma.empty() || (rng() & 1));
        if (take_mma) {
          op = std::move(mma.front());
          mma.pop_front();
        } else if (!tma.empty()) {
          const size_t i = rng() % tma.size();
          op = std::move(tma[i]);
          tma.erase(tma.begin() + (long)i);
        }
        if (op && (rng() % 4) == 0) {   // sometimes let the issuing threads run ahead first
          if (take_mma) mma.push_front(std::move(op));
          else tma.push_back(std::move(op));
          op = nullptr;
        }
      }
      if (op) {
        op();
        --in_flight;
      } else {
        std::this_thread::yield();
      }
    }
  }
  void start(unsigned seed) {
    on = true;
    rng.seed(seed);
    th = std::thread([this] { loop(); });
  }
  void finish() {
    if (!on) return;
    stop = true;
    th.join();
  }
};
Engine g_engine;
}  // namespace emu

// ------------------------------------------------------------------ CUDA surface of the kernel text
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define threadIdx emu::t_thread
#define blockIdx emu::t_block
#define blockDim emu::g_blockdim
#define gridDim emu::g_grid
using std::max;
using std::min;

static inline void __syncthreads() { emu::g_cta[emu::t_rank].block_bar->arrive_and_wait(); }
static inline void __syncwarp() { emu::g_cta[emu::t_rank].warp_bar[emu::t_tid >> 5]->arrive_and_wait(); }
static inline int __any_sync(unsigned, int pred) { return emu::warp_exchange_reduce<int>(pred ? 1 : 0, emu::op_or); }
static inline int __reduce_max_sync(unsigned, int v) { return emu::warp_exchange_reduce<int>(v, emu::op_max); }
static inline float __uint_as_float(unsigned u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline unsigned __float_as_uint(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  return u;
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicMin(unsigned* p, unsigned v) {
  std::atomic_ref<unsigned> r(*p);
  unsigned cur = r.load();
  while (v < cur && !r.compare_exchange_weak(cur, v)) {
  }
  return cur;
}

// ---- what the K1 / K3 kernels need on top
template <typename T>
static inline T emu_shfl(T v, int src_lane) {   // every lane of the warp calls; src_lane per caller
  static_assert(sizeof(T) == 4, "32-bit shuffles");
  emu::Cta& c = emu::g_cta[emu::t_rank];
  const int warp = emu::t_tid >> 5, lane = emu::t_tid & 31;
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  c.warp_x[warp][lane] = bits;
  c.warp_bar[warp]->arrive_and_wait();
  const uint32_t got = (src_lane >= 0 && src_lane < 32) ? c.warp_x[warp][src_lane] : bits;
  c.warp_bar[warp]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, 4);
  return out;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (emu::t_tid & 31) ^ m); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = emu::t_tid & 31;
  return emu_shfl(v, lane - d >= 0 ? lane - d : lane);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src & 31); }
template <typename T, typename U>
static inline T atomicAdd(T* p, U v) {
  return std::atomic_ref<T>(*p).fetch_add((T)v);
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  std::atomic_ref<unsigned> r(*p);
  unsigned cur = r.load();
  while (v > cur && !r.compare_exchange_weak(cur, v)) {
  }
  return cur;
}
static inline float __fadd_rn(float a, float b) { return a + b; }   // built with -ffp-contract=off
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline float __int_as_float(int i) {
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}
#define __align__(n) alignas(n)
struct float2 {
  float x, y;
};
struct __half {
  uint16_t bits;
};
struct __half2 {
  __half x, y;
};
static inline uint16_t float_to_half_rn(float f) {   // IEEE round to nearest even, like cvt.rn.f16.f32
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t a = u & 0x7FFFFFFFu;
  if (a >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (a > 0x7F800000u ? 0x200u : 0));
  if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);            // rounds to infinity
  if (a < 0x33000001u) return (uint16_t)sign;                           // below half of the smallest subnormal
  int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
  int shift = e >= -14 ? 13 : 13 + (-14 - e);                           // subnormal halves lose more bits
  uint32_t half_m = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1))) ++half_m;
  uint32_t h = e >= -14 ? (((uint32_t)(e + 15) << 10) + (half_m - 0x400u)) : half_m;   // carries propagate
  return (uint16_t)(sign | h);
}
static inline __half2 __floats2half2_rn(float a, float b) {
  return __half2{__half{float_to_half_rn(a)}, __half{float_to_half_rn(b)}};
}
static inline float2 __half22float2(__half2 h) {
  return float2{emu::half_to_float(h.x.bits), emu::half_to_float(h.y.bits)};
}


#endif  // TC_EMU_HPP_
