// Functional emulation of the shortlist GEMM kernel (gloc3d_b200/csrc/knn_shortlist.cu,
// knn_shortlist_gemm_kernel<kPair>) on the host: the kernel's own source text (regions marked
// [emu-*] in the .cu, extracted at build time into _gemm_*.inc) runs with one OS thread per CUDA
// thread while this file stands in for the hardware it talks to:
//   * mbarriers (init / arrive / expect_tx / complete_tx / try_wait.parity), also across the two
//     CTAs of a cluster (shared::cluster addresses = CTA rank in bit 24 + offset);
//   * TMA tile loads (cp.async.bulk.tensor.2d, 128B swizzle, out-of-bounds rows read as zero,
//     bytes counted on the given barrier -- the leader's for the cta_group::2 form);
//   * tcgen05.mma kind::f16 from shared-memory descriptors into tensor memory (cta_group::1:
//     M = 128; cta_group::2: M = 256 over both CTAs, N = 256 = leader's rows then peer's rows),
//     tcgen05.commit (multicast form: both CTAs), tcgen05.ld 32x32b.x32;
//   * warp votes / reductions, __syncthreads, the cluster barrier, atomicMin.
// Asynchronous operations complete at issue, or -- GLOC_EMU_ASYNC=<seed> -- late and (TMA) out of
// order through a background engine.  What this checks: the
// barrier protocol terminates (a watchdog reports which barriers are being waited on when
// nothing moves), tile coordinates, accumulator addressing, and the candidate lists the epilogue
// writes.  What it cannot check: anything about real hardware behaviour that the kernel's author
// and this model misunderstand in the same way (notably the order of the two CTAs' rows in the
// N dimension of a cta_group::2 MMA), and performance.
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

struct CUtensorMap {   // what the emulated TMA needs to know
  const uint16_t* base;
  uint64_t rows, dim;
  uint32_t box_rows;
};

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

inline void tma_copy(void* smem_dst, const CUtensorMap* map, int c_inner, int c_outer) {
  Cta& c = g_cta[t_rank];
  const uint32_t dst = (uint32_t)((unsigned char*)smem_dst - c.smem);
  if (dst & 1023u) {
    std::fprintf(stderr, "emu: TMA destination %x is not 1024-byte aligned\n", dst);
    std::abort();
  }
  for (uint32_t r = 0; r < map->box_rows; ++r)
    for (uint32_t kc = 0; kc < 64; ++kc) {
      const uint64_t row = (uint64_t)((long long)c_outer + r), col = (uint64_t)((long long)c_inner + kc);
      const bool in = c_outer + (long long)r >= 0 && row < map->rows && col < map->dim;
      const uint16_t v = in ? map->base[row * map->dim + col] : (uint16_t)0;
      std::memcpy(c.smem + swz(dst + r * 128 + kc * 2), &v, 2);
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

namespace gloc {
namespace {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;   // csrc/common.cuh
inline uint64_t pack_key(float d2, uint32_t idx) { return ((uint64_t)__float_as_uint(d2) << 32) | idx; }

// ---- the kernel's PTX wrappers, emulated (same names and signatures as in knn_shortlist.cu)
inline uint32_t smem_u32(const void* p) { return emu::addr_of(p); }
inline void mbar_init(uint64_t* bar, uint32_t count) { emu::bar_init(smem_u32(bar), count); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu::bar_arrive(smem_u32(bar), (int32_t)bytes); }
inline void mbar_arrive(uint64_t* bar) { emu::bar_arrive(smem_u32(bar), 0); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  emu::Waiting& w = emu::g_waiting[emu::t_rank][emu::t_tid];
  w.addr = a;
  w.parity = (int)parity;
  while (!emu::bar_try_wait(a, parity)) std::this_thread::yield();
  w.parity = -1;
}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void tcgen05_fence_before() {}
inline void tcgen05_fence_after() {}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  const int rank = emu::t_rank;
  const uint32_t b = smem_u32(bar);
  const CUtensorMap m = *map;
  emu::g_engine.submit(false, [=] {
    emu::t_rank = rank;
    emu::tma_copy(smem_dst, &m, c_inner, c_outer);
    emu::bar_complete_tx(b, (int32_t)(m.box_rows * 128));
  });
}
inline uint32_t cluster_ctarank() { return (uint32_t)emu::t_rank; }
inline void cluster_sync_all() { emu::g_cluster_bar->arrive_and_wait(); }
inline uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) { return (smem_addr & 0xFFFFFFu) | (rank << 24); }
inline void mbar_arrive_cluster(uint32_t cluster_addr) { emu::bar_arrive(cluster_addr, 0); }
inline void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c_inner,
                             int c_outer) {
  const int rank = emu::t_rank;
  const CUtensorMap m = *map;
  emu::g_engine.submit(false, [=] {
    emu::t_rank = rank;
    emu::tma_copy(smem_dst, &m, c_inner, c_outer);
    emu::bar_complete_tx(bar_cluster_addr, (int32_t)(m.box_rows * 128));
  });
}
// a commit arrives once every MMA issued before it has executed (the engine's MMA queue is in order)
inline void tcgen05_commit(uint64_t* bar) {
  const uint32_t a = smem_u32(bar);
  emu::g_engine.submit(true, [=] { emu::bar_arrive(a, 0); });
}
inline void tcgen05_commit_pair(uint64_t* bar) {
  const uint32_t a = smem_u32(bar);
  emu::g_engine.submit(true, [=] {
    emu::bar_arrive(map_to_cta(a, 0), 0);
    emu::bar_arrive(map_to_cta(a, 1), 0);
  });
}
inline void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  const int rank = emu::t_rank;
  emu::g_engine.submit(true, [=] {
    emu::t_rank = rank;
    emu::umma(false, tmem_d, da, db, idesc, acc);
  });
}
inline void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  const int rank = emu::t_rank;
  emu::g_engine.submit(true, [=] {
    emu::t_rank = rank;
    emu::umma(true, tmem_d, da, db, idesc, acc);
  });
}
inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  const emu::Cta& c = emu::g_cta[emu::t_rank];
  const int lane = (int)(taddr >> 16) + (emu::t_tid & 31), col = (int)(taddr & 0xFFFF);
  if (lane >= 128 || col + 32 > 512) {
    std::fprintf(stderr, "emu: tcgen05.ld outside tensor memory (lane %d col %d)\n", lane, col);
    std::abort();
  }
  std::memcpy(v, c.tmem.data() + (size_t)lane * 512 + col, 128);
}
inline void tmem_ld_wait() {}

thread_local unsigned char* t_smem_raw = nullptr;

#include "_gemm_a.inc"
#include "_gemm_b.inc"
#include "_gemm_k1.inc"
#include "_gemm_c.inc"
#include "_gemm_k3.inc"

}  // namespace
}  // namespace gloc

// ------------------------------------------------------------------ harness
using namespace gloc;

template <typename T>
static std::vector<T> read_vec(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
  return v;
}

template <bool kPair>
static void run_grid(int n_workers, const CUtensorMap& map_q, const CUtensorMap& map_db, const GemmArgs& a) {
  const int per = kPair ? 2 : 1;
  emu::g_n_cta = per;
  emu::g_grid.x = (unsigned)(n_workers * per);
  emu::g_blockdim.x = (unsigned)kThreads;
  for (int wk = 0; wk < n_workers; ++wk) {   // workers one after the other: a legal schedule
    for (int r = 0; r < per; ++r) {
      emu::Cta& c = emu::g_cta[r];
      if (!c.smem) c.smem = static_cast<unsigned char*>(std::aligned_alloc(1024, emu::kSmemBuf));
      std::memset(c.smem, 0xCD, emu::kSmemBuf);
      c.tmem.assign((size_t)128 * 512, NAN);
      c.block_bar.reset(new std::barrier<>(kThreads));
      c.warp_bar.clear();
      for (int w = 0; w < kThreads / 32; ++w) c.warp_bar.emplace_back(new std::barrier<>(32));
      c.warp_x.assign(kThreads / 32, std::vector<uint32_t>(32, 0));
    }
    emu::g_cluster_bar.reset(new std::barrier<>(kThreads * per));
    std::vector<std::thread> ts;
    for (int r = 0; r < per; ++r)
      for (int t = 0; t < kThreads; ++t)
        ts.emplace_back([&, r, t] {
          emu::t_rank = r;
          emu::t_tid = t;
          emu::t_thread.x = (unsigned)t;
          emu::t_block.x = (unsigned)(wk * per + r);
          t_smem_raw = emu::g_cta[r].smem + 16;   // dynamic shared memory does not start 1024-aligned
          knn_shortlist_gemm_kernel<kPair>(map_q, map_db, a);
        });
    for (auto& th : ts) th.join();
    if (emu::g_engine.in_flight.load() != 0) {   // a CTA must not exit under its own TMA loads / MMAs
      std::fprintf(stderr, "emu: %d asynchronous operations still in flight when the CTAs exited\n",
                   emu::g_engine.in_flight.load());
      std::_Exit(5);
    }
    // every byte that was announced has arrived and vice versa
    for (uint32_t a : emu::g_bars)
      if (emu::bar_at(a)->tx != 0) {
        std::fprintf(stderr, "emu: barrier %08x ends with transaction count %d\n", a, emu::bar_at(a)->tx);
        std::exit(4);
      }
    emu::g_bars.clear();
  }
}

// ordinary kernels (K1, K3): blocks one after the other, block_threads OS threads each
template <typename F>
static void launch_blocks(unsigned grid_x, unsigned block_threads, size_t dyn_smem, F body) {
  emu::g_n_cta = 1;
  emu::g_grid.x = grid_x;
  emu::g_blockdim.x = block_threads;
  emu::Cta& c = emu::g_cta[0];
  if (!c.smem) c.smem = static_cast<unsigned char*>(std::aligned_alloc(1024, emu::kSmemBuf));
  if (dyn_smem + 16 > emu::kSmemBuf) std::abort();
  for (unsigned bx = 0; bx < grid_x; ++bx) {
    c.block_bar.reset(new std::barrier<>(block_threads));
    c.warp_bar.clear();
    for (unsigned w = 0; w < (block_threads + 31) / 32; ++w)
      c.warp_bar.emplace_back(new std::barrier<>(std::min(32u, block_threads - 32 * w)));
    c.warp_x.assign((block_threads + 31) / 32, std::vector<uint32_t>(32, 0));
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block_threads; ++t)
      ts.emplace_back([&, t] {
        emu::t_rank = 0;
        emu::t_tid = (int)t;
        emu::t_thread.x = t;
        emu::t_block.x = bx;
        t_smem_raw = c.smem + 16;
        body();
        c.block_bar->arrive_and_drop();                  // early returns must not block the others
        c.warp_bar[t >> 5]->arrive_and_drop();
      });
    for (auto& th : ts) th.join();
  }
}

// mode 1: the whole shortlist path from float32 rows -- K1 (stats, FP16 copy, query prep), K2, K3 --
// with the launch geometry of shortlist_query(); writes the top-k and the overflow count.
static int run_full(FILE* f, const std::vector<int32_t>& h, const char* out_path) {
  const int nq = h[0], n_rows = h[1], n_pad = h[2], dim = h[3], k = h[4], cap = h[5], n_ranges = h[6],
            tiles_per_range = h[7], pair = h[8], workers = h[9];
  const std::vector<float> q = read_vec<float>(f, (size_t)nq * dim), db = read_vec<float>(f, (size_t)n_rows * dim);
  std::fclose(f);
  // K1, database
  std::vector<uint16_t> db_h((size_t)n_pad * dim, 0xCDCD), q_h((size_t)nq * dim, 0xCDCD);
  std::vector<float> xn((size_t)n_pad, NAN), qn((size_t)nq), qe((size_t)nq), qinv((size_t)nq);
  std::vector<unsigned> stats(4, 0u);
  const int wpb = 8;
  launch_blocks((unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, [&] {
    knn_db_stats_kernel(db.data(), n_rows, dim, xn.data(), stats.data(), stats.data() + 1);
  });
  const float scale_x = pow2_scale_for(__uint_as_float(stats[1]));
  launch_blocks((unsigned)((n_pad + wpb - 1) / wpb), wpb * 32, 0, [&] {
    knn_db_convert_kernel(db.data(), n_rows, n_pad, dim, scale_x, reinterpret_cast<__half*>(db_h.data()), xn.data(),
                          stats.data() + 2);
  });
  // K1, queries
  launch_blocks((unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, [&] {
    knn_query_prep_kernel(q.data(), nq, dim, reinterpret_cast<__half*>(q_h.data()), qn.data(), qe.data(), qinv.data());
  });
  // K2
  const int n_qtiles = (nq + BM - 1) / BM;
  const size_t lists = (size_t)nq * n_ranges * 2;
  std::vector<unsigned> thr((size_t)nq, 0xFF800000u), cand_g(lists * cap, 0xFFFFFFFFu), unit_cnt(lists, 0xFFFFFFFFu);
  std::vector<float> eps2((size_t)nq, -1.f);
  std::vector<float4> cand_v(lists * cap * 2, float4{NAN, NAN, NAN, NAN});
  GemmArgs g;
  g.nq = nq;
  g.n_qtiles = pair ? (n_qtiles + 1) / 2 : n_qtiles;
  g.n_ranges = n_ranges;
  g.tiles_per_range = tiles_per_range;
  g.n_kb = dim / BK;
  g.k = k;
  g.cap = cap;
  g.r_big = 1;
  g.n_rows = n_rows;
  g.xn = xn.data();
  g.qn = qn.data();
  g.qe = qe.data();
  g.qinv = qinv.data();
  g.inv_sx = 1.f / scale_x;
  g.max_norm2_bits = stats.data();
  g.max_dx2_bits = stats.data() + 2;
  g.thr_ord = thr.data();
  g.eps2 = eps2.data();
  g.cand_g = cand_g.data();
  g.cand_v = cand_v.data();
  g.unit_cnt = unit_cnt.data();
  const CUtensorMap map_q{q_h.data(), (uint64_t)nq, (uint64_t)dim, (uint32_t)BM};
  const CUtensorMap map_db{db_h.data(), (uint64_t)n_rows, (uint64_t)dim, (uint32_t)(pair ? BN / 2 : BN)};
  const int n_workers = std::min(g.n_qtiles * n_ranges, workers);
  if (pair) run_grid<true>(n_workers, map_q, map_db, g);
  else run_grid<false>(n_workers, map_q, map_db, g);
  emu::g_engine.finish();
  // K3
  std::vector<uint64_t> out_idx((size_t)nq * k, 0);
  std::vector<float> out_d2((size_t)nq * k, NAN);
  std::vector<int> ovf_list((size_t)nq, -1), ovf_count(1, 0);
  std::vector<unsigned long long> rows_ctr(2, 0);
  RerankArgs r;
  r.db = db.data();
  r.q = q.data();
  r.nq = nq;
  r.dim = dim;
  r.k = k;
  r.n_ranges = n_ranges * 2;
  r.cap = cap;
  r.cand_g = cand_g.data();
  r.cand_v = reinterpret_cast<const float*>(cand_v.data());
  r.unit_cnt = unit_cnt.data();
  r.thr_ord = thr.data();
  r.eps2 = eps2.data();
  r.offset = 0;
  r.out_idx = out_idx.data();
  r.out_d2 = out_d2.data();
  r.overflow_list = ovf_list.data();
  r.overflow_count = ovf_count.data();
  r.rows_reranked = rows_ctr.data();
  const size_t rr_smem = std::max((size_t)kCandMax * 8, (size_t)32 * (dim / 4 + 1) * 4) + (size_t)4 * 256 * 4 +
                         (size_t)kFinalMax * 8 + (size_t)dim * 4;
  launch_blocks((unsigned)nq, kRerankThreads, rr_smem, [&] { knn_shortlist_rerank_kernel(r); });
  FILE* o = std::fopen(out_path, "wb");
  if (!o) return 2;
  std::fwrite(out_idx.data(), 8, out_idx.size(), o);
  std::fwrite(out_d2.data(), 4, out_d2.size(), o);
  std::fwrite(ovf_count.data(), 4, 1, o);
  std::fwrite(rows_ctr.data(), 8, 2, o);
  std::fclose(o);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  // header: nq, n_rows, n_pad, dim, k, cap, n_ranges, tiles_per_range, pair, workers
  const std::vector<int32_t> h = read_vec<int32_t>(f, 11);
  const int nq = h[0], n_rows = h[1], n_pad = h[2], dim = h[3], k = h[4], cap = h[5], n_ranges = h[6],
            tiles_per_range = h[7], pair = h[8], workers = h[9], mode = h[10];
  if (mode == 1) {
    std::thread([] {   // the same watchdog, detached
      uint64_t last = emu::g_progress.load();
      for (int idle = 0;; ) {
        std::this_thread::sleep_for(std::chrono::seconds(1));
        const uint64_t now = emu::g_progress.load();
        idle = now == last ? idle + 1 : 0;
        last = now;
        if (idle >= 120) {
          std::fprintf(stderr, "emu: DEADLOCK -- no barrier progress for 120 s\n");
          std::_Exit(3);
        }
      }
    }).detach();
    if (const char* e = std::getenv("GLOC_EMU_ASYNC")) emu::g_engine.start((unsigned)std::atoi(e));
    return run_full(f, h, argv[2]);
  }
  const float inv_sx = read_vec<float>(f, 1)[0];
  const std::vector<unsigned> stats = read_vec<unsigned>(f, 3);          // max ||x||^2, -, max ||dx||^2 (float bits)
  const std::vector<uint16_t> q_h = read_vec<uint16_t>(f, (size_t)nq * dim);
  const std::vector<uint16_t> db_h = read_vec<uint16_t>(f, (size_t)n_pad * dim);
  const std::vector<float> xn = read_vec<float>(f, (size_t)n_pad), qn = read_vec<float>(f, (size_t)nq),
                           qe = read_vec<float>(f, (size_t)nq), qinv = read_vec<float>(f, (size_t)nq);
  std::fclose(f);

  const int n_qtiles = (nq + BM - 1) / BM;
  const size_t lists = (size_t)nq * n_ranges * 2;
  std::vector<unsigned> thr((size_t)nq, 0xFF800000u), cand_g(lists * cap, 0xFFFFFFFFu), unit_cnt(lists, 0xFFFFFFFFu);
  std::vector<float> eps2((size_t)nq, -1.f);
  std::vector<float4> cand_v(lists * cap * 2, float4{NAN, NAN, NAN, NAN});

  GemmArgs g;
  g.nq = nq;
  g.n_qtiles = pair ? (n_qtiles + 1) / 2 : n_qtiles;
  g.n_ranges = n_ranges;
  g.tiles_per_range = tiles_per_range;
  g.n_kb = dim / BK;
  g.k = k;
  g.cap = cap;
  g.r_big = 1;
  g.n_rows = n_rows;
  g.xn = xn.data();
  g.qn = qn.data();
  g.qe = qe.data();
  g.qinv = qinv.data();
  g.inv_sx = inv_sx;
  g.max_norm2_bits = stats.data();
  g.max_dx2_bits = stats.data() + 2;
  g.thr_ord = thr.data();
  g.eps2 = eps2.data();
  g.cand_g = cand_g.data();
  g.cand_v = cand_v.data();
  g.unit_cnt = unit_cnt.data();
  const CUtensorMap map_q{q_h.data(), (uint64_t)nq, (uint64_t)dim, (uint32_t)BM};
  const CUtensorMap map_db{db_h.data(), (uint64_t)n_rows, (uint64_t)dim, (uint32_t)(pair ? BN / 2 : BN)};
  const int n_units = g.n_qtiles * n_ranges;
  const int n_workers = std::min(n_units, workers);

  std::thread watchdog([] {   // no barrier traffic for 20 s: report who waits on what, give up
    uint64_t last = emu::g_progress.load();
    int idle = 0;
    while (!emu::g_done.load()) {
      std::this_thread::sleep_for(std::chrono::milliseconds(500));
      const uint64_t now = emu::g_progress.load();
      idle = now == last ? idle + 1 : 0;
      last = now;
      if (idle >= 40) {
        std::fprintf(stderr, "emu: DEADLOCK -- no barrier progress for 20 s.  Waiting threads:\n");
        for (int r = 0; r < 2; ++r)
          for (int t = 0; t < emu::kMaxThreads; ++t)
            if (emu::g_waiting[r][t].parity.load() >= 0)
              std::fprintf(stderr, "  cta %d thread %3d (warp %2d): barrier %08x parity %d\n", r, t, t >> 5,
                           emu::g_waiting[r][t].addr.load(), emu::g_waiting[r][t].parity.load());
        std::_Exit(3);
      }
    }
  });
  if (const char* e = std::getenv("GLOC_EMU_ASYNC")) emu::g_engine.start((unsigned)std::atoi(e));
  if (pair) run_grid<true>(n_workers, map_q, map_db, g);
  else run_grid<false>(n_workers, map_q, map_db, g);
  emu::g_engine.finish();
  emu::g_done = true;
  watchdog.join();

  FILE* o = std::fopen(argv[2], "wb");
  if (!o) return 2;
  std::fwrite(thr.data(), 4, thr.size(), o);
  std::fwrite(eps2.data(), 4, eps2.size(), o);
  std::fwrite(unit_cnt.data(), 4, unit_cnt.size(), o);
  std::fwrite(cand_g.data(), 4, cand_g.size(), o);
  std::fwrite(cand_v.data(), 16, cand_v.size(), o);
  std::fclose(o);
  return 0;
}
