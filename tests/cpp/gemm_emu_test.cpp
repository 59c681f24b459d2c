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
#include "tc_emu.hpp"

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
    emu::tma_copy(smem_dst, &m, {c_inner, c_outer, 0, 0});
    emu::bar_complete_tx(b, (int32_t)emu::tma_box_bytes(&m));
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
    emu::tma_copy(smem_dst, &m, {c_inner, c_outer, 0, 0});
    emu::bar_complete_tx(bar_cluster_addr, (int32_t)emu::tma_box_bytes(&m));
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

// ordinary kernels (K1, K3): blocks one after the other on ONE team of block_threads OS threads
// (a barrier between blocks rebuilds the per-block barriers; spawning a team per block is what
// made this slow)
template <typename F>
static void launch_blocks(unsigned grid_x, unsigned block_threads, size_t dyn_smem, F body) {
  emu::g_n_cta = 1;
  emu::g_grid.x = grid_x;
  emu::g_blockdim.x = block_threads;
  emu::Cta& c = emu::g_cta[0];
  if (!c.smem) c.smem = static_cast<unsigned char*>(std::aligned_alloc(1024, emu::kSmemBuf));
  if (dyn_smem + 16 > emu::kSmemBuf) std::abort();
  const unsigned n_warps = (block_threads + 31) / 32;
  auto rebuild = [&c, block_threads, n_warps]() noexcept {
    c.block_bar.reset(new std::barrier<>(block_threads));
    c.warp_bar.clear();
    for (unsigned w = 0; w < n_warps; ++w)
      c.warp_bar.emplace_back(new std::barrier<>(std::min(32u, block_threads - 32 * w)));
  };
  rebuild();
  c.warp_x.assign(n_warps, std::vector<uint32_t>(32, 0));
  std::barrier between_blocks(block_threads, rebuild);
  std::vector<std::thread> ts;
  for (unsigned t = 0; t < block_threads; ++t)
    ts.emplace_back([&, t] {
      emu::t_rank = 0;
      emu::t_tid = (int)t;
      emu::t_thread.x = t;
      t_smem_raw = c.smem + 16;
      for (unsigned bx = 0; bx < grid_x; ++bx) {
        emu::t_block.x = bx;
        body();
        c.block_bar->arrive_and_drop();                  // early returns must not block the others
        c.warp_bar[t >> 5]->arrive_and_drop();
        between_blocks.arrive_and_wait();
      }
    });
  for (auto& th : ts) th.join();
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
  const CUtensorMap map_q = emu_map_2d(q_h.data(), (uint64_t)nq, (uint64_t)dim, (uint32_t)BM);
  const CUtensorMap map_db = emu_map_2d(db_h.data(), (uint64_t)n_rows, (uint64_t)dim, (uint32_t)(pair ? BN / 2 : BN));
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
  std::fwrite(ovf_list.data(), 4, ovf_list.size(), o);
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
  const CUtensorMap map_q = emu_map_2d(q_h.data(), (uint64_t)nq, (uint64_t)dim, (uint32_t)BM);
  const CUtensorMap map_db = emu_map_2d(db_h.data(), (uint64_t)n_rows, (uint64_t)dim, (uint32_t)(pair ? BN / 2 : BN));
  const int n_units = g.n_qtiles * n_ranges;
  const int n_workers = std::min(n_units, workers);

  std::thread watchdog([] {   // no barrier traffic for 60 s: report who waits on what, give up
    uint64_t last = emu::g_progress.load();
    int idle = 0;
    while (!emu::g_done.load()) {
      std::this_thread::sleep_for(std::chrono::milliseconds(500));
      const uint64_t now = emu::g_progress.load();
      idle = now == last ? idle + 1 : 0;
      last = now;
      if (idle >= 120) {
        std::fprintf(stderr, "emu: DEADLOCK -- no barrier progress for 60 s.  Waiting threads:\n");
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
