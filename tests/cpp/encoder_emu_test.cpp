// Runs the kernels of gloc3d_b200/csrc/encoder.cu on the host: the tensor-core convolution against
// the functional mbarrier / TMA / tcgen05 model of tc_emu.hpp (4-D TMA boxes with zero fill =
// the convolution's padding), the two SIMT kernels one OS thread per CUDA thread.
// tests/test_encoder_emulated.py drives it and compares with numpy.
//   encoder_emu_test conv  <in> <out>    one 3x3 convolution layer (BN chosen like the library does)
//   encoder_emu_test conv1 <in> <out>    the folded first layer on a uint8 image
//   encoder_emu_test pool  <in> <out>    2x2 max-pool
#include "tc_emu.hpp"

struct uint4 {
  unsigned x, y, z, w;
};

namespace gloc {
namespace tc {
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
inline void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  const uint32_t b = smem_u32(bar);
  const CUtensorMap m = *map;
  emu::g_engine.submit(false, [=] {
    emu::t_rank = 0;
    emu::tma_copy(dst, &m, {c0, c1, 0, 0});
    emu::bar_complete_tx(b, (int32_t)emu::tma_box_bytes(&m));
  });
}
inline void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const uint32_t b = smem_u32(bar);
  const CUtensorMap m = *map;
  emu::g_engine.submit(false, [=] {
    emu::t_rank = 0;
    emu::tma_copy(dst, &m, {c0, c1, c2, c3});
    emu::bar_complete_tx(b, (int32_t)emu::tma_box_bytes(&m));
  });
}
inline void tcgen05_commit(uint64_t* bar) {
  const uint32_t a = smem_u32(bar);
  emu::g_engine.submit(true, [=] { emu::bar_arrive(a, 0); });
}
inline void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  emu::g_engine.submit(true, [=] {
    emu::t_rank = 0;
    emu::umma(false, tmem_d, da, db, idesc, acc);
  });
}
inline uint64_t make_sw128_desc(uint32_t smem_addr) {   // csrc/tc_ptx.cuh
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t instr_desc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
template <int kCols>
inline void tmem_alloc(uint32_t* slot) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM allocations are powers of two");
  *slot = 0;
}
template <int kCols>
inline void tmem_dealloc(uint32_t) {}
}  // namespace tc

namespace {
thread_local unsigned char* t_smem_raw = nullptr;
#include "_encoder_kernels.inc"
}  // namespace
}  // namespace gloc

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

static void prepare_cta(unsigned threads) {
  emu::Cta& c = emu::g_cta[0];
  if (!c.smem) c.smem = static_cast<unsigned char*>(std::aligned_alloc(1024, emu::kSmemBuf));
  std::memset(c.smem, 0xCD, emu::kSmemBuf);
  c.tmem.assign((size_t)128 * 512, NAN);
  c.block_bar.reset(new std::barrier<>(threads));
  c.warp_bar.clear();
  for (unsigned w = 0; w < (threads + 31) / 32; ++w)
    c.warp_bar.emplace_back(new std::barrier<>(std::min(32u, threads - 32 * w)));
  c.warp_x.assign((threads + 31) / 32, std::vector<uint32_t>(32, 0));
}

template <typename F>
static void launch_blocks(unsigned grid_x, unsigned block_threads, F body) {
  emu::g_n_cta = 1;
  emu::g_grid.x = grid_x;
  emu::g_blockdim.x = block_threads;
  for (unsigned bx = 0; bx < grid_x; ++bx) {
    prepare_cta(block_threads);
    emu::Cta& c = emu::g_cta[0];
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block_threads; ++t)
      ts.emplace_back([&, t] {
        emu::t_rank = 0;
        emu::t_tid = (int)t;
        emu::t_thread.x = t;
        emu::t_block.x = bx;
        t_smem_raw = c.smem + 16;
        body();
        c.block_bar->arrive_and_drop();
        c.warp_bar[t >> 5]->arrive_and_drop();
      });
    for (auto& th : ts) th.join();
    if (emu::g_engine.in_flight.load() != 0) {
      std::fprintf(stderr, "emu: %d asynchronous operations in flight at CTA exit\n", emu::g_engine.in_flight.load());
      std::_Exit(5);
    }
    for (uint32_t a : emu::g_bars)
      if (emu::bar_at(a)->tx != 0) {
        std::fprintf(stderr, "emu: barrier %08x ends with transaction count %d\n", a, emu::bar_at(a)->tx);
        std::exit(4);
      }
    emu::g_bars.clear();
  }
}

template <int BN>
static void run_conv(const CUtensorMap& map_in, const CUtensorMap& map_w, const ConvArgs& a, int workers) {
  const int n_tiles = a.B * (a.H / kEncTileH) * (a.W / kEncTileW) * (a.Cout / BN);
  launch_blocks((unsigned)std::min(n_tiles, workers), kEncThreads, [&] { enc_conv3x3_kernel<BN>(map_in, map_w, a); });
}

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  const std::string mode = argv[1];
  FILE* f = std::fopen(argv[2], "rb");
  if (!f) return 2;
  std::thread([] {
    uint64_t last = emu::g_progress.load();
    for (int idle = 0;;) {
      std::this_thread::sleep_for(std::chrono::seconds(1));
      const uint64_t now = emu::g_progress.load();
      idle = now == last ? idle + 1 : 0;
      last = now;
      if (idle >= 90 && !emu::g_done.load()) {
        std::fprintf(stderr, "emu: DEADLOCK -- no barrier progress for 90 s.  Waiting threads:\n");
        for (int t = 0; t < emu::kMaxThreads; ++t)
          if (emu::g_waiting[0][t].parity.load() >= 0)
            std::fprintf(stderr, "  thread %3d (warp %2d): barrier %08x parity %d\n", t, t >> 5,
                         emu::g_waiting[0][t].addr.load(), emu::g_waiting[0][t].parity.load());
        std::_Exit(3);
      }
    }
  }).detach();
  if (const char* e = std::getenv("GLOC_EMU_ASYNC")) emu::g_engine.start((unsigned)std::atoi(e));
  FILE* o = nullptr;
  if (mode == "conv") {
    const std::vector<int32_t> h = read_vec<int32_t>(f, 7);   // B H W Cin Cout last workers
    const int B = h[0], H = h[1], W = h[2], Cin = h[3], Cout = h[4], last = h[5], workers = h[6];
    const std::vector<uint16_t> act = read_vec<uint16_t>(f, (size_t)B * H * W * Cin),
                                wgt = read_vec<uint16_t>(f, (size_t)Cout * 9 * Cin);
    const std::vector<float> bias = read_vec<float>(f, (size_t)Cout);
    std::vector<uint16_t> out_h(last ? 0 : (size_t)B * H * W * Cout, 0xCDCD);
    std::vector<float> out_f(last ? (size_t)B * Cout * H * W : 0, NAN);
    // the library's maps: make_act_map / make_weight_map in encoder.cu
    const CUtensorMap map_in{act.data(), 4, {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B},
                             {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin},
                             {(uint32_t)kEncBK, (uint32_t)kEncTileW, (uint32_t)kEncTileH, 1}};
    const int bn = Cout >= 256 ? 256 : Cout;
    const CUtensorMap map_w = emu_map_2d(wgt.data(), (uint64_t)Cout, (uint64_t)9 * Cin, (uint32_t)bn);
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.bias = bias.data();
    a.out_nhwc = last ? nullptr : reinterpret_cast<__half*>(out_h.data());
    a.out_nchw = last ? out_f.data() : nullptr;
    if (bn == 64) run_conv<64>(map_in, map_w, a, workers);
    else if (bn == 128) run_conv<128>(map_in, map_w, a, workers);
    else run_conv<256>(map_in, map_w, a, workers);
    o = std::fopen(argv[3], "wb");
    if (last) std::fwrite(out_f.data(), 4, out_f.size(), o);
    else std::fwrite(out_h.data(), 2, out_h.size(), o);
  } else if (mode == "conv1") {
    const std::vector<int32_t> h = read_vec<int32_t>(f, 3);
    const int B = h[0], H = h[1], W = h[2];
    const std::vector<uint8_t> img = read_vec<uint8_t>(f, (size_t)B * H * W);
    const std::vector<float> w1 = read_vec<float>(f, 64 * 9), bias = read_vec<float>(f, 64);
    std::vector<uint16_t> out((size_t)B * H * W * 64, 0xCDCD);
    const size_t px = (size_t)B * H * W;
    launch_blocks((unsigned)((px + 255) / 256), 256, [&] {
      enc_conv1_kernel(img.data(), B, H, W, w1.data(), bias.data(), reinterpret_cast<__half*>(out.data()));
    });
    o = std::fopen(argv[3], "wb");
    std::fwrite(out.data(), 2, out.size(), o);
  } else if (mode == "pool") {
    const std::vector<int32_t> h = read_vec<int32_t>(f, 4);
    const int B = h[0], H = h[1], W = h[2], C = h[3];
    const std::vector<uint16_t> in = read_vec<uint16_t>(f, (size_t)B * H * W * C);
    std::vector<uint16_t> out((size_t)B * (H / 2) * (W / 2) * C, 0xCDCD);
    const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    launch_blocks((unsigned)((n + 255) / 256), 256, [&] {
      enc_maxpool2_kernel(reinterpret_cast<const __half*>(in.data()), B, H, W, C, reinterpret_cast<__half*>(out.data()));
    });
    o = std::fopen(argv[3], "wb");
    std::fwrite(out.data(), 2, out.size(), o);
  } else {
    return 2;
  }
  emu::g_done = true;
  emu::g_engine.finish();
  std::fclose(f);
  if (o) std::fclose(o);
  return 0;
}
