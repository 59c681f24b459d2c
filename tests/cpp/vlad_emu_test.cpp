// Runs the kernels of gloc3d_b200/csrc/vlad.cu on the host through tests/cpp/cuda_emu.hpp with
// the launch geometry of forward_device() and writes the descriptors: tests/test_vlad_emulated.py
// compares them with the oracle.  The kernel text is extracted from vlad.cu at build time
// (_vlad_kernels.inc, `extern __shared__` rewritten to `extern`).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_emu.hpp"

namespace gloc {
namespace {
float w_s[1 << 16];   // the dynamic shared memory of vlad_assign_kernel
#include "_vlad_kernels.inc"
}  // namespace
}  // namespace gloc

using namespace gloc;

static std::vector<float> read_floats(FILE* f, size_t n) {
  std::vector<float> v(n);
  if (n && std::fread(v.data(), 4, n, f) != n) {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[6];
  if (std::fread(hdr, 4, 6, f) != 6) return 2;
  const int B = hdr[0], C = hdr[1], S = hdr[2], K = hdr[3], D = hdr[4], has_bias = hdr[5];
  const int I = K * C;
  std::vector<float> x = read_floats(f, (size_t)B * C * S), conv_w = read_floats(f, (size_t)K * C),
                     conv_b = read_floats(f, (size_t)K), cent = read_floats(f, (size_t)K * C),
                     hidden = read_floats(f, (size_t)I * D);
  std::fclose(f);
  std::vector<float> a((size_t)B * K * S), inv((size_t)B * S), V((size_t)B * I), out((size_t)B * D);
  const int n_chunks = (I + kFcRows - 1) / kFcRows;
  std::vector<float> partial((size_t)n_chunks * B * D);

  // the launches of forward_device(), gloc3d_b200/csrc/vlad.cu
  emu::launch(dim3((S + kAssignThreads - 1) / kAssignThreads, B), dim3(kAssignThreads), vlad_assign_kernel,
              x.data(), conv_w.data(), has_bias ? conv_b.data() : nullptr, C, S, K, a.data(), inv.data());
  emu::launch(dim3((C + 31) / 32, B), dim3(256), vlad_aggregate_kernel, x.data(), a.data(), inv.data(),
              cent.data(), C, S, K, V.data());
  emu::launch(dim3(B), dim3(256), vlad_normalize_kernel, V.data(), C, K);
  for (int b0 = 0; b0 < B; b0 += kFcBatch) {
    const int nb = std::min(kFcBatch, B - b0);
    emu::launch(dim3((D + kFcCols - 1) / kFcCols, n_chunks), dim3(kFcCols), vlad_fc_kernel, V.data(),
                hidden.data(), I, D, b0, nb, partial.data(), B);
  }
  const size_t n_out = (size_t)B * D;
  emu::launch(dim3((unsigned)((n_out + 255) / 256)), dim3(256), vlad_fc_reduce_kernel, partial.data(), n_chunks, B,
              D, out.data());

  FILE* g = std::fopen(argv[2], "wb");
  if (!g) return 2;
  std::fwrite(out.data(), 4, out.size(), g);
  std::fwrite(V.data(), 4, V.size(), g);
  std::fclose(g);
  return 0;
}
