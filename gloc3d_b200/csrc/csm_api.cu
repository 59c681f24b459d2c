// csm_api.cu -- C ABI of stage 2 (scan-match verification).  See include/gloc3d.h
// for the reference interfaces each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "csm_store.cuh"

using namespace gloc;

// bev.cu
const uint8_t* gloc_bev_device_image(const gloc_bev_projector* b, gloc_bev_info* info);
int gloc_bev_device_of(const gloc_bev_projector* b);
cudaError_t gloc_bev_launch_level1(const uint8_t* img, size_t n, uint8_t* out, cudaStream_t s);
cudaError_t gloc_bev_launch_level1_aligned(const uint8_t* img, int W, int H, uint8_t* out, cudaStream_t s);

namespace {

// /root/reference/registration/3d/probability_values.h:64-67, float32 on purpose.
const float kMinProbability = 0.1f;
const float kMaxProbability = 1.f - kMinProbability;
const float kMinCorrespondenceCost = 1.f - kMaxProbability;
const float kMaxCorrespondenceCost = 1.f - kMinProbability;

// ValueToCorrespondenceCost table, 3d/probability_values.cpp:27-36,38-52,59-63
float value_to_cost(uint16_t value) {
  const uint16_t v = value & 32767u;  // the table repeats for update-marked values
  if (v == 0) return kMaxCorrespondenceCost;
  const float kScale = (kMaxCorrespondenceCost - kMinCorrespondenceCost) / 32766.f;
  return v * kScale + (kMinCorrespondenceCost - kScale);
}

size_t stack_bytes(int nx, int ny, int depth, long long* off) {
  size_t total = 0;
  for (int i = 0; i < depth; ++i) {
    const int w = 1 << i;
    if (off) off[i] = (long long)total;
    total += (size_t)(nx + w - 1) * (size_t)(ny + w - 1);
    total = (total + 15) & ~(size_t)15;
  }
  return total;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Build the uint16 -> uint8 table exactly as PrecomputationGrid2D does per cell:
// ComputeCellValue(1.f - |cost|), fast_..._2d.cpp:118-119,130-131,184-190.
void build_lut(std::vector<uint8_t>& lut) {
  lut.resize(65536);
  const float min_score = 1.f - kMaxCorrespondenceCost;
  const float max_score = 1.f - kMinCorrespondenceCost;
  for (int v = 0; v < 65536; ++v) {
    const float p = 1.f - std::fabs(value_to_cost((uint16_t)v));
    const long cell = std::lround((p - min_score) * (255.f / (max_score - min_score)));
    lut[v] = (uint8_t)std::min(255l, std::max(0l, cell));
  }
}

CsmBuf* all_bufs(gloc_csm_store* st, int i) {
  CsmBuf* b[] = {&st->d_recs, &st->pts, &st->pairs, &st->slots, &st->slot_gid, &st->hkeys, &st->hslot,
                 &st->ws, &st->rot, &st->bounds, &st->coarse, &st->top, &st->best, &st->survivors,
                 &st->nsurv, &st->nodes, &st->misc, &st->cells16, &st->stage_u8, &st->disc};
  return i < (int)(sizeof(b) / sizeof(b[0])) ? b[i] : nullptr;
}

// The width-1 grid sits in st->stage_u8 (device): pack it into the arena (bits when every
// cell is 0 or 255, else the bytes) and record it.  One host round trip when the encoding is
// not known in advance; none for grids that are binary by construction.
int add_from_stage(gloc_csm_store* st, int nx, int ny, double resolution, double max_x, double max_y,
                   bool known_binary, int* grid_id, const char* who) {
  const int stride = csm_bit_stride(nx);
  const size_t bit_bytes = (size_t)ny * stride * 4, n = (size_t)nx * ny;
  void* d_bits = nullptr;
  cudaError_t e = st->arena.alloc(bit_bytes, &d_bits);
  if (e == cudaSuccess) e = st->misc.reserve(64);
  if (e == cudaSuccess) e = cudaMemsetAsync(st->misc.p, 0, 4, st->stream);
  if (e == cudaSuccess)
    e = launch_csm_pack_bits((const uint8_t*)st->stage_u8.p, nx, ny, (unsigned*)d_bits, (int*)st->misc.p,
                             st->stream);
  st->stats.kernel_launches++;
  int not_binary = 0;
  if (e == cudaSuccess && !known_binary) {
    e = cudaMemcpyAsync(&not_binary, st->misc.p, 4, cudaMemcpyDeviceToHost, st->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  }
  CsmGridRec r{};
  r.nx = nx; r.ny = ny; r.resolution = resolution; r.max_x = max_x; r.max_y = max_y;
  r.data = d_bits;
  r.enc = 1;
  if (e == cudaSuccess && not_binary) {
    st->arena.rollback(d_bits, bit_bytes);
    void* d_raw = nullptr;
    e = st->arena.alloc(n, &d_raw);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_raw, st->stage_u8.p, n, cudaMemcpyDeviceToDevice, st->stream);
    r.data = d_raw;
    r.enc = 0;
  }
  // the staging buffer is reused by the next add: the kernels reading it must have finished
  if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);
  if (e != cudaSuccess) return fail(GLOC_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  if (r.enc == 0) st->n_graded++;
  st->max_nx = std::max(st->max_nx, nx);
  st->max_ny = std::max(st->max_ny, ny);
  st->recs.push_back(r);
  if (grid_id) *grid_id = (int)st->recs.size() - 1;
  return GLOC_OK;
}

int check_limits(int nx, int ny, double resolution, const char* who) {
  if (nx < 1 || ny < 1 || !(resolution > 0.))  // MapLimits ctor CHECKs, map_limits.h:44-46
    return fail(GLOC_ERR_INVALID, std::string(who) + ": bad limits");
  if ((long long)nx * ny > (1ll << 30)) return fail(GLOC_ERR_RANGE, std::string(who) + ": grid too large");
  return GLOC_OK;
}

}  // namespace

namespace gloc {

int csm_sync_recs(gloc_csm_store* st) {
  const size_t n = st->recs.size(), m = st->foreign.size();
  if (st->recs_on_device == n && !st->foreign_dirty) return GLOC_OK;
  const size_t need = (n + m) * sizeof(CsmGridRec);
  size_t first = st->recs_on_device;
  if (need > st->d_recs.bytes) {   // a growth drops the old contents: upload everything again
    GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
    GLOC_CUDA_TRY(st->d_recs.reserve(std::max(need * 2, (size_t)4096)));
    first = 0;
  }
  if (n > first)
    GLOC_CUDA_TRY(cudaMemcpyAsync((char*)st->d_recs.p + first * sizeof(CsmGridRec), st->recs.data() + first,
                                  (n - first) * sizeof(CsmGridRec), cudaMemcpyHostToDevice, st->stream));
  if (m)   // the peers' records follow the local ones
    GLOC_CUDA_TRY(cudaMemcpyAsync((char*)st->d_recs.p + n * sizeof(CsmGridRec), st->foreign.data(),
                                  m * sizeof(CsmGridRec), cudaMemcpyHostToDevice, st->stream));
  // the records live in std::vectors that may reallocate: the copies must not outlive this call
  GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
  st->recs_on_device = n;
  st->foreign_dirty = false;
  return GLOC_OK;
}

CsmParams csm_make_params(int n_lin, int n_ang, int depth, float min_score) {
  CsmParams prm;
  prm.n_lin = n_lin;
  prm.n_ang = n_ang;
  prm.S = 2 * n_ang + 1;
  prm.depth = depth;
  prm.step = 1 << (depth - 1);
  prm.max_side = (2 * n_lin) / prm.step + 1;
  prm.maxc = prm.max_side * prm.max_side;
  prm.W = (unsigned)(2 * n_lin + 1);
  prm.min_score = min_score;
  prm.min_s = 1.f - kMaxCorrespondenceCost;
  prm.coef = ((1.f - kMinCorrespondenceCost) - (1.f - kMaxCorrespondenceCost)) / 255.f;
  return prm;
}

void csm_host_rotations(int n_ang, double ang_step, std::vector<float2>* rot) {
  const long long S = 2ll * n_ang + 1;
  rot->resize((size_t)S);
  double delta_theta = -n_ang * ang_step;
  for (long long s = 0; s < S; ++s, delta_theta += ang_step) {
    const float ha = 0.5f * (float)delta_theta;
    (*rot)[(size_t)s] = make_float2(std::cos(ha), std::sin(ha));
  }
}

int csm_make_plan(int max_nx, int max_ny, bool all_binary, int n_lin, int depth, CsmBatchPlan* out) {
  CsmBatchPlan P;
  const int top = depth - 1, w = 1 << top;
  const int max_side = (2 * n_lin) / w + 1;
  const int wide_nx = max_nx + w - 1, wide_ny = max_ny + w - 1;
  P.dev.depth = depth;
  P.dev.n_lin = n_lin;
  P.dev.max_nx = max_nx;
  P.dev.max_ny = max_ny;
  // bit-sliced coarse scorer: binary grids whose coarsest level fits 64 columns per plane row
  // and shared memory, at most 16 lattice candidates per axis
  bool use_bits = all_binary && std::getenv("GLOC_CSM_NO_BITS") == nullptr && max_side <= 16 &&
                  ((wide_nx + n_lin - 1) >> top) < 64;
  P.dev.pmb_b0 = 0;
  P.dev.pmb_b1 = -1;
  if (use_bits && max_side <= 14 && std::getenv("GLOC_CSM_NO_PAIRED") == nullptr) {
    // Paired plane layout (CsmGridDev::pmb): the columns that can hold data, [c_lo, c_hi] (plane column
    // c = level bit w c + rx - px, px = n_lin), must fit two 32-column halves such that every window of
    // max_side columns ends inside the first or starts inside the second.
    const int c_lo = n_lin / w, c_hi = (wide_nx - 1 + n_lin) / w;
    const int b1 = std::max(c_lo, c_hi - 31);
    if (b1 - c_lo <= 33 - max_side) {
      P.dev.pmb_b0 = c_lo;
      P.dev.pmb_b1 = b1;
      P.bits_paired = true;
    }
  }
  if (use_bits) {
    const size_t smem = csm_coarse_bits_smem(top, csm_pmb_rows(wide_ny, n_lin, top, P.bits_paired));
    if (smem > (size_t)225 * 1024) {
      use_bits = false;
      P.bits_paired = false;
      P.dev.pmb_b1 = -1;
    } else {
      P.bits_smem = smem;
    }
  }
  P.use_bits = use_bits;
  if (use_bits) {
    P.dev.bits = 1;
    size_t off = 0;
    for (int l = 1; l < depth; ++l) {
      const int wl = 1 << l;
      P.dev.lvl_off[l] = off;
      off += align256((size_t)(max_ny + wl - 1) * (size_t)csm_bit_stride(max_nx + wl - 1) * 4);
    }
    P.dev.pmb_off = off;
    off += align256((size_t)w * w * (size_t)csm_pmb_rows(wide_ny, n_lin, top, P.bits_paired) * 8);
    P.dev.slot_bytes = std::max(off, (size_t)256);
    if (depth >= 2) {   // expand stage on the bit-packed level depth-2
      const int l2 = depth - 2, w2 = 1 << l2;
      const size_t smem = csm_expand_smem(max_ny + w2 - 1, csm_bit_stride(max_nx + w2 - 1));
      if (smem <= (size_t)112 * 1024) {
        P.exp_bits = true;
        P.exp_smem = smem;
      }
    }
  } else {
    P.dev.bits = 0;
    const size_t sb = align256(stack_bytes(max_nx, max_ny, depth, nullptr));
    const size_t pw = (size_t)(wide_nx + 2 * n_lin + w - 1) / w, ph = (size_t)(wide_ny + 2 * n_lin + w - 1) / w;
    const size_t cells = pw * ph * w * w;
    // absurd windows: bounds-checked kernel
    const bool use_pm = std::getenv("GLOC_CSM_NO_PM") == nullptr && cells <= ((size_t)1 << 27);
    P.dev.use_pm = use_pm ? 1 : 0;
    P.dev.pm_off = sb;
    P.dev.slot_bytes = sb + (use_pm ? align256(cells) : 0);
    P.pm_kernel = use_pm && (size_t)max_side * max_side * 4 + 20 * 4096 <= 150 * 1024;
  }
  *out = P;
  return GLOC_OK;
}

int csm_match_core(gloc_csm_store* st, const CsmBatchPlan& plan, const float* d_pts, CsmPairDev* d_pairs,
                   int n_pairs, const CsmParams& prm, const float2* d_rot, unsigned long long* h_best) {
  cudaStream_t stream = st->stream;
  int rc = csm_sync_recs(st);
  if (rc != GLOC_OK) return rc;
  const long long S = prm.S;
  const int depth = prm.depth;
  // sub-batches bound the coarse-score workspace and the working set, and keep candidate ids in 32 bits
  const long long per_pair = S * prm.maxc;
  long long sub = std::min<long long>(n_pairs, std::max<long long>(1, (1ll << 28) / per_pair));
  sub = std::min<long long>(sub, 65535);
  const size_t ws_budget = (size_t)(plan.use_bits ? 6 : 8) << 30;
  sub = std::min<long long>(sub, std::max<long long>(1, (long long)(ws_budget / plan.dev.slot_bytes)));
  if (const char* e = std::getenv("GLOC_CSM_SUB")) sub = std::max(1, std::min((int)sub, std::atoi(e)));
  unsigned long long init_key = 0ull;  // incumbent starts at (min_score, worst rank)
  if (prm.min_score > 0.f) {
    uint32_t u;
    std::memcpy(&u, &prm.min_score, 4);
    init_key = (unsigned long long)u << 32;
  }
  const int n_ctas = sm_count(st->device) * 8;   // persistent refinement CTAs (128 threads)
  int bits_warps = 12;   // rotations per CTA = 16 x warps (two lanes per rotation)
  if (const char* e = std::getenv("GLOC_CSM_BITS_WARPS")) bits_warps = std::min(12, std::max(1, std::atoi(e)));
  int hsize = 64;
  while (hsize < 2 * sub) hsize <<= 1;
  GLOC_CUDA_TRY(st->slots.reserve((size_t)sub * sizeof(CsmGridDev)));
  GLOC_CUDA_TRY(st->slot_gid.reserve((size_t)sub * sizeof(int)));
  GLOC_CUDA_TRY(st->hkeys.reserve((size_t)hsize * sizeof(int)));
  GLOC_CUDA_TRY(st->hslot.reserve((size_t)hsize * sizeof(int)));
  GLOC_CUDA_TRY(st->ws.reserve((size_t)sub * plan.dev.slot_bytes));
  GLOC_CUDA_TRY(st->bounds.reserve((size_t)sub * S * sizeof(CsmBounds)));
  GLOC_CUDA_TRY(st->coarse.reserve((size_t)sub * per_pair * sizeof(int)));
  GLOC_CUDA_TRY(st->survivors.reserve((size_t)sub * per_pair * sizeof(unsigned)));
  GLOC_CUDA_TRY(st->top.reserve((size_t)sub * 8));
  GLOC_CUDA_TRY(st->best.reserve((size_t)sub * 8));
  GLOC_CUDA_TRY(st->nsurv.reserve((size_t)sub * 4));
  // nodes handed from the expand stage to the depth-first refinement (24 B each)
  const unsigned node_cap = (unsigned)std::min<long long>(4ll * sub * per_pair, 16ll << 20);
  GLOC_CUDA_TRY(st->nodes.reserve((size_t)node_cap * sizeof(CsmNode)));
  GLOC_CUDA_TRY(st->misc.reserve(64));
  std::vector<unsigned long long> init((size_t)sub, init_key);
  const bool timing = std::getenv("GLOC_CSM_TIMING") != nullptr;
  for (long long p0 = 0; p0 < n_pairs; p0 += sub) {
    const int np = (int)std::min<long long>(sub, n_pairs - p0);
    CsmPairDev* dp = d_pairs + p0;
    GLOC_CUDA_TRY(cudaMemsetAsync(st->top.p, 0, (size_t)np * 8, stream));
    GLOC_CUDA_TRY(cudaMemsetAsync(st->nsurv.p, 0, (size_t)np * 4, stream));
    GLOC_CUDA_TRY(cudaMemsetAsync(st->misc.p, 0, 64, stream));
    GLOC_CUDA_TRY(cudaMemcpyAsync(st->best.p, init.data(), (size_t)np * 8, cudaMemcpyHostToDevice, stream));
    unsigned* n_surv = (unsigned*)st->nsurv.p;
    unsigned* n_nodes = (unsigned*)st->misc.p;
    unsigned* cursor = n_nodes + 1;
    int* n_slots = (int*)st->misc.p + 2;
    unsigned long long* counters = (unsigned long long*)((char*)st->misc.p + 16);
    const CsmGridDev* dg = (const CsmGridDev*)st->slots.p;
    // tuning aid: GLOC_CSM_TIMING=1 prints the duration of every stage of this sub-batch
    cudaEvent_t tev[7];
    if (timing) {
      for (auto& e : tev) cudaEventCreate(&e);
      cudaEventRecord(tev[6], stream);
    }
    // ---- working set: distinct grids -> slots, derived structures rebuilt on the device
    GLOC_CUDA_TRY(launch_csm_assign_slots(dp, np, (int*)st->hkeys.p, (int*)st->hslot.p, hsize,
                                          (int*)st->slot_gid.p, n_slots, stream));
    GLOC_CUDA_TRY(launch_csm_prepare_slots((const CsmGridRec*)st->d_recs.p, (const int*)st->slot_gid.p,
                                           n_slots, np, plan.dev, (unsigned char*)st->ws.p,
                                           (CsmGridDev*)st->slots.p, stream));
    st->stats.kernel_launches += 3;
    GLOC_CUDA_TRY(launch_csm_build_slots((const CsmGridRec*)st->d_recs.p, (const int*)st->slot_gid.p, dg,
                                         n_slots, np, plan.dev, stream, &st->stats.kernel_launches));
    if (timing) cudaEventRecord(tev[0], stream);
    st->prof.begin(stream);
    cudaError_t ce;
    if (plan.use_bits)
      ce = launch_csm_coarse_bits(dg, dp, np, d_pts, d_rot, prm, (CsmBounds*)st->bounds.p,
                                  (int*)st->coarse.p, (unsigned long long*)st->top.p, plan.bits_smem,
                                  bits_warps, plan.bits_paired, stream);
    else
      ce = launch_csm_coarse(dg, dp, np, d_pts, d_rot, prm, (CsmBounds*)st->bounds.p, (int*)st->coarse.p,
                             (unsigned long long*)st->top.p, stream, plan.pm_kernel);
    st->prof.end(stream);
    GLOC_CUDA_TRY(ce);
    if (timing) cudaEventRecord(tev[1], stream);
    GLOC_CUDA_TRY(launch_csm_seed(dg, dp, np, d_pts, d_rot, prm, (const CsmBounds*)st->bounds.p,
                                  (const unsigned long long*)st->top.p, (unsigned long long*)st->best.p,
                                  stream));
    if (timing) cudaEventRecord(tev[2], stream);
    GLOC_CUDA_TRY(launch_csm_filter(dp, np, prm, (const CsmBounds*)st->bounds.p, (const int*)st->coarse.p,
                                    (const unsigned long long*)st->best.p, (unsigned*)st->survivors.p,
                                    n_surv, stream));
    if (timing) cudaEventRecord(tev[3], stream);
    if (depth >= 2) {
      GLOC_CUDA_TRY(launch_csm_expand(dg, dp, np, d_pts, d_rot, prm, (const CsmBounds*)st->bounds.p,
                                      (const int*)st->coarse.p, (const unsigned*)st->survivors.p, n_surv,
                                      (unsigned long long*)st->best.p, (CsmNode*)st->nodes.p, n_nodes,
                                      node_cap, counters, 128, plan.exp_bits && plan.use_bits,
                                      plan.exp_smem, stream));
      if (timing) cudaEventRecord(tev[5], stream);
      GLOC_CUDA_TRY(launch_csm_refine(dg, dp, d_pts, d_rot, prm, (const CsmBounds*)st->bounds.p,
                                      (const CsmNode*)st->nodes.p, n_nodes, node_cap, cursor,
                                      (unsigned long long*)st->best.p, counters, n_ctas, stream));
    }
    if (timing) cudaEventRecord(tev[4], stream);
    GLOC_CUDA_TRY(cudaMemcpyAsync(h_best + p0, st->best.p, (size_t)np * 8, cudaMemcpyDeviceToHost, stream));
    unsigned long long hc = 0;
    unsigned hn = 0;
    GLOC_CUDA_TRY(cudaMemcpyAsync(&hc, counters, 8, cudaMemcpyDeviceToHost, stream));
    GLOC_CUDA_TRY(cudaMemcpyAsync(&hn, n_nodes, 4, cudaMemcpyDeviceToHost, stream));
    GLOC_CUDA_TRY(cudaStreamSynchronize(stream));
    if (hn > node_cap)
      return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: branch-and-bound node list overflowed "
                                  "(more than 16M open nodes in one sub-batch); match fewer pairs per call");
    if (timing) {
      float t[4] = {0, 0, 0, 0}, t_build = 0.f;
      std::vector<unsigned> hs((size_t)np);
      cudaMemcpy(hs.data(), n_surv, (size_t)np * 4, cudaMemcpyDeviceToHost);
      int hslots = 0;
      cudaMemcpy(&hslots, n_slots, 4, cudaMemcpyDeviceToHost);
      unsigned long long ns = 0;
      for (unsigned v : hs) ns += v;
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
      cudaEventElapsedTime(&t_build, tev[6], tev[0]);
      float t_exp = 0.f;
      if (depth >= 2) cudaEventElapsedTime(&t_exp, tev[3], tev[5]);
      fprintf(stderr, "[csm] pairs=%d grids=%d build=%.3f ms coarse(%s)=%.3f ms seed=%.3f filter=%.3f "
                      "expand+refine=%.3f (expand %.3f) | survivors=%llu nodes=%u expanded=%llu\n",
              np, hslots, t_build, plan.use_bits ? (plan.bits_paired ? "bits, paired rows" : "bits") : "u8", t[0], t[1], t[2], t[3], t_exp, ns, hn, hc);
      unsigned mx = 0;
      for (unsigned v : hs) mx = std::max(mx, v);
      // where the survivors come from: pairs that end up matching vs the rest
      unsigned long long ns_found = 0, n_found = 0;
      unsigned hist[6] = {0, 0, 0, 0, 0, 0};   // survivors per pair: 0, <64, <256, <1024, <4096, more
      for (int i = 0; i < np; ++i) {
        const unsigned long long key = h_best[p0 + i];
        uint32_t sb = (uint32_t)(key >> 32);
        float sc;
        std::memcpy(&sc, &sb, 4);
        if (key != 0 && sc > prm.min_score) {
          ns_found += hs[(size_t)i];
          ++n_found;
        }
        const unsigned v = hs[(size_t)i];
        hist[v == 0 ? 0 : v < 64 ? 1 : v < 256 ? 2 : v < 1024 ? 3 : v < 4096 ? 4 : 5]++;
      }
      fprintf(stderr, "[csm] max survivors in one pair = %u; %llu pairs matched and hold %llu of the survivors; "
                      "pairs by survivors 0/<64/<256/<1024/<4096/more: %u %u %u %u %u %u\n", mx, n_found, ns_found,
              hist[0], hist[1], hist[2], hist[3], hist[4], hist[5]);
      for (auto& e : tev) cudaEventDestroy(e);
    }
    st->stats.kernel_launches += depth >= 2 ? 5 : 3;
    st->stats.refined_nodes += hc;
    st->stats.coarse_candidates += (uint64_t)np * (uint64_t)per_pair;  // upper bound (slots)
  }
  st->stats.matches += (uint64_t)n_pairs;
  return GLOC_OK;
}

void csm_decode(unsigned long long key, const CsmParams& prm, double ang_step, double resolution,
                const double* init, float min_score, gloc_csm_result* rp) {
  gloc_csm_result& r = *rp;
  std::memset(&r, 0, sizeof(r));
  uint32_t sb = (uint32_t)(key >> 32);
  float score;
  std::memcpy(&score, &sb, 4);
  r.score = min_score;
  if (key != 0 && score > min_score) {
    const uint32_t rank = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    const int s = (int)(rank / (prm.W * prm.W));
    const int xo = (int)((rank / prm.W) % prm.W) - prm.n_lin;
    const int yo = (int)(rank % prm.W) - prm.n_lin;
    const double cx = -yo * resolution;
    const double cy = -xo * resolution;
    const double orientation = (s - prm.n_ang) * ang_step;
    r.found = 1;
    r.score = score;
    r.scan_index = s;
    r.x_offset = xo;
    r.y_offset = yo;
    r.pose_x = init[0] + cx;
    r.pose_y = init[1] + cy;
    r.pose_yaw = init[2] + orientation;
  }
}

}  // namespace gloc

extern "C" {

int gloc_csm_create(gloc_csm_store** out, int device) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_csm_create: out is null");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_csm_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(GLOC_ERR_INVALID, "gloc_csm_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return fail(GLOC_ERR_CUDA, "gloc_csm_create: device is not sm_100 (kernels are sm_100a only)");
  DeviceGuard g(device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_csm_create: cudaSetDevice failed");
  gloc_csm_store* st = new (std::nothrow) gloc_csm_store;
  if (!st) return fail(GLOC_ERR_NOMEM, "gloc_csm_create: out of host memory");
  st->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_lut, 65536);
  if (e == cudaSuccess) {
    std::vector<uint8_t> lut;
    build_lut(lut);
    e = cudaMemcpy(st->d_lut, lut.data(), 65536, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    gloc_csm_destroy(st);
    return fail(GLOC_ERR_CUDA, std::string("gloc_csm_create: ") + cudaGetErrorString(e));
  }
  *out = st;
  return GLOC_OK;
}

void gloc_csm_destroy(gloc_csm_store* st) {
  if (!st) return;
  DeviceGuard g(st->device);
  if (st->stream) {
    cudaStreamSynchronize(st->stream);
    cudaStreamDestroy(st->stream);
  }
  st->arena.release();
  if (st->d_lut) cudaFree(st->d_lut);
  for (int i = 0; CsmBuf* b = all_bufs(st, i); ++i) b->release();
  delete st;
}

int gloc_csm_add_grid_cells(gloc_csm_store* st, const uint16_t* cells, int nx, int ny,
                            double resolution, double max_x, double max_y, int* grid_id) {
  if (!st || !cells) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_cells: null argument");
  int rc = check_limits(nx, ny, resolution, "gloc_csm_add_grid_cells");
  if (rc != GLOC_OK) return rc;
  DeviceGuard g(st->device);
  const size_t n = (size_t)nx * ny;
  GLOC_CUDA_TRY(st->cells16.reserve(n * 2));
  GLOC_CUDA_TRY(st->stage_u8.reserve(n));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->cells16.p, cells, n * 2, cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(launch_csm_level1_from_cells((const uint16_t*)st->cells16.p, st->d_lut, n,
                                             (uint8_t*)st->stage_u8.p, st->stream));
  st->stats.kernel_launches++;
  return add_from_stage(st, nx, ny, resolution, max_x, max_y, false, grid_id, "gloc_csm_add_grid_cells");
}

int gloc_csm_add_grid_u8(gloc_csm_store* st, const uint8_t* level1, int nx, int ny,
                         double resolution, double max_x, double max_y, int* grid_id) {
  if (!st || !level1) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_u8: null argument");
  int rc = check_limits(nx, ny, resolution, "gloc_csm_add_grid_u8");
  if (rc != GLOC_OK) return rc;
  DeviceGuard g(st->device);
  const size_t n = (size_t)nx * ny;
  GLOC_CUDA_TRY(st->stage_u8.reserve(n));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->stage_u8.p, level1, n, cudaMemcpyHostToDevice, st->stream));
  return add_from_stage(st, nx, ny, resolution, max_x, max_y, false, grid_id, "gloc_csm_add_grid_u8");
}

int gloc_csm_add_grid_from_bev(gloc_csm_store* st, gloc_bev_projector* bev, int* grid_id) {
  if (!st || !bev) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev: null argument");
  gloc_bev_info I;
  const uint8_t* d_img = gloc_bev_device_image(bev, &I);
  if (!d_img || I.width < 1 || I.height < 1)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_csm_add_grid_from_bev: the projector holds no image");
  if (gloc_bev_device_of(bev) != st->device)
    return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev: projector and store live on different devices");
  DeviceGuard g(st->device);
  // ProjectToGrid, 3d/submap_3d.cpp:376-388: limits from the voxel bounding box
  const double max_x = (I.min_ix + I.width - 1) * I.resolution;
  const double max_y = (I.min_iy + I.height - 1) * I.resolution;
  int rc = check_limits(I.width, I.height, I.resolution, "gloc_csm_add_grid_from_bev");
  if (rc != GLOC_OK) return rc;
  const size_t n = (size_t)I.width * I.height;
  GLOC_CUDA_TRY(st->stage_u8.reserve(n));
  GLOC_CUDA_TRY(gloc_bev_launch_level1(d_img, n, (uint8_t*)st->stage_u8.p, st->stream));
  st->stats.kernel_launches++;
  return add_from_stage(st, I.width, I.height, I.resolution, max_x, max_y, true, grid_id,
                        "gloc_csm_add_grid_from_bev");
}

int gloc_csm_add_grid_from_bev_aligned(gloc_csm_store* st, gloc_bev_projector* bev, int* grid_id) {
  if (!st || !bev) return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev_aligned: null argument");
  gloc_bev_info I;
  const uint8_t* d_img = gloc_bev_device_image(bev, &I);
  if (!d_img || I.width < 1 || I.height < 1)
    return fail(GLOC_ERR_NOT_BUILT, "gloc_csm_add_grid_from_bev_aligned: the projector holds no image");
  if (gloc_bev_device_of(bev) != st->device)
    return fail(GLOC_ERR_INVALID, "gloc_csm_add_grid_from_bev_aligned: projector and store live on different devices");
  DeviceGuard g(st->device);
  // cell x runs along world -y and cell y along world -x (2d/map_limits.h:69-76): the grid has
  // height x width cells and max() half a cell beyond the last voxel centre
  const int nx = I.height, ny = I.width;
  const double max_x = (I.min_ix + I.width - 1 + 0.5) * I.resolution;
  const double max_y = (I.min_iy + I.height - 1 + 0.5) * I.resolution;
  int rc = check_limits(nx, ny, I.resolution, "gloc_csm_add_grid_from_bev_aligned");
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(st->stage_u8.reserve((size_t)nx * ny));
  GLOC_CUDA_TRY(gloc_bev_launch_level1_aligned(d_img, I.width, I.height, (uint8_t*)st->stage_u8.p, st->stream));
  st->stats.kernel_launches++;
  return add_from_stage(st, nx, ny, I.resolution, max_x, max_y, true, grid_id,
                        "gloc_csm_add_grid_from_bev_aligned");
}

int gloc_csm_num_grids(const gloc_csm_store* st) { return st ? (int)st->recs.size() : 0; }

int gloc_csm_get_grid_info(const gloc_csm_store* st, int grid_id, gloc_grid_info* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_grid_info: null argument");
  if (grid_id < 0 || grid_id >= (int)st->recs.size())
    return fail(GLOC_ERR_INVALID, "gloc_csm_get_grid_info: bad grid id");
  const CsmGridRec& r = st->recs[grid_id];
  out->nx = r.nx;
  out->ny = r.ny;
  out->resolution = r.resolution;
  out->max_x = r.max_x;
  out->max_y = r.max_y;
  return GLOC_OK;
}

int gloc_csm_store_bytes(const gloc_csm_store* st, uint64_t* grid_bytes, uint64_t* workspace_bytes) {
  if (!st) return fail(GLOC_ERR_INVALID, "gloc_csm_store_bytes: null store");
  if (grid_bytes) *grid_bytes = st->arena.total;
  if (workspace_bytes) {
    uint64_t t = 0;
    for (int i = 0; const CsmBuf* b = all_bufs(const_cast<gloc_csm_store*>(st), i); ++i) t += b->bytes;
    *workspace_bytes = t;
  }
  return GLOC_OK;
}

int gloc_csm_get_precomputation_grid(gloc_csm_store* st, int grid_id, int width, uint8_t* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_precomputation_grid: null argument");
  if (grid_id < 0 || grid_id >= (int)st->recs.size())
    return fail(GLOC_ERR_INVALID, "gloc_csm_get_precomputation_grid: bad grid id");
  int level = 0;
  while ((1 << level) < width) ++level;
  if (width < 1 || (1 << level) != width || level >= kCsmMaxDepth)  // CHECK_GE(width, 1)
    return fail(GLOC_ERR_RANGE, "gloc_csm_get_precomputation_grid: width must be 1,2,4,...,128");
  DeviceGuard g(st->device);
  const CsmGridRec& r = st->recs[grid_id];
  int rc = csm_sync_recs(st);
  if (rc != GLOC_OK) return rc;
  // one slot of the uint8 working set, built exactly as a batch builds it
  CsmPlan plan{};
  plan.bits = 0;
  plan.depth = level + 1;
  plan.n_lin = 0;
  plan.use_pm = 0;
  plan.max_nx = r.nx;
  plan.max_ny = r.ny;
  long long off[kCsmMaxDepth] = {0};
  plan.slot_bytes = align256(stack_bytes(r.nx, r.ny, level + 1, off));
  GLOC_CUDA_TRY(st->ws.reserve(plan.slot_bytes));
  GLOC_CUDA_TRY(st->slots.reserve(sizeof(CsmGridDev)));
  GLOC_CUDA_TRY(st->slot_gid.reserve(sizeof(int)));
  GLOC_CUDA_TRY(st->misc.reserve(64));
  const int one = 1;
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->slot_gid.p, &grid_id, 4, cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->misc.p, &one, 4, cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(launch_csm_prepare_slots((const CsmGridRec*)st->d_recs.p, (const int*)st->slot_gid.p,
                                         (const int*)st->misc.p, 1, plan, (unsigned char*)st->ws.p,
                                         (CsmGridDev*)st->slots.p, st->stream));
  st->stats.kernel_launches++;
  GLOC_CUDA_TRY(launch_csm_build_slots((const CsmGridRec*)st->d_recs.p, (const int*)st->slot_gid.p,
                                       (const CsmGridDev*)st->slots.p, (const int*)st->misc.p, 1, plan,
                                       st->stream, &st->stats.kernel_launches));
  const size_t n = (size_t)(r.nx + width - 1) * (size_t)(r.ny + width - 1);
  GLOC_CUDA_TRY(cudaMemcpyAsync(out, (const uint8_t*)st->ws.p + off[level], n, cudaMemcpyDeviceToHost,
                                st->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
  return GLOC_OK;
}

int gloc_csm_match_batch(gloc_csm_store* st, const float* pts, const int64_t* scan_offsets,
                         int n_scans, const int* grid_ids, const int* scan_ids,
                         const double* init_xyyaw, int n_pairs, int n_lin, int n_ang,
                         double ang_step, int depth, float min_score, gloc_csm_result* results) {
  if (!st) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: null store");
  if (n_pairs == 0) return GLOC_OK;
  if (!pts || !scan_offsets || !grid_ids || !scan_ids || !init_xyyaw || !results || n_pairs < 0 ||
      n_scans < 1)
    return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: null argument");
  if (depth < 1 || depth > kCsmMaxDepth)  // CHECK_GE(branch_and_bound_depth, 1), fast_..._2d.cpp:195
    return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: depth must be in [1, 8]");
  if (n_lin < 0 || n_ang < 0) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: negative window");
  const long long S = 2ll * n_ang + 1, W = 2ll * n_lin + 1;
  if (S * W * W >= (1ll << 32) || S > 65535)
    return fail(GLOC_ERR_RANGE, "gloc_csm_match_batch: search window too large (scans*(2*n_lin+1)^2 must be < 2^32)");
  DeviceGuard guard(st->device);
  if (!guard.ok) return fail(GLOC_ERR_CUDA, "gloc_csm_match_batch: cudaSetDevice failed");
  const CsmParams prm = csm_make_params(n_lin, n_ang, depth, min_score);
  std::vector<float2> rot;
  csm_host_rotations(n_ang, ang_step, &rot);
  const int64_t total_pts = scan_offsets[n_scans];
  if (total_pts < 0) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad scan offsets");

  // validate; the plan depends on the dimensions of the batch's grids only
  std::vector<CsmPairDev> hp((size_t)n_pairs);
  int max_nx = 0, max_ny = 0;
  bool all_binary = true;
  for (int i = 0; i < n_pairs; ++i) {
    const int gi = grid_ids[i], si = scan_ids[i];
    if (gi < 0 || gi >= (int)st->recs.size() || si < 0 || si >= n_scans)
      return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad grid or scan id");
    const int64_t b = scan_offsets[si], e = scan_offsets[si + 1];
    if (b < 0 || e < b || e > total_pts || e - b > INT32_MAX)
      return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: bad scan offsets");
    if (e == b) return fail(GLOC_ERR_INVALID, "gloc_csm_match_batch: empty scan");
    const CsmGridRec& r = st->recs[gi];
    max_nx = std::max(max_nx, r.nx);
    max_ny = std::max(max_ny, r.ny);
    all_binary = all_binary && r.enc == 1;
    CsmPairDev& p = hp[(size_t)i];
    p.grid = 0;
    p.gid = gi;
    p.pt_begin = b;
    p.n_pts = (int)(e - b);
    // initial_rotation.cast<float>().angle() -> Quaternionf(AngleAxisf) (fast_..._2d.cpp:278-283)
    const float ha = 0.5f * (float)init_xyyaw[3 * i + 2];
    p.w0 = std::cos(ha);
    p.z0 = std::sin(ha);
    p.tx = (float)init_xyyaw[3 * i];      // Eigen::Translation2f(double, double), :287-288
    p.ty = (float)init_xyyaw[3 * i + 1];
  }
  CsmBatchPlan plan;
  csm_make_plan(max_nx, max_ny, all_binary, n_lin, depth, &plan);

  cudaStream_t stream = st->stream;
  GLOC_CUDA_TRY(st->pts.reserve((size_t)total_pts * 3 * sizeof(float)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(st->pairs.reserve((size_t)n_pairs * sizeof(CsmPairDev)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->pts.p, pts, (size_t)total_pts * 3 * sizeof(float),
                                cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2),
                                cudaMemcpyHostToDevice, stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->pairs.p, hp.data(), (size_t)n_pairs * sizeof(CsmPairDev),
                                cudaMemcpyHostToDevice, stream));
  std::vector<unsigned long long> hbest((size_t)n_pairs);
  int rc = csm_match_core(st, plan, (const float*)st->pts.p, (CsmPairDev*)st->pairs.p, n_pairs, prm,
                          (const float2*)st->rot.p, hbest.data());
  if (rc != GLOC_OK) return rc;
  for (int i = 0; i < n_pairs; ++i)
    csm_decode(hbest[(size_t)i], prm, ang_step, st->recs[grid_ids[i]].resolution, init_xyyaw + 3 * i,
               min_score, &results[i]);
  return GLOC_OK;
}

int gloc_csm_discretize(gloc_csm_store* st, const float* pts, int n_pts, double init_x,
                        double init_y, double init_yaw, int n_ang, double ang_step,
                        double resolution, double max_x, double max_y, int32_t* out_cells) {
  if (!st || !pts || !out_cells || n_pts < 0 || n_ang < 0)
    return fail(GLOC_ERR_INVALID, "gloc_csm_discretize: bad argument");
  if (n_pts == 0) return GLOC_OK;
  DeviceGuard guard(st->device);
  const int S = 2 * n_ang + 1;
  std::vector<float2> rot((size_t)S);
  double delta_theta = -n_ang * ang_step;
  for (int s = 0; s < S; ++s, delta_theta += ang_step) {
    const float ha = 0.5f * (float)delta_theta;
    rot[(size_t)s] = make_float2(std::cos(ha), std::sin(ha));
  }
  const float ha0 = 0.5f * (float)init_yaw;
  GLOC_CUDA_TRY(st->pts.reserve((size_t)n_pts * 3 * sizeof(float)));
  GLOC_CUDA_TRY(st->rot.reserve((size_t)S * sizeof(float2)));
  GLOC_CUDA_TRY(st->disc.reserve((size_t)S * n_pts * 2 * sizeof(int)));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->pts.p, pts, (size_t)n_pts * 3 * sizeof(float),
                                cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(st->rot.p, rot.data(), (size_t)S * sizeof(float2),
                                cudaMemcpyHostToDevice, st->stream));
  GLOC_CUDA_TRY(launch_csm_discretize((const float*)st->pts.p, n_pts, std::cos(ha0), std::sin(ha0),
                                      (float)init_x, (float)init_y, (const float2*)st->rot.p, S,
                                      resolution, max_x, max_y, (int*)st->disc.p, st->stream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(out_cells, st->disc.p, (size_t)S * n_pts * 2 * sizeof(int),
                                cudaMemcpyDeviceToHost, st->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(st->stream));
  st->stats.kernel_launches++;
  return GLOC_OK;
}

// SearchParameters production ctor, correlative_scan_matcher_2d.cpp:27-55 (host-side, O(P))
int gloc_csm_search_params(double linear_window, double angular_window, const float* pts,
                           int n_pts, double resolution, int* n_lin, int* n_ang,
                           double* ang_step) {
  if ((n_pts > 0 && !pts) || !n_lin || !n_ang || !ang_step || !(resolution > 0.))
    return fail(GLOC_ERR_INVALID, "gloc_csm_search_params: bad argument");
  float max_scan_range = 3.f * resolution;
  for (int i = 0; i < n_pts; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1];
    const float range = std::sqrt(x * x + y * y);
    max_scan_range = std::max(range, max_scan_range);
  }
  const double kSafetyMargin = 1. - 1e-3;
  const float r2 = max_scan_range * max_scan_range;
  *ang_step = kSafetyMargin * std::acos(1. - (resolution * resolution) / (2. * r2));
  *n_ang = (int)std::ceil(angular_window / *ang_step);
  *n_lin = (int)std::ceil(linear_window / resolution);
  return GLOC_OK;
}

// GridToVirtualPointCloud, fast_..._2d.cpp:78-95 (host-side, O(cells))
int gloc_csm_grid_to_points(const uint16_t* cells, int nx, int ny, double resolution, double ox,
                            double oy, float* pts, int capacity, int* n_out) {
  if (!cells || nx < 1 || ny < 1 || !n_out)
    return fail(GLOC_ERR_INVALID, "gloc_csm_grid_to_points: bad argument");
  int n = 0;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const float cost = value_to_cost(cells[(size_t)nx * j + i]);
      if (cost < 0.11) {
        if (pts && n < capacity) {
          pts[3 * n] = (float)(ox + i * resolution);
          pts[3 * n + 1] = (float)(oy + j * resolution);
          pts[3 * n + 2] = 0.f;
        }
        ++n;
      }
    }
  *n_out = n;
  return GLOC_OK;
}

int gloc_csm_set_profiling(gloc_csm_store* st, int enabled) {
  if (!st) return fail(GLOC_ERR_INVALID, "gloc_csm_set_profiling: null store");
  st->prof.enabled = enabled != 0;
  return GLOC_OK;
}

int gloc_csm_get_profile(gloc_csm_store* st, gloc_profile* out) {
  if (!st || !out) return fail(GLOC_ERR_INVALID, "gloc_csm_get_profile: null argument");
  DeviceGuard g(st->device);
  st->prof.collect(&out->dominant_ms, &out->dominant_launches);
  return GLOC_OK;
}

int gloc_csm_get_stats(const gloc_csm_store* st, gloc_csm_stats* stats) {
  if (!st || !stats) return fail(GLOC_ERR_INVALID, "gloc_csm_get_stats: null argument");
  *stats = st->stats;
  return GLOC_OK;
}

}  // extern "C"
