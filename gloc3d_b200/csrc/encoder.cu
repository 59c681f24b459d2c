// encoder.cu -- the convolutional encoder in front of the NetVLAD_fc head:
// VGG16 `features[:-2]` as the reference assembles it (/root/reference/main.py:531-536: the 13
// 3x3 convolutions with ReLU and the first four 2x2 max-pools; the last ReLU and pool dropped),
// run on the 768 x 768 BEV occupancy image of RpyPCLoopDetector::get_place_feature
// (loop_detector.cpp:137-172: CV_8UC3 with three identical channels, scaled by 1/255).
// Batched, device to device: images from gloc_bev_get_cnn_input's plane in, the [B][512][S]
// feature maps gloc_vlad_forward_device takes out.
//
// * conv1_1 sees three identical channels: it is folded on the host into a one-channel 3x3
//   convolution on the uint8 image (weights summed over the input channels, 1/255 included) and
//   runs as a small SIMT kernel writing FP16 NHWC.
// * every other convolution is an implicit GEMM on the tensor cores: M = 128 pixels (an 8 x 16
//   patch of one image), N = 64 / 128 / 256 output channels, K = 9 taps x Cin in k-blocks of 64
//   channels.  The activation tile of a tap is ONE 4-D TMA box {64 ch, 16 x, 8 y, 1 image} of
//   the FP16 NHWC tensor at (x0 + dx - 1, y0 + dy - 1): the zero padding is the TMA's
//   out-of-bounds fill, the box lands in shared memory as 128 rows of 128 bytes (pixel-major,
//   128B swizzle) -- exactly the K-major operand tcgen05.mma wants; weights [Cout][9 Cin] come
//   through a 2-D map.  tcgen05.mma kind::f16 (FP16 operands, FP32 accumulation in tensor
//   memory, two accumulator stages), warp-specialised like the shortlist GEMM: warp 0 TMA,
//   warp 1 MMA, warp 2 TMEM allocation, warps 4-7 epilogue (bias, ReLU, FP16 NHWC store; the
//   last layer stores FP32 NCHW without ReLU).
// * 2x2 max-pools: one pass over FP16 NHWC (post-ReLU values: the order of non-negative halves
//   is the order of their bit patterns, so the pool is a signed 16-bit max).
// Precision: FP16 operands / FP32 accumulation against the reference's FP32 weights run through
// cuDNN with TF32 allowed (PyTorch's default for convolutions): the same 10-bit significand.
//
// STATUS: written after round 1 ran out of GPU minutes.  Compiles for sm_100a; the kernel
// source runs on the host against the functional tcgen05/TMA model of tests/cpp
// (tests/test_encoder_emulated.py) and matches a float64 convolution there.  Not yet run on a
// GPU: tests/test_encoder_gpu.py is opt-in (GLOC_TEST_UNVERIFIED=1).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gloc3d.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gloc {

namespace {

// [enc-kernels-begin] (tests/cpp/encoder_emu_test.cpp compiles the text up to [enc-kernels-end] for the host)
constexpr int kEncBM = 128;                      // pixels per tile = TMEM lanes = UMMA M
constexpr int kEncTileH = 8, kEncTileW = 16;     // the tile is an 8 x 16 patch of one image
constexpr int kEncBK = 64;                       // channels per k-block (128 B of FP16: one swizzle row)
constexpr int kEncUK = 16;                       // UMMA K for 16-bit inputs
constexpr int kEncThreads = 256;                 // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-7 epilogue
constexpr int kEncABytes = kEncBM * kEncBK * 2;  // 16 KB
constexpr int kEncSmemBudget = 200 * 1024;

template <int BN>
struct EncCfg {
  static constexpr int kBBytes = BN * kEncBK * 2;
  static constexpr int kStageBytes = kEncABytes + kBBytes;
  static constexpr int kStages = kEncSmemBudget / kStageBytes > 6 ? 6 : kEncSmemBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;       // two accumulator stages (a power of two >= 32)
  static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;
};

struct ConvArgs {
  int B, H, W, Cin, Cout;      // H % 8 == 0, W % 16 == 0, Cin % 64 == 0, Cout % BN == 0
  const float* bias;           // [Cout]
  __half* out_nhwc;            // [B][H][W][Cout] FP16 with ReLU, or null
  float* out_nchw;             // [B][Cout][H*W] FP32 without ReLU (the last layer), or null
};

template <int BN>
__global__ void __launch_bounds__(kEncThreads, 1)
enc_conv3x3_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                   ConvArgs a) {
  using Cfg = EncCfg<BN>;
  extern __shared__ unsigned char enc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(enc_smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                        // [kStages]
  uint64_t* empty = bars + Cfg::kStages;        // [kStages]
  uint64_t* tm_full = bars + 2 * Cfg::kStages;  // [2]
  uint64_t* tm_empty = tm_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = a.W / kEncTileW, tiles_h = a.H / kEncTileH, tiles_n = a.Cout / BN;
  const int n_tiles = a.B * tiles_h * tiles_w * tiles_n;
  const int cblocks = a.Cin / kEncBK, n_kblocks = 9 * cblocks;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tm_full + s, 1);
      tc::mbar_init(tm_empty + s, 4);   // one arrive per epilogue warp
    }
    tc::fence_barrier_init();
    tc::fence_proxy_async();
  }
  if (warp == 2) tc::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> (image, patch row, patch column, channel block); the channel block runs fastest so
  // that CTAs working side by side share their activation tiles in L2
  auto decode = [&](int t, int& b, int& y0, int& x0, int& n0) {
    n0 = (t % tiles_n) * BN;
    const int m = t / tiles_n;
    x0 = (m % tiles_w) * kEncTileW;
    y0 = ((m / tiles_w) % tiles_h) * kEncTileH;
    b = m / (tiles_w * tiles_h);
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int b, y0, x0, n0;
        decode(t, b, y0, x0, n0);
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int cb = 0; cb < cblocks; ++cb) {
            unsigned char* sA = smem + (size_t)stage * Cfg::kStageBytes;
            tc::mbar_wait(empty + stage, phase ^ 1);
            tc::mbar_expect_tx(full + stage, (uint32_t)Cfg::kStageBytes);
            // the shifted patch; rows/columns outside the image arrive as zeros (= the padding)
            tc::tma_load_4d(sA, &map_in, full + stage, cb * kEncBK, x0 + dx, y0 + dy, b);
            tc::tma_load_2d(sA + kEncABytes, &map_w, full + stage, tap * a.Cin + cb * kEncBK, n0);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      constexpr uint32_t idesc = tc::instr_desc_f16(kEncBM, BN);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        tc::mbar_wait(tm_empty + as, aphase ^ 1);   // the epilogue drained this accumulator stage
        tc::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < n_kblocks; ++kb) {
          tc::mbar_wait(full + stage, phase);
          tc::tcgen05_fence_after();
          const uint32_t sA = tc::smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint64_t ad = tc::make_sw128_desc(sA), bd = tc::make_sw128_desc(sA + kEncABytes);
#pragma unroll
          for (int k4 = 0; k4 < kEncBK / kEncUK; ++k4)
            tc::umma_f16(d_tmem, ad + (uint64_t)(k4 * (kEncUK * 2 >> 4)), bd + (uint64_t)(k4 * (kEncUK * 2 >> 4)),
                         idesc, (kb | k4) != 0 ? 1u : 0u);
          tc::tcgen05_commit(empty + stage);        // the slot is reusable once these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        tc::tcgen05_commit(tm_full + as);           // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: bias, ReLU, store
    const int ew = warp & 3;                        // this warp's TMEM lane quadrant
    const int row = ew * 32 + lane;                 // pixel of the patch = TMEM lane
    const int yy = row / kEncTileW, xx = row % kEncTileW;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int b, y0, x0, n0;
      decode(t, b, y0, x0, n0);
      const int y = y0 + yy, x = x0 + xx;
      tc::mbar_wait(tm_full + as, aphase);
      tc::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        const int ch0 = n0 + c * 32;
        if (a.out_nhwc) {
          __align__(16) __half2 h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float f0 = fmaxf(__uint_as_float(v[2 * j]) + __ldg(a.bias + ch0 + 2 * j), 0.f);
            const float f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + __ldg(a.bias + ch0 + 2 * j + 1), 0.f);
            h[j] = __floats2half2_rn(f0, f1);
          }
          uint4* dst = reinterpret_cast<uint4*>(a.out_nhwc + (((size_t)b * a.H + y) * a.W + x) * a.Cout + ch0);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = reinterpret_cast<const uint4*>(h)[j];
        } else {
          float* dst = a.out_nchw + ((size_t)b * a.Cout + ch0) * ((size_t)a.H * a.W) + (size_t)y * a.W + x;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            dst[(size_t)j * a.H * a.W] = __uint_as_float(v[j]) + __ldg(a.bias + ch0 + j);
        }
      }
      // this warp's TMEM reads of the stage are complete: hand it back to the MMA
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tm_empty + as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// conv1_1 folded to one input channel: uint8 image [B][H][W] -> FP16 NHWC [B][H][W][64], ReLU.
// w1: [64][9] (weights summed over the three identical input channels, times 1/255), thread = pixel.
// The reference pads its canvas with cv::Mat::ones(h, w, CV_8UC3) * 255 (loop_detector.cpp:84), and
// Mat::ones sets only channel 0 of a multi-channel matrix: pad pixels are (255, 0, 0), the copied
// BEV image has three identical channels.  With `rois` ([B][4]: x0, y0, width, height of the copied
// image inside the plane) the taps that fall on padding use w1p: [64][9], channel 0's weights
// alone (times 1/255).  rois == nullptr: every pixel is image.
__global__ void __launch_bounds__(256)
enc_conv1_kernel(const uint8_t* __restrict__ img, int B, int H, int W, const float* __restrict__ w1,
                 const float* __restrict__ bias, __half* __restrict__ out,
                 const float* __restrict__ w1p = nullptr, const int* __restrict__ rois = nullptr) {
  __shared__ float w_s[64 * 9], b_s[64], wp_s[64 * 9];
  const bool padded = rois != nullptr && w1p != nullptr;
  for (int i = threadIdx.x; i < 64 * 9; i += 256) {
    w_s[i] = w1[i];
    wp_s[i] = padded ? w1p[i] : 0.f;
  }
  if (threadIdx.x < 64) b_s[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= (size_t)B * H * W) return;
  const int x = (int)(p % W), y = (int)((p / W) % H);
  const uint8_t* plane = img + (p - (size_t)y * W - x);
  int rx0 = 0, ry0 = 0, rx1 = W, ry1 = H;
  if (padded) {
    const int b = (int)(p / ((size_t)H * W));
    rx0 = rois[4 * b]; ry0 = rois[4 * b + 1]; rx1 = rx0 + rois[4 * b + 2]; ry1 = ry0 + rois[4 * b + 3];
  }
  float v[9], vp[9];
  bool any_pad = false;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    const bool in_img = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const bool in_roi = in_img && yy >= ry0 && yy < ry1 && xx >= rx0 && xx < rx1;
    const float val = in_img ? (float)plane[(size_t)yy * W + xx] : 0.f;
    v[tap] = in_roi ? val : 0.f;
    vp[tap] = in_img && !in_roi ? val : 0.f;
    any_pad |= in_img && !in_roi;
  }
  uint4* dst = reinterpret_cast<uint4*>(out + p * 64);
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    __align__(16) __half2 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int co = g * 8 + 2 * j + e;
        float acc = b_s[co];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) acc = fmaf(w_s[co * 9 + tap], v[tap], acc);
        if (any_pad) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) acc = fmaf(wp_s[co * 9 + tap], vp[tap], acc);
        }
        f[e] = fmaxf(acc, 0.f);
      }
      h[j] = __floats2half2_rn(f[0], f[1]);
    }
    dst[g] = *reinterpret_cast<const uint4*>(h);
  }
}

// per 16-bit lane signed maximum of two packed pairs (post-ReLU halves order like their bit patterns)
__device__ __forceinline__ unsigned enc_max2(unsigned p, unsigned q) {
  const short p0 = (short)(p & 0xFFFFu), p1 = (short)(p >> 16), q0 = (short)(q & 0xFFFFu), q1 = (short)(q >> 16);
  const unsigned lo = (unsigned short)(p0 > q0 ? p0 : q0), hi = (unsigned short)(p1 > q1 ? p1 : q1);
  return lo | (hi << 16);
}

// 2x2 max-pool, FP16 NHWC [B][H][W][C] -> [B][H/2][W/2][C]; thread = 8 channels of one output pixel
__global__ void __launch_bounds__(256)
enc_maxpool2_kernel(const __half* __restrict__ in, int B, int H, int W, int C, __half* __restrict__ out) {
  const int cg = C / 8, Ho = H / 2, Wo = W / 2;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)B * Ho * Wo * cg) return;
  const int g = (int)(i % cg);
  const size_t pix = i / cg;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((size_t)Wo * Ho));
  const uint4* src = reinterpret_cast<const uint4*>(in);
  const size_t r0 = (((size_t)b * H + 2 * oy) * W + 2 * ox) * cg + g, r1 = r0 + (size_t)W * cg;
  const uint4 a0 = __ldg(src + r0), a1 = __ldg(src + r0 + cg), a2 = __ldg(src + r1), a3 = __ldg(src + r1 + cg);
  uint4 m;
  m.x = enc_max2(enc_max2(a0.x, a1.x), enc_max2(a2.x, a3.x));
  m.y = enc_max2(enc_max2(a0.y, a1.y), enc_max2(a2.y, a3.y));
  m.z = enc_max2(enc_max2(a0.z, a1.z), enc_max2(a2.z, a3.z));
  m.w = enc_max2(enc_max2(a0.w, a1.w), enc_max2(a2.w, a3.w));
  reinterpret_cast<uint4*>(out)[i] = m;
}
// [enc-kernels-end]

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// FP16 NHWC activations [B][H][W][C]: box {64 channels, 16 x, 8 y, 1 image}, 128B swizzle, zero fill
bool make_act_map(CUtensorMap* map, const void* base, int B, int H, int W, int C) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kEncBK, (cuuint32_t)kEncTileW, (cuuint32_t)kEncTileH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// FP16 weights [Cout][9 Cin] (K-major): box {64, BN}
bool make_weight_map(CUtensorMap* map, const void* base, int Cout, int K, int bn) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kEncBK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int kEncLayers = 13;
// VGG16 features: output channels of the 13 convolutions; a pool follows layers 1, 3, 6, 9
// (and 12, which the reference drops together with the last ReLU)
constexpr int kEncCout[kEncLayers] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
constexpr bool kEncPoolAfter[kEncLayers] = {false, true, false, true, false, false, true,
                                            false, false, true, false, false, false};

inline int enc_bn(int cout) { return cout >= 256 ? 256 : cout; }

}  // namespace

}  // namespace gloc

struct gloc_encoder {
  int device = 0, H = 0, W = 0;
  float* d_w1 = nullptr;                          // folded conv1_1 [64][9]
  float* d_w1p = nullptr;                         // conv1_1 on padding: channel 0 alone [64][9]
  int* d_rois = nullptr;
  int rois_cap = 0;
  float* d_bias[gloc::kEncLayers] = {nullptr};    // [Cout]
  __half* d_w[gloc::kEncLayers] = {nullptr};      // [Cout][9 Cin] FP16 (layer 0 unused)
  __half* d_act[2] = {nullptr, nullptr};          // ping-pong activations
  size_t act_elems = 0;
  uint8_t* d_img = nullptr;
  float* d_feat = nullptr;
  float* d_desc = nullptr;
  size_t img_bytes = 0, feat_elems = 0, desc_elems = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  int sms = 0;
};

using gloc::fail;

namespace {

template <int BN>
cudaError_t launch_conv(const CUtensorMap& map_in, const CUtensorMap& map_w, const gloc::ConvArgs& a, int sms,
                        cudaStream_t st) {
  using Cfg = gloc::EncCfg<BN>;
  static unsigned long long attr_mask = 0;
  if (gloc::first_use_on_current_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(gloc::enc_conv3x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
  }
  const int n_tiles = a.B * (a.H / gloc::kEncTileH) * (a.W / gloc::kEncTileW) * (a.Cout / BN);
  gloc::enc_conv3x3_kernel<BN><<<std::min(n_tiles, sms), gloc::kEncThreads, Cfg::kSmemBytes, st>>>(map_in, map_w, a);
  return cudaGetLastError();
}

int forward_device(gloc_encoder* e, const uint8_t* d_img, int B, float* d_feat, const int32_t* rois = nullptr) {
  using namespace gloc;
  const size_t need = (size_t)B * e->H * e->W * 64;   // the largest activation: conv1 output
  if (need > e->act_elems) {
    for (int i = 0; i < 2; ++i) {
      if (e->d_act[i]) cudaFree(e->d_act[i]);
      e->d_act[i] = nullptr;
    }
    e->act_elems = 0;
    GLOC_CUDA_TRY(cudaMalloc(&e->d_act[0], need * sizeof(__half)));
    GLOC_CUDA_TRY(cudaMalloc(&e->d_act[1], need * sizeof(__half)));
    e->act_elems = need;
  }
  cudaStream_t st = e->stream;
  int H = e->H, W = e->W, C = 64, cur = 0;
  {
    const size_t px = (size_t)B * H * W;
    const int* d_rois = nullptr;
    if (rois) {
      for (int b = 0; b < B; ++b)
        if (rois[4 * b] < 0 || rois[4 * b + 1] < 0 || rois[4 * b + 2] < 0 || rois[4 * b + 3] < 0 ||
            rois[4 * b] + rois[4 * b + 2] > W || rois[4 * b + 1] + rois[4 * b + 3] > H)
          return fail(GLOC_ERR_INVALID, "gloc_enc_forward: image rectangle outside the plane");
      if (B > e->rois_cap) {
        if (e->d_rois) cudaFree(e->d_rois);
        e->d_rois = nullptr;
        e->rois_cap = 0;
        GLOC_CUDA_TRY(cudaMalloc(&e->d_rois, (size_t)B * 16));
        e->rois_cap = B;
      }
      GLOC_CUDA_TRY(cudaMemcpyAsync(e->d_rois, rois, (size_t)B * 16, cudaMemcpyHostToDevice, st));
      d_rois = e->d_rois;
    }
    enc_conv1_kernel<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(d_img, B, H, W, e->d_w1, e->d_bias[0], e->d_act[0],
                                                                   e->d_w1p, d_rois);
    GLOC_CUDA_TRY(cudaGetLastError());
    ++e->launches;
  }
  for (int l = 1; l < kEncLayers; ++l) {
    const int Cout = kEncCout[l], bn = enc_bn(Cout);
    CUtensorMap map_in, map_w;
    if (!make_act_map(&map_in, e->d_act[cur], B, H, W, C) || !make_weight_map(&map_w, e->d_w[l], Cout, 9 * C, bn))
      return fail(GLOC_ERR_CUDA, "gloc_enc_forward: cuTensorMapEncodeTiled failed");
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = C; a.Cout = Cout;
    a.bias = e->d_bias[l];
    const bool last = l == kEncLayers - 1;
    a.out_nhwc = last ? nullptr : e->d_act[cur ^ 1];
    a.out_nchw = last ? d_feat : nullptr;
    cudaError_t ce = bn == 64 ? launch_conv<64>(map_in, map_w, a, e->sms, st)
                   : bn == 128 ? launch_conv<128>(map_in, map_w, a, e->sms, st)
                               : launch_conv<256>(map_in, map_w, a, e->sms, st);
    GLOC_CUDA_TRY(ce);
    ++e->launches;
    cur ^= 1;
    C = Cout;
    if (kEncPoolAfter[l]) {
      const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
      enc_maxpool2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->d_act[cur], B, H, W, C, e->d_act[cur ^ 1]);
      GLOC_CUDA_TRY(cudaGetLastError());
      ++e->launches;
      cur ^= 1;
      H /= 2;
      W /= 2;
    }
  }
  return GLOC_OK;
}

}  // namespace

extern "C" {

int gloc_enc_create(gloc_encoder** out, int device, int height, int width, const float* const* conv_w,
                    const float* const* conv_b) {
  using namespace gloc;
  if (!out || !conv_w || !conv_b) return fail(GLOC_ERR_INVALID, "gloc_enc_create: null argument");
  *out = nullptr;
  for (int l = 0; l < kEncLayers; ++l)
    if (!conv_w[l] || !conv_b[l]) return fail(GLOC_ERR_INVALID, "gloc_enc_create: 13 weight and 13 bias arrays expected");
  // four pools, then 8 x 16 patches on the last feature map
  if (height < 128 || width < 256 || height % 128 != 0 || width % 256 != 0 || height > 4096 || width > 4096)
    return fail(GLOC_ERR_RANGE, "gloc_enc_create: height % 128 == 0 and width % 256 == 0 required (768 x 768 in the reference)");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    (void)cudaGetLastError();
    return fail(GLOC_ERR_CUDA, "gloc_enc_create: no CUDA device (there is no CPU fallback)");
  }
  if (device < 0 || device >= n_dev) return fail(GLOC_ERR_INVALID, "gloc_enc_create: bad device");
  int major = 0;
  GLOC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(GLOC_ERR_CUDA, "gloc_enc_create: device is not sm_100 (kernels are sm_100a only)");
  DeviceGuard scope(device);
  gloc_encoder* e = new gloc_encoder;
  e->device = device;
  e->H = height;
  e->W = width;
  e->sms = sm_count(device);
  cudaError_t ce = cudaStreamCreate(&e->stream);
  int cin = 3;
  for (int l = 0; l < kEncLayers && ce == cudaSuccess; ++l) {
    const int cout = kEncCout[l];
    ce = cudaMalloc(&e->d_bias[l], (size_t)cout * 4);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_bias[l], conv_b[l], (size_t)cout * 4, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) break;
    if (l == 0) {   // three identical input channels scaled by 1/255: one channel, weights summed
      std::vector<float> w1((size_t)64 * 9);
      for (int co = 0; co < 64; ++co)
        for (int tap = 0; tap < 9; ++tap) {
          float s = 0.f;
          for (int ci = 0; ci < 3; ++ci) s += conv_w[0][((size_t)co * 3 + ci) * 9 + tap];
          w1[(size_t)co * 9 + tap] = s / 255.f;
        }
      ce = cudaMalloc(&e->d_w1, w1.size() * 4);
      if (ce == cudaSuccess) ce = cudaMemcpy(e->d_w1, w1.data(), w1.size() * 4, cudaMemcpyHostToDevice);
      // padding of the reference's canvas is (255, 0, 0): only channel 0 sees it
      for (int co = 0; co < 64; ++co)
        for (int tap = 0; tap < 9; ++tap) w1[(size_t)co * 9 + tap] = conv_w[0][((size_t)co * 3 + 0) * 9 + tap] / 255.f;
      if (ce == cudaSuccess) ce = cudaMalloc(&e->d_w1p, w1.size() * 4);
      if (ce == cudaSuccess) ce = cudaMemcpy(e->d_w1p, w1.data(), w1.size() * 4, cudaMemcpyHostToDevice);
    } else {        // [Cout][Cin][3][3] float32 -> [Cout][tap][Cin] FP16
      std::vector<__half> w((size_t)cout * 9 * cin);
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int tap = 0; tap < 9; ++tap)
            w[((size_t)co * 9 + tap) * cin + ci] = __float2half_rn(conv_w[l][((size_t)co * cin + ci) * 9 + tap]);
      ce = cudaMalloc(&e->d_w[l], w.size() * sizeof(__half));
      if (ce == cudaSuccess) ce = cudaMemcpy(e->d_w[l], w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice);
    }
    cin = cout;
  }
  if (ce != cudaSuccess) {
    const std::string msg = std::string("gloc_enc_create: ") + cudaGetErrorString(ce);
    gloc_enc_destroy(e);
    return fail(GLOC_ERR_CUDA, msg);
  }
  *out = e;
  return GLOC_OK;
}

void gloc_enc_destroy(gloc_encoder* e) {
  if (!e) return;
  gloc::DeviceGuard scope(e->device);
  if (e->d_w1) cudaFree(e->d_w1);
  if (e->d_w1p) cudaFree(e->d_w1p);
  if (e->d_rois) cudaFree(e->d_rois);
  for (int l = 0; l < gloc::kEncLayers; ++l) {
    if (e->d_bias[l]) cudaFree(e->d_bias[l]);
    if (e->d_w[l]) cudaFree(e->d_w[l]);
  }
  for (int i = 0; i < 2; ++i)
    if (e->d_act[i]) cudaFree(e->d_act[i]);
  if (e->d_img) cudaFree(e->d_img);
  if (e->d_feat) cudaFree(e->d_feat);
  if (e->d_desc) cudaFree(e->d_desc);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int gloc_enc_feature_shape(const gloc_encoder* e, int* channels, int* n_loc) {
  if (!e) return fail(GLOC_ERR_INVALID, "gloc_enc_feature_shape: null encoder");
  if (channels) *channels = 512;
  if (n_loc) *n_loc = (e->H / 16) * (e->W / 16);
  return GLOC_OK;
}

int gloc_enc_forward_device(gloc_encoder* e, const uint8_t* d_images, int batch, float* d_feat) {
  if (!e || !d_images || !d_feat) return fail(GLOC_ERR_INVALID, "gloc_enc_forward_device: null argument");
  if (batch < 0) return fail(GLOC_ERR_INVALID, "gloc_enc_forward_device: negative batch");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(e->device);
  const int rc = forward_device(e, d_images, batch, d_feat);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GLOC_OK;
}

// host planes -> device, encoder -> e->d_feat (both buffers grown on demand); asynchronous on e->stream
static int stage_and_encode(gloc_encoder* e, const uint8_t* images, int batch, const int32_t* rois = nullptr) {
  const size_t in_bytes = (size_t)batch * e->H * e->W, out_elems = (size_t)batch * 512 * (e->H / 16) * (e->W / 16);
  if (in_bytes > e->img_bytes) {
    if (e->d_img) cudaFree(e->d_img);
    e->d_img = nullptr;
    e->img_bytes = 0;
    GLOC_CUDA_TRY(cudaMalloc(&e->d_img, in_bytes));
    e->img_bytes = in_bytes;
  }
  if (out_elems > e->feat_elems) {
    if (e->d_feat) cudaFree(e->d_feat);
    e->d_feat = nullptr;
    e->feat_elems = 0;
    GLOC_CUDA_TRY(cudaMalloc(&e->d_feat, out_elems * 4));
    e->feat_elems = out_elems;
  }
  GLOC_CUDA_TRY(cudaMemcpyAsync(e->d_img, images, in_bytes, cudaMemcpyHostToDevice, e->stream));
  return forward_device(e, e->d_img, batch, e->d_feat, rois);
}

int gloc_desc_extract(gloc_encoder* e, gloc_vlad_head* head, int out_dim, const uint8_t* images, int batch,
                      float* desc) {
  return gloc_desc_extract_padded(e, head, out_dim, images, nullptr, batch, desc);
}

int gloc_desc_extract_padded(gloc_encoder* e, gloc_vlad_head* head, int out_dim, const uint8_t* images,
                             const int32_t* rois, int batch, float* desc) {
  if (!e || !head || !images || !desc) return fail(GLOC_ERR_INVALID, "gloc_desc_extract: null argument");
  if (batch < 0 || out_dim < 1) return fail(GLOC_ERR_INVALID, "gloc_desc_extract: bad batch / out_dim");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(e->device);
  int rc = stage_and_encode(e, images, batch, rois);
  if (rc != GLOC_OK) return rc;
  const size_t n_desc = (size_t)batch * out_dim;
  if (n_desc > e->desc_elems) {
    if (e->d_desc) cudaFree(e->d_desc);
    e->d_desc = nullptr;
    e->desc_elems = 0;
    GLOC_CUDA_TRY(cudaMalloc(&e->d_desc, n_desc * 4));
    e->desc_elems = n_desc;
  }
  GLOC_CUDA_TRY(cudaStreamSynchronize(e->stream));   // the head runs on its own stream
  rc = gloc_vlad_forward_device(head, e->d_feat, batch, (e->H / 16) * (e->W / 16), e->d_desc);   // synchronous
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpy(desc, e->d_desc, n_desc * 4, cudaMemcpyDeviceToHost));
  return GLOC_OK;
}

int gloc_enc_forward(gloc_encoder* e, const uint8_t* images, int batch, float* feat) {
  return gloc_enc_forward_padded(e, images, nullptr, batch, feat);
}

int gloc_enc_forward_padded_device(gloc_encoder* e, const uint8_t* d_images, const int32_t* rois, int batch,
                                   float* d_feat) {
  if (!e || !d_images || !d_feat) return fail(GLOC_ERR_INVALID, "gloc_enc_forward_device: null argument");
  if (batch < 0) return fail(GLOC_ERR_INVALID, "gloc_enc_forward_device: negative batch");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(e->device);
  const int rc = forward_device(e, d_images, batch, d_feat, rois);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GLOC_OK;
}

int gloc_enc_forward_padded(gloc_encoder* e, const uint8_t* images, const int32_t* rois, int batch, float* feat) {
  if (!e || !images || !feat) return fail(GLOC_ERR_INVALID, "gloc_enc_forward: null argument");
  if (batch < 0) return fail(GLOC_ERR_INVALID, "gloc_enc_forward: negative batch");
  if (batch == 0) return GLOC_OK;
  gloc::DeviceGuard scope(e->device);
  const size_t out_elems = (size_t)batch * 512 * (e->H / 16) * (e->W / 16);
  const int rc = stage_and_encode(e, images, batch, rois);
  if (rc != GLOC_OK) return rc;
  GLOC_CUDA_TRY(cudaMemcpyAsync(feat, e->d_feat, out_elems * 4, cudaMemcpyDeviceToHost, e->stream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GLOC_OK;
}

uint64_t gloc_enc_kernel_launches(const gloc_encoder* e) { return e ? e->launches : 0; }

}  // extern "C"
