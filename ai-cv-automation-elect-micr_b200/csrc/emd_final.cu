// The last layer of the denoiser on the tensor cores: 3x3 convolution 64 -> 1 channel, BatchNorm, ReLU6 and the in-graph
// clip (misc_py/denoiser-multi-gpu.py:531-538), FP32 image out.
//
// A 64 -> 1 convolution is 576 multiply-adds per pixel for 128 bytes read -- on the CUDA cores that is ~4x the time the
// bytes take from HBM.  Here the channel reduction runs as one small GEMM per tile instead:
//     G[q, t] = sum_c X[q, c] * W[t, c]        q = the 10 x 18 halo pixels of an 8 x 16 output tile, t = the 9 taps
//     out[p]  = sum_t G[p + offset(t), t]
// One 4-D TMA box brings the halo into shared memory as 180 rows of 128 bytes (128-byte swizzle) -- which is already the
// K-major A operand of a tcgen05.mma with M = 2 x 128 rows, K = 64, N = 16 (9 used); the 16 x 64 weight operand stays
// resident.  The accumulators (2 x 16 TMEM columns, double-buffered) are read back by four warps, scattered to shared
// memory as G[t][q], and each thread adds the nine shifted values of its output pixel.  Per tile: 8 MMAs, ~60 CUDA-core
// instructions per thread, 23 KB from HBM; zero padding (TF SAME) is the TMA's out-of-range fill.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "emd_kernels.h"
#include "emd_tma.h"

namespace emd {
namespace {

using namespace ptx;

constexpr int kTH = 8, kTW = 16;                      // output tile
constexpr int kHaloH = kTH + 2, kHaloW = kTW + 2;     // 10 x 18
constexpr int kHaloPx = kHaloH * kHaloW;              // 180 rows of the A operand
constexpr int kC = 64;                                // input channels = one 128-byte swizzle row
constexpr int kHaloBytes = kHaloPx * kC * 2;          // 23040 bytes per TMA box
constexpr int kStageBytes = 256 * kC * 2;             // two M = 128 operand tiles (rows 180..255 are never read back)
constexpr int kStages = 6;
constexpr int kNB = 16;                               // N of the MMA: 9 taps padded to 16
constexpr int kGPitch = 192;                          // floats per tap row of G
constexpr int kEpiThreads = 128, kMmaWarp = 4, kProdWarp = 5, kThreads = 192;
constexpr int kTmemCols = 64;                         // 2 buffers x (2 M tiles x 16 columns)

struct FinalArgs {
  const float* w;        // FP32 [9][64], values already rounded to the 16-bit operand type
  float* out;            // [N][H][W] FP32
  int N, H, W;
  int tiles_x, tiles_per_img, tiles;
  FastDiv d_tpi, d_tx;
  float scale, shift;
  int relu6, clip01;
};

template <typename T> struct Fmt;
template <> struct Fmt<__nv_bfloat16> {
  static constexpr uint32_t k = 1;
  static __device__ __forceinline__ uint16_t from(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
};
template <> struct Fmt<__half> {
  static constexpr uint32_t k = 0;
  static __device__ __forceinline__ uint16_t from(float f) { return __half_as_ushort(__float2half_rn(f)); }
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tile_origin(const FinalArgs& a, int tile, int& n, int& y0, int& x0) {
  n = (int)fdiv((uint32_t)tile, a.d_tpi);
  const int rem = tile - n * a.tiles_per_img;
  const int by = (int)fdiv((uint32_t)rem, a.d_tx);
  y0 = by * kTH;
  x0 = (rem - by * a.tiles_x) * kTW;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
final_umma_kernel(const __grid_constant__ FinalArgs a, const __grid_constant__ CUtensorMap tmap_in) {
  extern __shared__ uint8_t smem_raw[];
  griddep_launch();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~(uintptr_t)1023u);
  uint8_t* s_b = smem + (size_t)kStages * kStageBytes;                        // 16 rows x 128 B, swizzled (2 KB)
  float* s_g = reinterpret_cast<float*>(s_b + 2048);                          // [2][9][kGPitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_g + 2 * 9 * kGPitch);
  const uint32_t bar_hfull = smem_u32(bars), bar_hempty = bar_hfull + 8u * kStages, bar_tfull = bar_hempty + 8u * kStages,
                 bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform roles

  // weight operand: row = tap (rows 9..15 zero), 64 channels = eight 16-byte chunks, chunk j of row r at (j ^ (r & 7))
  for (int i = threadIdx.x; i < kNB * kC; i += kThreads) {
    const int r = i >> 6, c = i & 63;
    const float v = r < 9 ? a.w[r * kC + c] : 0.f;
    *reinterpret_cast<uint16_t*>(s_b + r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1))) = Fmt<T>::from(v);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_hfull + 8u * s, 1); mbar_init(bar_hempty + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, kEpiThreads / 32); }
    fence_barrier_init();
    prefetch_tmap(&tmap_in);
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  fence_proxy_async();          // the generic-proxy stores of the weight operand become visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();               // the layer before has finished (emd_tma.h)
  const int first = blockIdx.x, step = gridDim.x;

  if (warp == kProdWarp) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = first; tile < a.tiles; tile += step) {
        int n, y0, x0;
        tile_origin(a, tile, n, y0, x0);
        mbar_wait(bar_hempty + 8u * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_hfull + 8u * s, kHaloBytes);
        tma_load_4d(smem_u32(smem) + (uint32_t)s * kStageBytes, &tmap_in, 0, x0 - 1, y0 - 1, n, bar_hfull + 8u * s);
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t idesc = (1u << 4) | (Fmt<T>::k << 7) | (Fmt<T>::k << 10) | ((uint32_t)(kNB >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(smem)), b_lo = sdesc_lo(smem_u32(s_b));
    int s = 0, tcount = 0;
    uint32_t ph = 0;
    for (int tile = first; tile < a.tiles; tile += step, ++tcount) {
      const int acc = tcount & 1;
      mbar_wait(bar_tempty + 8u * acc, (uint32_t)(((tcount >> 1) & 1) ^ 1));
      mbar_wait(bar_hfull + 8u * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (kStageBytes >> 4);
        const uint32_t d = tmem_base + (uint32_t)(acc * 2 * kNB);
        umma_kblock<false, 4>(d, a_lo, b_lo, idesc, 0u);                               // halo rows 0..127
        umma_kblock<false, 4>(d + kNB, a_lo + ((128 * kC * 2) >> 4), b_lo, idesc, 0u);  // halo rows 128..255
        umma_commit(bar_hempty + 8u * s);
        umma_commit(bar_tfull + 8u * acc);
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== accumulator read-back, tap gather, BN / ReLU6 / clip, store =====================
    const int q0 = warp * 32 + lane, q1 = 128 + q0;     // halo pixels (A rows) whose accumulators this thread reads
    const int oy = q0 >> 4, ox = q0 & 15;               // output pixel of the tile this thread writes
    int tcount = 0;
    for (int tile = first; tile < a.tiles; tile += step, ++tcount) {
      const int acc = tcount & 1;
      int n, y0, x0;
      tile_origin(a, tile, n, y0, x0);
      mbar_wait(bar_tfull + 8u * acc, (uint32_t)((tcount >> 1) & 1));
      tc_fence_after();
      uint32_t v0[16], v1[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 2 * kNB);
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + kNB, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
      float* g = s_g + acc * 9 * kGPitch;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        g[t * kGPitch + q0] = __uint_as_float(v0[t]);
        if (q1 < kHaloPx) g[t * kGPitch + q1] = __uint_as_float(v1[t]);
      }
      named_bar_sync(1, kEpiThreads);   // one barrier per tile: the other G buffer is rewritten only after everyone passed this one
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) sum += g[t * kGPitch + (oy + t / 3) * kHaloW + ox + t % 3];
      float y = fmaf(sum, a.scale, a.shift);
      if (a.relu6) y = fminf(fmaxf(y, 0.f), 6.f);
      if (a.clip01) y = fminf(fmaxf(y, 0.f), 1.f);
      a.out[((size_t)n * a.H + (y0 + oy)) * a.W + x0 + ox] = y;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <typename T>
cudaError_t launch_t(const FinalArgs& a, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t s) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t r = cudaFuncSetAttribute(final_umma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r != cudaSuccess) return r;
    attr_dev = dev;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, final_umma_kernel<T>, a, tmap);
}

}  // namespace

bool final_umma_supported(const ConvParams& p, int et) {
  return tuning().final_umma && final_tma_supported(p, et);   // A/B switch: the CUDA-core kernel of emd_dw.cu
}

cudaError_t launch_final_umma(const ConvParams& c, float scale, float shift, int et, int num_sms, cudaStream_t s) {
  FinalArgs a;
  a.w = c.w;
  a.out = reinterpret_cast<float*>(c.out.ptr);
  a.N = c.N; a.H = c.MH; a.W = c.MW;
  a.tiles_x = c.MW / kTW;
  a.tiles_per_img = (c.MH / kTH) * a.tiles_x;
  a.tiles = c.N * a.tiles_per_img;
  a.d_tpi = make_fastdiv((uint32_t)a.tiles_per_img); a.d_tx = make_fastdiv((uint32_t)a.tiles_x);
  a.scale = scale; a.shift = shift; a.relu6 = c.relu6; a.clip01 = c.clip01;
  CUtensorMap tmap;
  void* base = reinterpret_cast<char*>(c.in.ptr) + (size_t)c.in.coff * 2;
  if (!tma_encode_nhwc(&tmap, et == ET_BF16, base, kC, c.in.W, c.in.H, c.N, c.in.pitch, kC, kHaloW, kHaloH, 1, true))
    return cudaErrorInvalidValue;
  const size_t smem = 1024 + (size_t)kStages * kStageBytes + 2048 + 2 * 9 * kGPitch * sizeof(float) + (2 * kStages + 4) * 8 + 16;
  const int grid = a.tiles < num_sms ? a.tiles : num_sms;
  return et == ET_BF16 ? launch_t<__nv_bfloat16>(a, tmap, grid, smem, s) : launch_t<__half>(a, tmap, grid, smem, s);
}

}  // namespace emd
