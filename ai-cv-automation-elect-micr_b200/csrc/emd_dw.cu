// emd_dw.cu -- depthwise 3x3 convolution (stride 1, rate 1) for the 16-bit modes, TMA-fed.
//
// The DepthwiseConv2dNative half of slim.separable_convolution2d (DMG:253-273) is memory-bound
// (9 MACs per element).  One persistent CTA walks (8 x 16 pixel tile, 64-channel chunk) items:
//   producer warp : one 4-D TMA load per item brings the 10 x 18 pixel halo of the chunk into shared
//                   memory (zero fill outside the image = SAME padding), 3-deep mbarrier ring
//   8 math warps  : thread = (8 channels, one pixel column, 4 output rows); the 3x3 window slides
//                   down the column in registers (FP32 accumulate), each quarter-warp reads one
//                   full 128-byte pixel row per LDS.128 and writes one full 128-byte segment per STG.128
#include "emd_kernels.h"
#include "emd_tma.h"

namespace emd {
namespace {

constexpr int kTH = 8, kTW = 16, kHaloH = kTH + 2, kHaloW = kTW + 2;
constexpr int kChunk = 64;
constexpr int kStageBytes = kHaloH * kHaloW * kChunk * 2;  // 23040
constexpr int kMaxStages = 8;     // stage count is a launch argument (even, 4..8): two groups own the stages of their parity
constexpr int kGroupWarps = 8;
constexpr int kMathThreads = 2 * kGroupWarps * 32;   // two groups of 8 warps take alternate items

struct DwTmaArgs {
  DwParams p;
  int tiles_x, tiles_per_img, nchunks, items;
  FastDiv d_chunks, d_tpi, d_tx;   // by nchunks, tiles_per_img, tiles_x
  // kReduce (final 3x3 conv 64 -> 1, DMG:531-538): p.w holds the [9][64] kernel, the 64 products are
  // summed over channels, then scale/shift, ReLU6, clip and an FP32 store
  float scale, shift;
  int relu6, clip01;
  int stages;   // depth of the halo ring (even)
  int cols;     // thread mapping: 1 = 2 channels x 4 columns x 4 rows (conflict-free 128-byte LDS, 40 % fewer unpacks), 0 = 4 channels x 1 column x 8 rows
};

template <typename T> struct Up;
template <> struct Up<__nv_bfloat16> {
  static __device__ __forceinline__ float2 up(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct Up<__half> {
  static __device__ __forceinline__ float2 up(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

// Thread = 4 channels x 1 pixel column x 8 rows of an (8 x 16 pixel, 64-channel) item: 16 lanes read one 128-byte
// halo pixel per LDS.64, the 3x3 window slides down the column in registers (FP32, packed FFMA2), and the same
// 16 lanes write one full 128-byte output pixel.
template <typename T, bool kReduce>
__global__ void __launch_bounds__(kMathThreads + 32, 1) dw_tma_kernel(const __grid_constant__ DwTmaArgs a,
                                                                      const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  ptx::griddep_launch();
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* smem = smem_raw + (base - raw);
  const int kStages = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  const uint32_t bar_full = ptx::smem_u32(bars), bar_empty = bar_full + 8u * kStages;
  const DwParams& p = a.p;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(bar_full + 8u * s, 1); ptx::mbar_init(bar_empty + 8u * s, kGroupWarps); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
  }
  __syncthreads();
  ptx::griddep_wait();   // the layer before has finished (emd_tma.h)
  // Work split with CHUNK AFFINITY: worker = (CTA, math group); worker w keeps channel chunk w % nchunks for the whole launch
  // and walks the tiles j, j + Wc, ... (j = its rank among the Wc workers of that chunk), so its 36 depthwise weights are
  // loaded once instead of once per item and the per-item index arithmetic is one tile decode.
  const int n_workers = 2 * (int)gridDim.x, n_tiles = a.items / a.nchunks;
  auto worker = [&](int grp, int& c, int& j, int& wc, int& cnt) {
    const int w = 2 * (int)blockIdx.x + grp;
    c = w % a.nchunks; j = w / a.nchunks;
    wc = (n_workers - c + a.nchunks - 1) / a.nchunks;
    cnt = j < n_tiles ? (n_tiles - j + wc - 1) / wc : 0;
  };

  if (threadIdx.x >= kMathThreads) {
    // ---------------- producer ----------------
    if (threadIdx.x == kMathThreads) {
      int c[2], j[2], wc[2], cnt[2];
      worker(0, c[0], j[0], wc[0], cnt[0]);
      worker(1, c[1], j[1], wc[1], cnt[1]);
      const int rounds = cnt[0] > cnt[1] ? cnt[0] : cnt[1];
      for (int k = 0; k < rounds; ++k) {
        for (int g = 0; g < 2; ++g) {
          if (k >= cnt[g]) continue;
          const int tile = j[g] + k * wc[g];
          const int n_img = tile / a.tiles_per_img;
          const int rem = tile - n_img * a.tiles_per_img;
          const int by = rem / a.tiles_x, bx = rem - by * a.tiles_x;
          const int q = g + 2 * k, s = q % kStages;          // group g owns the stages of its own parity (the stage count is even)
          const uint32_t ph = (uint32_t)((q / kStages) & 1);
          ptx::mbar_wait(bar_empty + 8u * s, ph ^ 1u);
          ptx::mbar_arrive_expect_tx(bar_full + 8u * s, kStageBytes);
          ptx::tma_load_4d(base + (uint32_t)s * kStageBytes, &tmap, c[g] * kChunk, bx * kTW - 1, by * kTH - 1, n_img, bar_full + 8u * s);
        }
      }
    }
    return;
  }

  // ---------------- math warps ----------------
  const int grp = threadIdx.x >> 8, tg = threadIdx.x & 255, lane = threadIdx.x & 31;
  int c, j0, wc, cnt;
  worker(grp, c, j0, wc, cnt);
  if (!kReduce && a.cols) {
    // Thread = 2 channels x 4 pixel columns x 4 rows (the mapping of the fused kernel's depthwise producer, emd_fused.cu): a
    // warp's lanes are the 32 channel pairs of one pixel, every LDS.32 / STG.32 is one 128-byte line; halo row j adds into
    // output rows j-2 .. j.  Same tap order as the other mapping: bit-identical results.
    const int cp = tg & 31, sub = tg >> 5;
    const int x0 = (sub & 3) * 4, y0 = (sub >> 2) * 4;
    const int ch = c * kChunk + cp * 2;
    const bool ch_ok = ch < p.in.C;
    float2 w[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t] = (ch_ok && cnt > 0) ? __ldg(reinterpret_cast<const float2*>(p.w + t * p.in.C + ch)) : make_float2(0.f, 0.f);
    int s = grp % kStages;
    uint32_t ph = (uint32_t)((grp / kStages) & 1);
    const uint32_t ostep = (uint32_t)p.out.W * (uint32_t)p.out.pitch, opitch = (uint32_t)p.out.pitch;   // element offsets inside a tile fit 32 bits: one IMAD.WIDE.U32 per store address
    for (int k = 0; k < cnt; ++k) {
      const int tile = j0 + k * wc;
      const int n_img = (int)fdiv((uint32_t)tile, a.d_tpi);
      const int rem = tile - n_img * a.tiles_per_img;
      const int by = (int)fdiv((uint32_t)rem, a.d_tx), bx = rem - by * a.tiles_x;
      ptx::mbar_wait(bar_full + 8u * s, ph);
      const uint8_t* hb = smem + (size_t)s * kStageBytes + (y0 * kHaloW + x0) * (kChunk * 2) + cp * 4;
      const size_t opix = ((size_t)n_img * p.out.H + by * kTH + y0) * p.out.W + bx * kTW + x0;
      T* orow = reinterpret_cast<T*>(p.out.ptr) + opix * p.out.pitch + p.out.coff + ch;
      float2 acc[3][4];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        float2 x[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = Up<T>::up(*reinterpret_cast<const uint32_t*>(hb + (j * kHaloW + i) * (kChunk * 2)));
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int r = j - ky;
          if (r < 0 || r > 3) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2& d = acc[r % 3][i];
            d = ky == 0 ? ptx::fmul2(x[i], w[0]) : ptx::ffma2(x[i], w[ky * 3], d);
            d = ptx::ffma2(x[i + 1], w[ky * 3 + 1], d);
            d = ptx::ffma2(x[i + 2], w[ky * 3 + 2], d);
          }
        }
        if (j >= 2 && ch_ok) {
          const int r = j - 2;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint32_t*>(orow + ((uint32_t)r * ostep + (uint32_t)i * opitch)) = Up<T>::pack(acc[r % 3][i].x, acc[r % 3][i].y);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_empty + 8u * s);
      s += 2;
      if (s >= kStages) { s -= kStages; ph ^= 1u; }
    }
    return;
  }
  const int cq = tg & 15, col = tg >> 4;
  const int ch = c * kChunk + cq * 4;
  const bool ch_ok = ch < p.in.C;
  float2 w[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {   // this thread's 9 x 4 depthwise weights
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch_ok && cnt > 0) w0 = __ldg(reinterpret_cast<const float4*>(p.w + t * p.in.C + ch));
    w[t][0] = make_float2(w0.x, w0.y); w[t][1] = make_float2(w0.z, w0.w);
  }
  int s = grp % kStages;
  uint32_t ph = (uint32_t)((grp / kStages) & 1);
  for (int k = 0; k < cnt; ++k) {
    const int tile = j0 + k * wc;
    const int n_img = (int)fdiv((uint32_t)tile, a.d_tpi);
    const int rem = tile - n_img * a.tiles_per_img;
    const int by = (int)fdiv((uint32_t)rem, a.d_tx), bx = rem - by * a.tiles_x;
    ptx::mbar_wait(bar_full + 8u * s, ph);
    const uint8_t* hb = smem + (size_t)s * kStageBytes + col * (kChunk * 2) + cq * 8;
    float2 win[3][3][2];
    auto load_row = [&](int hy, float2 (&dst)[3][2]) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint2 u = *reinterpret_cast<const uint2*>(hb + (hy * kHaloW + kx) * (kChunk * 2));
        dst[kx][0] = Up<T>::up(u.x); dst[kx][1] = Up<T>::up(u.y);
      }
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
    const size_t opix = ((size_t)n_img * p.out.H + by * kTH) * p.out.W + bx * kTW + col;
    T* orow = reinterpret_cast<T*>(p.out.ptr) + opix * p.out.pitch + p.out.coff + ch;   // this thread's 4 channels of output row 0
    const size_t ostep = (size_t)p.out.W * p.out.pitch;
#pragma unroll
    for (int i = 0; i < kTH; ++i) {
      load_row(i + 2, win[(i + 2) % 3]);
      float2 acc0 = ptx::fmul2(win[i % 3][0][0], w[0][0]), acc1 = ptx::fmul2(win[i % 3][0][1], w[0][1]);
#pragma unroll
      for (int t = 1; t < 9; ++t) {
        acc0 = ptx::ffma2(win[(i + t / 3) % 3][t % 3][0], w[t][0], acc0);
        acc1 = ptx::ffma2(win[(i + t / 3) % 3][t % 3][1], w[t][1], acc1);
      }
      if (kReduce) {
        float t = (acc0.x + acc0.y) + (acc1.x + acc1.y);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);   // the 16 lanes of a half-warp hold the 16 channel quads of one pixel
        if (cq == 0) {
          float v = fmaf(t, a.scale, a.shift);
          if (a.relu6) v = fminf(fmaxf(v, 0.f), 6.f);
          if (a.clip01) v = fminf(fmaxf(v, 0.f), 1.f);
          reinterpret_cast<float*>(p.out.ptr)[opix + (size_t)i * p.out.W] = v;
        }
      } else if (ch_ok) {
        *reinterpret_cast<uint2*>(orow + i * ostep) = make_uint2(Up<T>::pack(acc0.x, acc0.y), Up<T>::pack(acc1.x, acc1.y));
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_empty + 8u * s);
    s += 2;
    if (s >= kStages) { s -= kStages; ph ^= 1u; }
  }
}

// ---------------------------------------------------------------------------------------------
// stride 2 (the DepthwiseConv2dNative of the cnnN_strided blocks, DMG:402/417/432/447; TF SAME on even sizes pads
// 0 before / 1 after, so output (y, x) reads input rows 2y..2y+2, columns 2x..2x+2).  Item = 8 x 8 output pixels of a
// 64-channel chunk; TMA brings the 17 x 17 input patch.  Four groups of four warps, group g owns pipeline stage g;
// thread = 4 channels x 1 output column x 8 output rows, the bottom window row is carried to the next output row.
// ---------------------------------------------------------------------------------------------
constexpr int kS2T = 8, kS2Halo = 2 * kS2T + 1;
constexpr int kS2StageBytes = kS2Halo * kS2Halo * kChunk * 2;   // 36992
constexpr int kS2Groups = 4, kS2GroupThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kS2Groups* kS2GroupThreads + 32, 1) dw_s2_tma_kernel(const __grid_constant__ DwTmaArgs a,
                                                                                       const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  ptx::griddep_launch();
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kS2Groups * kS2StageBytes);
  const uint32_t bar_full = ptx::smem_u32(bars), bar_empty = bar_full + 8u * kS2Groups;
  const DwParams& p = a.p;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kS2Groups; ++s) { ptx::mbar_init(bar_full + 8u * s, 1); ptx::mbar_init(bar_empty + 8u * s, kS2GroupThreads / 32); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
  }
  __syncthreads();
  ptx::griddep_wait();   // the layer before has finished (emd_tma.h)
  const int my_items = (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x >= kS2Groups * kS2GroupThreads) {
    if (threadIdx.x == kS2Groups * kS2GroupThreads) {   // producer
      for (int k = 0, item = blockIdx.x; k < my_items; ++k, item += gridDim.x) {
        const int tile = item / a.nchunks, c = item - tile * a.nchunks;
        const int n_img = tile / a.tiles_per_img;
        const int rem = tile - n_img * a.tiles_per_img;
        const int by = rem / a.tiles_x, bx = rem - by * a.tiles_x;
        const int s = k % kS2Groups;
        ptx::mbar_wait(bar_empty + 8u * s, (uint32_t)(((k / kS2Groups) & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(bar_full + 8u * s, kS2StageBytes);
        ptx::tma_load_4d(base + (uint32_t)s * kS2StageBytes, &tmap, c * kChunk, 2 * bx * kS2T, 2 * by * kS2T, n_img, bar_full + 8u * s);
      }
    }
    return;
  }

  const int grp = threadIdx.x >> 7, tg = threadIdx.x & 127, lane = threadIdx.x & 31;
  const int cq = tg & 15, col = tg >> 4;      // 4-channel quad, output column 0..7
  int cur_c = -1;
  float2 w[9][2];
  for (int k = grp; k < my_items; k += kS2Groups) {
    const int item = blockIdx.x + k * gridDim.x;
    const int tile = (int)fdiv((uint32_t)item, a.d_chunks), c = item - tile * a.nchunks;
    const int n_img = (int)fdiv((uint32_t)tile, a.d_tpi);
    const int rem = tile - n_img * a.tiles_per_img;
    const int by = (int)fdiv((uint32_t)rem, a.d_tx), bx = rem - by * a.tiles_x;
    const int ch = c * kChunk + cq * 4;
    const bool ch_ok = ch < p.in.C;
    if (c != cur_c) {
      cur_c = c;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch_ok) w0 = __ldg(reinterpret_cast<const float4*>(p.w + t * p.in.C + ch));
        w[t][0] = make_float2(w0.x, w0.y); w[t][1] = make_float2(w0.z, w0.w);
      }
    }
    ptx::mbar_wait(bar_full + 8u * grp, (uint32_t)((k / kS2Groups) & 1));
    const uint8_t* hb = smem + (size_t)grp * kS2StageBytes + (2 * col) * (kChunk * 2) + cq * 8;
    float2 win[3][3][2];
    auto load_row = [&](int hy, float2 (&dst)[3][2]) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint2 u = *reinterpret_cast<const uint2*>(hb + (hy * kS2Halo + kx) * (kChunk * 2));
        dst[kx][0] = Up<T>::up(u.x); dst[kx][1] = Up<T>::up(u.y);
      }
    };
    load_row(0, win[0]);
    const size_t opix = ((size_t)n_img * p.out.H + by * kS2T) * p.out.W + bx * kS2T + col;
    T* orow = reinterpret_cast<T*>(p.out.ptr) + opix * p.out.pitch + p.out.coff + ch;
    const size_t ostep = (size_t)p.out.W * p.out.pitch;
#pragma unroll
    for (int i = 0; i < kS2T; ++i) {
      // window rows for output row i: halo rows 2i, 2i+1, 2i+2 live in win[(2i)%3], win[(2i+1)%3], win[(2i+2)%3]
      load_row(2 * i + 1, win[(2 * i + 1) % 3]);
      load_row(2 * i + 2, win[(2 * i + 2) % 3]);
      float2 acc0 = ptx::fmul2(win[(2 * i) % 3][0][0], w[0][0]), acc1 = ptx::fmul2(win[(2 * i) % 3][0][1], w[0][1]);
#pragma unroll
      for (int t = 1; t < 9; ++t) {
        acc0 = ptx::ffma2(win[(2 * i + t / 3) % 3][t % 3][0], w[t][0], acc0);
        acc1 = ptx::ffma2(win[(2 * i + t / 3) % 3][t % 3][1], w[t][1], acc1);
      }
      if (ch_ok) *reinterpret_cast<uint2*>(orow + i * ostep) = make_uint2(Up<T>::pack(acc0.x, acc0.y), Up<T>::pack(acc1.x, acc1.y));
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_empty + 8u * grp);
  }
}

}  // namespace

bool dw_tma_supported(const DwParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (p.in_f32 || p.stride != 1 || p.rate != 1 || p.pad != 1) return false;
  if (p.OH % kTH || p.OW % kTW) return false;
  if ((p.in.C & 7) || (p.in.pitch & 7) || (p.in.coff & 7) || (p.out.pitch & 7) || (p.out.coff & 7)) return false;
  return tma_encoder() != nullptr;
}

template <typename T, bool kReduce>
static cudaError_t launch_t(const DwTmaArgs& a, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t s) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t r = cudaFuncSetAttribute(dw_tma_kernel<T, kReduce>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r != cudaSuccess) return r;
    attr_dev = dev;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(kMathThreads + 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, dw_tma_kernel<T, kReduce>, a, tmap);
}

static cudaError_t launch_common(DwTmaArgs& a, int et, bool reduce, int num_sms, cudaStream_t s) {
  const DwParams& p = a.p;
  a.tiles_x = p.OW / kTW;
  a.tiles_per_img = (p.OH / kTH) * a.tiles_x;
  a.nchunks = (p.in.C + kChunk - 1) / kChunk;
  a.items = p.N * a.tiles_per_img * a.nchunks;
  a.d_chunks = make_fastdiv((uint32_t)a.nchunks); a.d_tpi = make_fastdiv((uint32_t)a.tiles_per_img); a.d_tx = make_fastdiv((uint32_t)a.tiles_x);
  CUtensorMap tmap;
  void* base = reinterpret_cast<char*>(p.in.ptr) + (size_t)p.in.coff * 2;
  if (!tma_encode_nhwc(&tmap, et == ET_BF16, base, p.in.C, p.in.W, p.in.H, p.N, p.in.pitch, kChunk, kHaloW, kHaloH, 1, false))
    return cudaErrorInvalidValue;
  int st = tuning().dw_stages ? tuning().dw_stages : 4;
  st = st < 2 ? 2 : st > kMaxStages ? kMaxStages : st & ~1;
  a.stages = st;
  const size_t smem = 128 + (size_t)kMaxStages * kStageBytes + 2 * kMaxStages * 8;   // attribute set once for the deepest ring
  const int grid = a.items < num_sms ? a.items : num_sms;
  if (et == ET_BF16)
    return reduce ? launch_t<__nv_bfloat16, true>(a, tmap, grid, smem, s) : launch_t<__nv_bfloat16, false>(a, tmap, grid, smem, s);
  return reduce ? launch_t<__half, true>(a, tmap, grid, smem, s) : launch_t<__half, false>(a, tmap, grid, smem, s);
}

cudaError_t launch_dw_tma(const DwParams& p, int et, int num_sms, cudaStream_t s) {
  DwTmaArgs a;
  a.p = p; a.scale = 1.f; a.shift = 0.f; a.relu6 = a.clip01 = 0;
  a.cols = tuning().dw_cols;
  return launch_common(a, et, false, num_sms, s);
}

// stride 2, TF SAME on even input sizes (pad 0 before / 1 after)
bool dw_s2_tma_supported(const DwParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (p.in_f32 || p.stride != 2 || p.rate != 1 || p.pad != 0) return false;
  if (p.OH % kS2T || p.OW % kS2T || p.in.H != 2 * p.OH || p.in.W != 2 * p.OW) return false;
  if ((p.in.C & 7) || (p.in.pitch & 7) || (p.in.coff & 7) || (p.out.pitch & 7) || (p.out.coff & 7)) return false;
  return tma_encoder() != nullptr;
}

template <typename T>
static cudaError_t launch_s2_t(const DwTmaArgs& a, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t s) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t r = cudaFuncSetAttribute(dw_s2_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r != cudaSuccess) return r;
    attr_dev = dev;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(kS2Groups * kS2GroupThreads + 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, dw_s2_tma_kernel<T>, a, tmap);
}

cudaError_t launch_dw_s2_tma(const DwParams& p, int et, int num_sms, cudaStream_t s) {
  DwTmaArgs a;
  a.p = p; a.scale = 1.f; a.shift = 0.f; a.relu6 = a.clip01 = 0; a.cols = 0;
  a.tiles_x = p.OW / kS2T;
  a.tiles_per_img = (p.OH / kS2T) * a.tiles_x;
  a.nchunks = (p.in.C + kChunk - 1) / kChunk;
  a.items = p.N * a.tiles_per_img * a.nchunks;
  a.d_chunks = make_fastdiv((uint32_t)a.nchunks); a.d_tpi = make_fastdiv((uint32_t)a.tiles_per_img); a.d_tx = make_fastdiv((uint32_t)a.tiles_x);
  CUtensorMap tmap;
  void* base = reinterpret_cast<char*>(p.in.ptr) + (size_t)p.in.coff * 2;
  if (!tma_encode_nhwc(&tmap, et == ET_BF16, base, p.in.C, p.in.W, p.in.H, p.N, p.in.pitch, kChunk, kS2Halo, kS2Halo, 1, false))
    return cudaErrorInvalidValue;
  const size_t smem = 128 + (size_t)kS2Groups * kS2StageBytes + 2 * kS2Groups * 8;
  const int grid = a.items < num_sms ? a.items : num_sms;
  return et == ET_BF16 ? launch_s2_t<__nv_bfloat16>(a, tmap, grid, smem, s) : launch_s2_t<__half>(a, tmap, grid, smem, s);
}

// final 3x3 conv 64 -> 1 (conv_block_not_sep(deconv0, 1), DMG:531) + clip (DMG:534-538) as a channel-reducing
// depthwise pass; only taken when the conv is exactly that shape
bool final_tma_supported(const ConvParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (p.Cout != 1 || p.Cin != 64 || p.ntaps != 9 || p.wtaps != 9 || p.istride != 1 || p.ostride != 1 || !p.out_f32 || p.in_f32) return false;
  for (int t = 0; t < 9; ++t)
    if (p.dy[t] != t / 3 - 1 || p.dx[t] != t % 3 - 1 || p.wrow[t] != t) return false;
  if (p.MH % kTH || p.MW % kTW || (p.in.pitch & 7) || (p.in.coff & 7) || p.res.ptr) return false;
  if (p.out.pitch != 1 || p.out.coff != 0) return false;
  return tma_encoder() != nullptr;
}

cudaError_t launch_final_tma(const ConvParams& c, float scale, float shift, int et, int num_sms, cudaStream_t s) {
  DwTmaArgs a;
  DwParams& p = a.p;
  p.in = c.in; p.out = c.out; p.N = c.N; p.OH = c.MH; p.OW = c.MW; p.stride = 1; p.rate = 1; p.pad = 1;
  p.w = c.w;   // FP32 [9*64][1] == [9][64]
  p.in_f32 = 0;
  a.scale = scale; a.shift = shift; a.relu6 = c.relu6; a.clip01 = c.clip01;
  a.cols = 0;
  return launch_common(a, et, true, num_sms, s);
}

}  // namespace emd
