// emd_fused.cu -- second-generation tcgen05 / TMEM convolution kernel for sm_100a (16-bit operands, FP32
// accumulate), for layers whose output grid tiles into 8 x 16 pixel blocks.
//
//   D[m, co] = sum_{tap t} sum_{ci}  A_t[m, ci] * W[t][ci][co]        (ConvParams, emd_kernels.h)
//
// Two sources for the A operand:
//   taps mode  A_t = the input shifted by tap t: one 4-D TMA box (64 ch x 16 x 8 px, 128B swizzle, zero fill =
//              TF SAME padding) per (tap, 64-channel chunk).  1x1 convs (DMG:299, 353, 365, 503...), dilated
//              3x3 ASPP branches (DMG:306-327), sub-pixel phases of deconv_block (DMG:278-289).
//   dw mode    A = depthwise 3x3 of the input (the DepthwiseConv2dNative half of slim.separable_convolution2d,
//              DMG:253-273), computed on the fly: TMA brings the 10 x 18 pixel halo of a 64-channel chunk, 8 math
//              warps slide the 3x3 window in registers (FP32, packed FFMA2) and write the 16-bit result
//              straight into the swizzled A stage -- the depthwise intermediate never goes to HBM.
//
// Warp roles: 0-3 epilogue, 4 MMA issuer, 5 TMA producer, 6-13 depthwise math (dw mode only).
// Epilogue: tcgen05.ld (one pixel row per thread) -> folded BN scale/shift (FFMA2) -> ReLU6 -> (+ residual,
// which a TMA load has already put into the output staging slab) -> 16-bit pack -> swizzled smem slab of
// 128 px x 64 ch -> one TMA store per slab (clipped at the view's channel count; a channel slice of a concat
// buffer or one sub-pixel phase of a transposed conv are just different tensor maps).
#include "emd_kernels.h"
#include "emd_tma.h"

#include <cstring>

namespace emd {

namespace {

using namespace ptx;

constexpr int kBM = 128, kBK = 64, kTH = 8, kTW = 16;
constexpr int kAStageBytes = kBM * kBK * 2;                 // 16 KB
constexpr int kHaloH = kTH + 2, kHaloW = kTW + 2;
constexpr int kHaloBytes = kHaloH * kHaloW * kBK * 2;       // 23040
constexpr int kSlabBytes = kBM * 64 * 2;                    // 128 pixels x 64 channels
constexpr int kMaxRing = 3;
constexpr int kMaxC = 768;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kEpiThreads = 128;
constexpr int kBaseThreads = 192;                           // epilogue + MMA + producer
constexpr int kMathThreads = 256;
constexpr int kMaxStages = 8;

struct FusedArgs {
  ConvParams p;
  NTiling nt;
  int dw_mode;
  int SA, SB, SH, ring;      // taps mode: SA stages of (A + B), SB == SA, SH == 0
  int b_stage_bytes, nchunks, w_kblocks, m_tiles, tiles_x, tiles_per_img;
  int tmem_cols, acc_stride;
  int has_res;
  const float* dw_w;         // [9][Cin] FP32, tap-major (dw mode)
};

template <typename T> struct Cv;
template <> struct Cv<__nv_bfloat16> {
  static constexpr uint32_t kFmt = 1;  // UMMA F16F32Format::BF16
  static __device__ __forceinline__ float2 up(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  static __device__ __forceinline__ uint32_t pack_relu6(float a, float b) {
    uint32_t r, m;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    asm("min.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(r), "r"(0x40c040c0u));
    return m;
  }
};
template <> struct Cv<__half> {
  static constexpr uint32_t kFmt = 0;  // UMMA F16F32Format::F16
  static __device__ __forceinline__ float2 up(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  static __device__ __forceinline__ uint32_t pack_relu6(float a, float b) {
    uint32_t r, m;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    asm("min.f16x2 %0, %1, %2;" : "=r"(m) : "r"(r), "r"(0x46004600u));
    return m;
  }
};

__device__ __forceinline__ void tile_coords(const FusedArgs& a, int mt, int& n_img, int& y0, int& x0) {
  n_img = mt / a.tiles_per_img;
  const int rem = mt - n_img * a.tiles_per_img;
  const int by = rem / a.tiles_x;
  y0 = by * kTH;
  x0 = (rem - by * a.tiles_x) * kTW;
}

template <typename T, bool kDw>
__global__ void __launch_bounds__(kDw ? kBaseThreads + kMathThreads : kBaseThreads, 1)
fused_conv_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ CUtensorMap tmap_in,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = a.p;
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_u32);
  const int SA = a.SA, SB = a.SB, SH = a.SH, R = a.ring;
  // carve: [A stages][B stages][output staging ring][halo stages][scale][shift][barriers][tmem slot]
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + (uint32_t)SA * kAStageBytes;
  const uint32_t sO = sB + (uint32_t)SB * a.b_stage_bytes;
  const uint32_t sH = sO + (uint32_t)R * kSlabBytes;
  uint8_t* g_stage = smem + (sO - smem_base);
  uint8_t* g_halo = smem + (sH - smem_base);
  float* s_scale = reinterpret_cast<float*>(g_halo + (size_t)SH * kHaloBytes);
  float* s_shift = s_scale + kMaxC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + kMaxC);
  const uint32_t bar_afull = smem_u32(bars), bar_aempty = bar_afull + 8u * kMaxStages, bar_bfull = bar_aempty + 8u * kMaxStages,
                 bar_bempty = bar_bfull + 8u * kMaxStages, bar_hfull = bar_bempty + 8u * kMaxStages,
                 bar_hempty = bar_hfull + 8u * kMaxStages, bar_tfull = bar_hempty + 8u * kMaxStages, bar_tempty = bar_tfull + 16u,
                 bar_rfull = bar_tempty + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 * kMaxStages + 4 + kMaxRing);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int kblocks = p.ntaps * a.nchunks;

  for (int i = threadIdx.x; i < kMaxC; i += blockDim.x) {
    s_scale[i] = i < p.Cout ? p.scale[i] : 0.f;
    s_shift[i] = i < p.Cout ? p.shift[i] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_afull + 8u * s, kDw ? kMathThreads / 32 : 1);   // dw: one arrive per math warp; taps: expect_tx arrive
      mbar_init(bar_aempty + 8u * s, 1);
      mbar_init(bar_bfull + 8u * s, 1);
      mbar_init(bar_bempty + 8u * s, 1);
      mbar_init(bar_hfull + 8u * s, 1);
      mbar_init(bar_hempty + 8u * s, kMathThreads / 32);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, kEpiThreads); }
    for (int i = 0; i < kMaxRing; ++i) mbar_init(bar_rfull + 8u * i, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_in);
    prefetch_tmap(&tmap_out);
    if (a.has_res) prefetch_tmap(&tmap_res);
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const char* wbase = reinterpret_cast<const char*>(p.w16);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / a.nt.nt, ntile = tile - mt * a.nt.nt;
        const uint32_t bytes = (uint32_t)a.nt.rows[ntile] * 128u;
        // packed weights: [n_tile][weight k-block = tap*nchunks + chunk][rows x 128 B, swizzled]
        const char* tbase = wbase + (size_t)a.nt.rows_before[ntile] * a.w_kblocks * 128;
        int n_img, y0, x0;
        tile_coords(a, mt, n_img, y0, x0);
        if (kDw) {
          for (int c = 0; c < a.nchunks; ++c, ++it) {
            const int sh = it % SH, sb = it % SB;
            mbar_wait(bar_hempty + 8u * sh, (uint32_t)(((it / SH) & 1) ^ 1));
            mbar_arrive_expect_tx(bar_hfull + 8u * sh, kHaloBytes);
            tma_load_4d(sH + (uint32_t)sh * kHaloBytes, &tmap_in, c * kBK, x0 - 1, y0 - 1, n_img, bar_hfull + 8u * sh);
            mbar_wait(bar_bempty + 8u * sb, (uint32_t)(((it / SB) & 1) ^ 1));
            mbar_arrive_expect_tx(bar_bfull + 8u * sb, bytes);
            bulk_g2s(sB + (uint32_t)sb * a.b_stage_bytes, tbase + (size_t)(p.wrow[0] * a.nchunks + c) * bytes, bytes, bar_bfull + 8u * sb);
          }
        } else {
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            const int t = kb / a.nchunks, c = kb - t * a.nchunks;
            const int s = it % SA;
            mbar_wait(bar_aempty + 8u * s, (uint32_t)(((it / SA) & 1) ^ 1));
            const uint32_t bar = bar_afull + 8u * s;
            mbar_arrive_expect_tx(bar, bytes + (uint32_t)kAStageBytes);
            tma_load_4d(sA + (uint32_t)s * kAStageBytes, &tmap_in, c * kBK, x0 * p.istride + p.dx[t], y0 * p.istride + p.dy[t], n_img, bar);
            bulk_g2s(sB + (uint32_t)s * a.b_stage_bytes, tbase + (size_t)(p.wrow[t] * a.nchunks + c) * bytes, bytes, bar);
          }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int ntile = tile % a.nt.nt;
      const uint32_t n = (uint32_t)a.nt.rows[ntile];
      // instruction descriptor: D = F32 (bits 4-5), A/B format (bits 7-9 / 10-12), K-major A and B, N>>3 at 17-22, M>>4 at 24-28
      const uint32_t idesc = (1u << 4) | (Cv<T>::kFmt << 7) | (Cv<T>::kFmt << 10) | ((n >> 3) << 17) | ((kBM >> 4) << 24);
      const int acc = tcount & 1;
      mbar_wait(bar_tempty + 8u * acc, (uint32_t)(((tcount >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.acc_stride);
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int c = kb % a.nchunks;
        const int sa = it % SA, sb = kDw ? it % SB : sa;
        mbar_wait(bar_afull + 8u * sa, (uint32_t)((it / SA) & 1));
        if (kDw) mbar_wait(bar_bfull + 8u * sb, (uint32_t)((it / SB) & 1));
        tc_fence_after();
        if (lane == 0) {
          const int kvalid = min(kBK, p.Cin - c * kBK);
          const int ksteps = (kvalid + 15) >> 4;
          const uint64_t adesc = make_sdesc(sA + (uint32_t)sa * kAStageBytes);
          const uint64_t bdesc = make_sdesc(sB + (uint32_t)sb * a.b_stage_bytes);
          for (int ks = 0; ks < ksteps; ++ks)  // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
            umma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
          umma_commit(bar_aempty + 8u * sa);                         // frees the stage(s) when these MMAs retire
          if (kDw) umma_commit(bar_bempty + 8u * sb);
          if (kb == kblocks - 1) umma_commit(bar_tfull + 8u * acc);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue (warp w owns TMEM lanes 32w..32w+31 = pixel rows of the tile) =====================
    const int tid = threadIdx.x;             // 0..127 = row of the tile
    const uint32_t row_off = (uint32_t)tid * 128u;
    const int rsw = tid & 7;
    // residual prefetch cursor (thread 0): slab sequence number -> (tile, slab)
    int pf_tile = blockIdx.x, pf_slab = 0, pf_q = 0;
    auto prefetch_res = [&]() {
      if (pf_tile >= total_tiles) return;
      const int mt = pf_tile / a.nt.nt, ntile = pf_tile - mt * a.nt.nt;
      int n_img, y0, x0;
      tile_coords(a, mt, n_img, y0, x0);
      const int buf = pf_q % R;
      mbar_arrive_expect_tx(bar_rfull + 8u * buf, kSlabBytes);
      tma_load_4d(sO + (uint32_t)buf * kSlabBytes, &tmap_res, a.nt.n0[ntile] + pf_slab * 64, x0, y0, n_img, bar_rfull + 8u * buf);
      ++pf_q;
      if (++pf_slab * 64 >= a.nt.rows[ntile]) { pf_slab = 0; pf_tile += gridDim.x; }
    };
    if (a.has_res && tid == 0)
      for (int i = 0; i < R - 1; ++i) prefetch_res();
    int tcount = 0, q = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / a.nt.nt, ntile = tile - mt * a.nt.nt;
      const int n0 = a.nt.n0[ntile], n = a.nt.rows[ntile];
      int n_img, y0, x0;
      tile_coords(a, mt, n_img, y0, x0);
      const int acc = tcount & 1;
      mbar_wait(bar_tfull + 8u * acc, (uint32_t)((tcount >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * a.acc_stride);
      const int nslabs = (n + 63) >> 6;
      for (int j = 0; j < nslabs; ++j, ++q) {
        const int buf = q % R;
        uint8_t* srow = g_stage + (size_t)buf * kSlabBytes + row_off;
        if (a.has_res) mbar_wait(bar_rfull + 8u * buf, (uint32_t)((q / R) & 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = j * 64 + h * 32;        // column within the N tile
          if (c0 < n) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (j == nslabs - 1 && (h == 1 || c0 + 32 >= n)) {  // last read of this accumulator: hand it back to the MMA warp
              tc_fence_before();
              mbar_arrive(bar_tempty + 8u * acc);
            }
            const float* sc = s_scale + n0 + c0;
            const float* sh = s_shift + n0 + c0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {         // 16-byte chunk = 8 channels
              const float4 sc0 = *reinterpret_cast<const float4*>(sc + 8 * k), sc1 = *reinterpret_cast<const float4*>(sc + 8 * k + 4);
              const float4 sh0 = *reinterpret_cast<const float4*>(sh + 8 * k), sh1 = *reinterpret_cast<const float4*>(sh + 8 * k + 4);
              float2 y[4];
              y[0] = ffma2(make_float2(__uint_as_float(v[8 * k + 0]), __uint_as_float(v[8 * k + 1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
              y[1] = ffma2(make_float2(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
              y[2] = ffma2(make_float2(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
              y[3] = ffma2(make_float2(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
              uint4* slot = reinterpret_cast<uint4*>(srow + ((((h * 4 + k) ^ rsw)) << 4));
              uint4 o;
              if (a.has_res) {
                const uint4 r = *slot;
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
                uint32_t ow[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float lo = y[e].x, hi = y[e].y;
                  if (p.relu6) { lo = fminf(fmaxf(lo, 0.f), 6.f); hi = fminf(fmaxf(hi, 0.f), 6.f); }
                  const float2 rr = Cv<T>::up(rw[e]);
                  ow[e] = Cv<T>::pack(lo + rr.x, hi + rr.y);
                }
                o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              } else if (p.relu6) {
                o = make_uint4(Cv<T>::pack_relu6(y[0].x, y[0].y), Cv<T>::pack_relu6(y[1].x, y[1].y), Cv<T>::pack_relu6(y[2].x, y[2].y),
                               Cv<T>::pack_relu6(y[3].x, y[3].y));
              } else {
                o = make_uint4(Cv<T>::pack(y[0].x, y[0].y), Cv<T>::pack(y[1].x, y[1].y), Cv<T>::pack(y[2].x, y[2].y),
                               Cv<T>::pack(y[3].x, y[3].y));
              }
              *slot = o;
            }
          }
        }
        fence_proxy_async();               // make this thread's slab writes visible to the TMA engine
        if (tid == 0) {                    // the slab the NEXT iteration writes must have been read out by its last store
          if (R == 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) {
          tma_store_4d(&tmap_out, sO + (uint32_t)buf * kSlabBytes, n0 + j * 64, x0, y0, n_img);
          bulk_commit();
          if (a.has_res) {                 // refill the slab stored one iteration ago with the residual of slab q + R - 1
            bulk_wait_read<1>();
            prefetch_res();
          }
        }
      }
    }
    if (tid == 0) bulk_wait<0>();
  } else if (kDw) {
    // ===================== depthwise math warps: halo -> 3x3 window in registers -> swizzled A stage =====================
    const int tm = threadIdx.x - kBaseThreads;           // 0..255
    const int qc = tm & 7, col = (tm >> 3) & 15, half = tm >> 7;
    const uint32_t a_thread = (uint32_t)((4 * half) * kTW + col) * 128u + (uint32_t)((qc ^ (col & 7)) << 4);
    float2 w[9][4];
    int cur_c = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int ch = c * kBK + qc * 8;
        if (c != cur_c) {
          cur_c = c;
          const bool ch_ok = ch < p.Cin;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
            if (ch_ok) {
              w0 = __ldg(reinterpret_cast<const float4*>(a.dw_w + t * p.Cin + ch));
              w1 = __ldg(reinterpret_cast<const float4*>(a.dw_w + t * p.Cin + ch + 4));
            }
            w[t][0] = make_float2(w0.x, w0.y); w[t][1] = make_float2(w0.z, w0.w);
            w[t][2] = make_float2(w1.x, w1.y); w[t][3] = make_float2(w1.z, w1.w);
          }
        }
        const int sh = it % SH, sa = it % SA;
        mbar_wait(bar_hfull + 8u * sh, (uint32_t)((it / SH) & 1));
        mbar_wait(bar_aempty + 8u * sa, (uint32_t)(((it / SA) & 1) ^ 1));
        const uint8_t* hb = g_halo + (size_t)sh * kHaloBytes + qc * 16;
        uint8_t* ab = smem + (size_t)sa * kAStageBytes + a_thread;
        float2 win[3][3][4];
        auto load_row = [&](int hy, float2 (&dst)[3][4]) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint4 u = *reinterpret_cast<const uint4*>(hb + (size_t)(hy * kHaloW + col + kx) * (kBK * 2));
            dst[kx][0] = Cv<T>::up(u.x); dst[kx][1] = Cv<T>::up(u.y); dst[kx][2] = Cv<T>::up(u.z); dst[kx][3] = Cv<T>::up(u.w);
          }
        };
        load_row(4 * half + 0, win[0]);
        load_row(4 * half + 1, win[1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          load_row(4 * half + i + 2, win[(i + 2) % 3]);
          float2 acc[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[j] = ffma2(win[(i + ky) % 3][kx][j], w[ky * 3 + kx][j], acc[j]);
          const uint4 o = make_uint4(Cv<T>::pack(acc[0].x, acc[0].y), Cv<T>::pack(acc[1].x, acc[1].y), Cv<T>::pack(acc[2].x, acc[2].y),
                                     Cv<T>::pack(acc[3].x, acc[3].y));
          *reinterpret_cast<uint4*>(ab + (size_t)i * kTW * 128) = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_afull + 8u * sa);
          mbar_arrive(bar_hempty + 8u * sh);
        }
      }
    }
  }

  // teardown: everyone done with TMEM, then the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

size_t fused_smem_bytes(const FusedArgs& a) {
  return 1024 + (size_t)a.SA * kAStageBytes + (size_t)a.SB * a.b_stage_bytes + (size_t)a.ring * kSlabBytes +
         (size_t)a.SH * kHaloBytes + 2 * kMaxC * sizeof(float) + (6 * kMaxStages + 4 + kMaxRing) * 8 + 16;
}

template <typename T, bool kDw>
cudaError_t launch_t(const FusedArgs& a, const CUtensorMap& tin, const CUtensorMap& tout, const CUtensorMap& tres, int grid,
                     size_t smem, cudaStream_t s) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t r = cudaFuncSetAttribute(fused_conv_kernel<T, kDw>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (r != cudaSuccess) return r;
    attr_dev = dev;
  }
  fused_conv_kernel<T, kDw><<<grid, kDw ? kBaseThreads + kMathThreads : kBaseThreads, smem, s>>>(a, tin, tout, tres);
  return cudaGetLastError();
}

}  // namespace

static bool g_use_fused = true;
void fused_set_enabled(bool on) { g_use_fused = on; }

// dw != nullptr: the GEMM's A operand is the depthwise 3x3 (stride 1, rate 1, SAME) of p.in with weights dw [9][Cin]
bool fused_supported(const ConvParams& p, int et, const float* dw) {
  if (!g_use_fused) return false;
  if (et != ET_BF16 && et != ET_F16) return false;
  if (!p.w16 || p.in_f32 || p.out_f32 || p.clip01) return false;
  if (p.Cout < 8 || p.Cout > kMaxC || (p.Cout & 7) || (p.Cin & 7)) return false;
  if ((p.in.pitch & 7) || (p.in.coff & 7) || (p.out.pitch & 7) || (p.out.coff & 7)) return false;
  if (p.res.ptr && ((p.res.pitch & 7) || (p.res.coff & 7))) return false;
  if (p.MH % kTH || p.MW % kTW) return false;
  if (kTW * p.istride > 256) return false;
  const NTiling nt = make_ntiling(p.Cout);
  if (nt.nt > kMaxNTiles) return false;
  if (nt.nt > 1 && (nt.rows[0] & 63)) return false;  // a 64-channel output slab must not straddle two N tiles
  if (dw) {
    if (p.ntaps != 1 || p.dy[0] || p.dx[0] || p.istride != 1 || p.ostride != 1) return false;
    if (make_ntiling(p.Cout).nt != 1) return false;   // the depthwise would be recomputed per N tile
  }
  if ((long long)p.N * p.MH * p.MW >= (1ll << 31) || (long long)p.N * p.in.H * p.in.W >= (1ll << 31) ||
      (long long)p.N * p.out.H * p.out.W >= (1ll << 31))
    return false;
  return tma_encoder() != nullptr;
}

cudaError_t launch_conv_fused(const ConvParams& p, int et, const float* dw, int num_sms, cudaStream_t s) {
  FusedArgs a;
  memset(&a, 0, sizeof a);
  a.p = p;
  a.nt = make_ntiling(p.Cout);
  a.dw_mode = dw ? 1 : 0;
  a.dw_w = dw;
  a.has_res = p.res.ptr ? 1 : 0;
  a.nchunks = (p.Cin + kBK - 1) / kBK;
  a.w_kblocks = p.wtaps * a.nchunks;
  a.tiles_x = p.MW / kTW;
  a.tiles_per_img = (p.MH / kTH) * a.tiles_x;
  a.m_tiles = p.N * a.tiles_per_img;
  a.b_stage_bytes = ((a.nt.maxrows * 128) + 1023) & ~1023;
  int accs = 32;
  while (accs < a.nt.maxrows) accs <<= 1;
  a.acc_stride = accs;
  a.tmem_cols = 2 * accs;
  // stage counts from the shared-memory budget
  a.ring = kMaxRing;
  if (a.dw_mode) {
    a.SA = a.SB = a.SH = 4;
    while (fused_smem_bytes(a) > (size_t)kSmemLimit) {
      if (a.ring == 3 && a.nt.maxrows > 128) { a.ring = 2; continue; }
      if (a.SB >= a.SA && a.SB >= a.SH && a.SB > 2) { --a.SB; continue; }
      if (a.SH >= a.SA && a.SH > 2) { --a.SH; continue; }
      if (a.SA > 2) { --a.SA; continue; }
      if (a.SB > 2) { --a.SB; continue; }
      if (a.ring == 3) { a.ring = 2; continue; }
      return cudaErrorInvalidValue;
    }
  } else {
    a.SA = kMaxStages; a.SH = 0;
    a.SB = a.SA;
    while (fused_smem_bytes(a) > (size_t)kSmemLimit) {
      if (a.ring == 3 && a.SA <= 4) { a.ring = 2; continue; }
      if (a.SA > 2) { --a.SA; a.SB = a.SA; continue; }
      return cudaErrorInvalidValue;
    }
  }
  const bool bf16 = et == ET_BF16;
  CUtensorMap tin, tout, tres;
  memset(&tres, 0, sizeof tres);
  void* in_base = reinterpret_cast<char*>(p.in.ptr) + (size_t)p.in.coff * 2;
  if (a.dw_mode) {
    if (!tma_encode_nhwc(&tin, bf16, in_base, p.Cin, p.in.W, p.in.H, p.N, p.in.pitch, kBK, kHaloW, kHaloH, 1, false))
      return cudaErrorInvalidValue;
  } else {
    if (!tma_encode_nhwc(&tin, bf16, in_base, p.Cin, p.in.W, p.in.H, p.N, p.in.pitch, kBK, kTW * p.istride, kTH * p.istride,
                         p.istride, true))
      return cudaErrorInvalidValue;
  }
  {  // output view on the virtual grid: pixel (my, mx) -> out pixel (my*ostride + oy0, mx*ostride + ox0)
    const size_t sx = (size_t)p.ostride * p.out.pitch, sy = (size_t)p.ostride * p.out.W * p.out.pitch,
                 sn = (size_t)p.out.H * p.out.W * p.out.pitch;
    void* ob = reinterpret_cast<char*>(p.out.ptr) + (((size_t)p.oy0 * p.out.W + p.ox0) * p.out.pitch + p.out.coff) * 2;
    if (!tma_encode_view(&tout, bf16, ob, p.Cout, p.MW, p.MH, p.N, sx, sy, sn, 64, kTW, kTH, true)) return cudaErrorInvalidValue;
    if (a.has_res) {
      const size_t rx = (size_t)p.ostride * p.res.pitch, ry = (size_t)p.ostride * p.res.W * p.res.pitch,
                   rn = (size_t)p.res.H * p.res.W * p.res.pitch;
      void* rb = reinterpret_cast<char*>(p.res.ptr) + (((size_t)p.oy0 * p.res.W + p.ox0) * p.res.pitch + p.res.coff) * 2;
      if (!tma_encode_view(&tres, bf16, rb, p.Cout, p.MW, p.MH, p.N, rx, ry, rn, 64, kTW, kTH, true)) return cudaErrorInvalidValue;
    }
  }
  const size_t smem = fused_smem_bytes(a);
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (bf16)
    return a.dw_mode ? launch_t<__nv_bfloat16, true>(a, tin, tout, tres, grid, smem, s)
                     : launch_t<__nv_bfloat16, false>(a, tin, tout, tres, grid, smem, s);
  return a.dw_mode ? launch_t<__half, true>(a, tin, tout, tres, grid, smem, s) : launch_t<__half, false>(a, tin, tout, tres, grid, smem, s);
}

}  // namespace emd
