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
// Warp roles: 0-7 epilogue, 8 MMA issuer, 9 TMA producer, 10-25 depthwise math (dw mode only).
// Epilogue: tcgen05.ld (one pixel row per thread) -> folded BN scale/shift (FFMA2) -> ReLU6 -> (+ residual,
// which a TMA load has already put into the output staging slab) -> 16-bit pack -> swizzled smem slab of
// 128 px x 64 ch -> one TMA store per slab (clipped at the view's channel count; a channel slice of a concat
// buffer or one sub-pixel phase of a transposed conv are just different tensor maps).
#include "emd_kernels.h"
#include "emd_tma.h"

#include <cstdlib>
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
// epilogue warps: two per TMEM lane quadrant, each converting one 32-column half of every 64-column slab (with 4 warps the
// output-heavy layers were bound by the epilogue's serial chain); dw mode still fits 16 math warps at 72 registers
constexpr int epi_warps(bool) { return 8; }
constexpr int base_threads(bool dw) { return epi_warps(dw) * 32 + 64; }   // epilogue + MMA + producer
constexpr int kMathThreads = 512;                           // two groups of 8 warps, alternating chunks
constexpr int kGroupWarps = 8;
constexpr int kMaxStages = 8;

struct FusedArgs {
  ConvParams p;
  NTiling nt;
  int dw_mode;
  int dww_bytes;             // dw mode: bytes of the depthwise-weight image in shared memory (nchunks x 9 x 64 floats)
  int skip_taps;             // taps mode: per-tile skipping of taps that fall wholly into the zero padding
  int dw_cols;               // dw mode: depthwise thread mapping (1 = 2 channels x 4 columns x 4 rows, 0 = 4 channels x 1 column x 8 rows)
  int SA, SB, SH, ring;      // taps mode: SA stages of (A + B), SB == SA, SH == 0
  int b_stage_bytes, nchunks, w_kblocks, m_tiles, tiles_x, tiles_per_img;
  // tile geometry: an M tile is a box of bw x bh pixels from each of bn consecutive images (<= 128 pixels; 16 x 8 x 1 wherever the
  // map tiles that way, whole-row boxes over several images for the small maps of small crops); tiles_per_img = tiles per image GROUP
  int bw, bh, bn, a_tile_bytes;
  int tmem_cols, acc_stride;
  int has_res;
  // pair mode (taps mode, wide N tiles): 2-CTA clusters; tcgen05.mma.cta_group::2 with M = 256 (each CTA's 128 pixels)
  // and each CTA holding only HALF of the B stage -- operand bytes into each SM per MAC drop by a third
  int pair;
  int b_row0[kMaxNTiles];    // first row of each N tile's blocks in the 2-D weight tensor map
  // variants (taps mode): work items are (M tile, N tile, variant); a variant has its own tap list and its own output
  // tensor map -- the 4 sub-pixel phases of a transposed conv run as ONE launch, neighbouring CTAs working on the phases
  // of the same input tile at the same time, so the input comes from HBM once (the other phases hit L2)
  int nvar;
  int v_ntaps[4];
  int v_dy[4][9], v_dx[4][9], v_wrow[4][9];
  int b_res;                 // taps mode: every B block of the launch stays in shared memory, loaded once per CTA (block = tap*nchunks + chunk)
  FastDiv d_nt, d_tpi, d_tx, d_chunks;   // by nt.nt, tiles_per_img, tiles_x, nchunks
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
  const int grp = (int)fdiv((uint32_t)mt, a.d_tpi);
  const int rem = mt - grp * a.tiles_per_img;
  const int by = (int)fdiv((uint32_t)rem, a.d_tx);
  n_img = grp * a.bn;
  y0 = by * a.bh;
  x0 = (rem - by * a.tiles_x) * a.bw;
}

// Taps of variant `var` whose shifted input box touches the image for M tile `mt` (bit t): a dilated tap whose box lies wholly in
// the zero padding contributes nothing, so its K blocks are neither loaded nor multiplied -- at rate 18 on a 32 x 32 map that is
// 5 of the 9 taps on average (SURVEY App. B counts only the in-bounds MACs).  The centre tap is always on.
__device__ __forceinline__ uint32_t tap_mask(const FusedArgs& a, int var, int mt) {
  int n_img, y0, x0;
  tile_coords(a, mt, n_img, y0, x0);
  const int s = a.p.istride;
  uint32_t m = 0;
  for (int t = 0; t < a.v_ntaps[var]; ++t) {
    const int ys = y0 * s + a.v_dy[var][t], xs = x0 * s + a.v_dx[var][t];
    if (ys < a.p.in.H && ys + a.bh * s > 0 && xs < a.p.in.W && xs + a.bw * s > 0) m |= 1u << t;
  }
  return m;
}

// position in a ring of n pipeline stages plus the mbarrier phase parity of the current pass
struct Ring {
  int idx, n;
  uint32_t phase;
  __device__ __forceinline__ Ring(int start, int n_) : idx(start % n_), n(n_), phase((uint32_t)((start / n_) & 1)) {}
  __device__ __forceinline__ void advance(int by) {   // by <= n
    idx += by;
    if (idx >= n) { idx -= n; phase ^= 1u; }
  }
};

// one 32-column half of an output slab: registers v (FP32 accumulators of this thread's pixel row) -> scale/shift ->
// ReLU6 -> (+ residual already sitting in the slab) -> 16-bit -> swizzled slab row
template <typename T, bool kRes, bool kRelu6>
__device__ __forceinline__ void epi_half(const uint32_t (&v)[32], const float* __restrict__ sc, const float* __restrict__ sh,
                                         uint8_t* srow, int h, int rsw) {
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
    if (kRes) {
      const uint4 r = *slot;
      const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = y[e].x, hi = y[e].y;
        if (kRelu6) { lo = fminf(fmaxf(lo, 0.f), 6.f); hi = fminf(fmaxf(hi, 0.f), 6.f); }
        const float2 rr = Cv<T>::up(rw[e]);
        ow[e] = Cv<T>::pack(lo + rr.x, hi + rr.y);
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (kRelu6) {
      o = make_uint4(Cv<T>::pack_relu6(y[0].x, y[0].y), Cv<T>::pack_relu6(y[1].x, y[1].y), Cv<T>::pack_relu6(y[2].x, y[2].y),
                     Cv<T>::pack_relu6(y[3].x, y[3].y));
    } else {
      o = make_uint4(Cv<T>::pack(y[0].x, y[0].y), Cv<T>::pack(y[1].x, y[1].y), Cv<T>::pack(y[2].x, y[2].y), Cv<T>::pack(y[3].x, y[3].y));
    }
    *slot = o;
  }
}

struct OutMaps { CUtensorMap m[4]; };   // output tensor map per variant

template <typename T, bool kDw, bool kRes, bool kPair>
__global__ void __launch_bounds__(kDw ? base_threads(true) + kMathThreads : base_threads(false), 1)
fused_conv_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ CUtensorMap tmap_in,
                  const __grid_constant__ OutMaps tmaps_out, const __grid_constant__ CUtensorMap tmap_res,
                  const __grid_constant__ CUtensorMap tmap_w) {
  extern __shared__ uint8_t smem_raw[];
  griddep_launch();
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
  // dw mode, where the stage budget leaves room (a.dww_bytes > 0): the depthwise weights as a [chunk][tap][64 channels] image
  // at the very end of the carve, used by the math warps only
  float* s_dww = reinterpret_cast<float*>(tmem_slot + 4);

  // warp index through a shuffle: the compiler can then prove the role branches warp-uniform and keep the MMA issuer's
  // descriptors in uniform registers (otherwise every tcgen05.mma sits in an ELECT / 7x R2UR waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int kEpiWarps = epi_warps(kDw), kEpiThreads = kEpiWarps * 32, kBaseThreads = base_threads(kDw);
  constexpr int kMmaWarp = kEpiWarps, kProdWarp = kEpiWarps + 1;
  // pair mode: the cluster (not the CTA) walks the item list; an item covers M tiles 2m and 2m+1 (one per CTA of the pair)
  uint32_t crank = 0;
  if constexpr (kPair) crank = cluster_ctarank();
  const int total_tiles = (kPair ? a.m_tiles >> 1 : a.m_tiles) * a.nt.nt * a.nvar;
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // work item -> (M tile, N tile, variant), variant fastest
  auto split = [&](int item, int& mt, int& ntile, int& var) {
    int q = item;
    var = 0;
    if (a.nvar > 1) { q = item >> 2; var = item & 3; }     // nvar is 1 or 4
    mt = (int)fdiv((uint32_t)q, a.d_nt);
    ntile = q - mt * a.nt.nt;
    if (kPair) mt = 2 * mt + (int)crank;
  };

  for (int i = threadIdx.x; i < kMaxC; i += blockDim.x) {
    s_scale[i] = i < p.Cout ? p.scale[i] : 0.f;
    s_shift[i] = i < p.Cout ? p.shift[i] : 0.f;
  }
  if (kDw && a.dww_bytes) {     // constants of the layer, like scale / shift: loaded before the dependency wait
    for (int i = threadIdx.x; i < a.nchunks * 9 * kBK; i += blockDim.x) {
      const int ch = i & (kBK - 1), t = (i >> 6) % 9, c = i / (9 * kBK);
      s_dww[i] = (c * kBK + ch < p.Cin) ? a.dw_w[t * p.Cin + c * kBK + ch] : 0.f;
    }
  }
  // barrier set-up spread over the first threads (one pipeline stage each) instead of ~55 serial inits in thread 0: the
  // prologue is part of the fixed cost every launch pays
  if (threadIdx.x < kMaxStages) {
    const uint32_t s = threadIdx.x;
    mbar_init(bar_afull + 8u * s, kDw ? kGroupWarps : 1);   // dw: one arrive per math warp of the group; taps: expect_tx arrive
    mbar_init(bar_aempty + 8u * s, 1);
    mbar_init(bar_bfull + 8u * s, 1);
    mbar_init(bar_bempty + 8u * s, 1);
    mbar_init(bar_hfull + 8u * s, 1);
    mbar_init(bar_hempty + 8u * s, kGroupWarps);
    fence_barrier_init();
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, (kPair ? 2 : 1) * kEpiWarps); }
    for (int i = 0; i < kMaxRing; ++i) mbar_init(bar_rfull + 8u * i, 1);
    fence_barrier_init();
  } else if (threadIdx.x == 64) {
    prefetch_tmap(&tmap_in);
    for (int v = 0; v < a.nvar; ++v) prefetch_tmap(&tmaps_out.m[v]);
    if (kRes) prefetch_tmap(&tmap_res);
  }
  if (warp == kMmaWarp) {
    if constexpr (kPair) tmem_alloc_2sm(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
    else tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();      // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                               // the layer before has finished: activations may be read and written from here on

  if (warp == kProdWarp) {
    // ===================== TMA producer: the whole warp walks the (uniform) loops, one elected lane issues =====================
    {
      const char* wbase = reinterpret_cast<const char*>(p.w16);
      Ring ra(0, SA), rb(0, SB), rh(0, kDw ? SH : 1);
      if constexpr (kDw) {
        // Two independent cursors over this CTA's (tile, chunk) items: halo tiles run ahead as far as the halo ring allows,
        // weight blocks follow the MMA's consumption.  (One in-order stream would hold every halo behind a weight slot that
        // only frees when an MMA retires, i.e. limit the halo prefetch distance to the weight ring's depth.)
        int h_tile = first, h_c = 0, b_tile = first, b_c = 0;
        int h_img = 0, h_y0 = 0, h_x0 = 0;
        uint32_t b_bytes = 0;
        const char* b_src = nullptr;
        bool h_new = true, b_new = true;
        while (h_tile < total_tiles || b_tile < total_tiles) {
          bool progress = false;
          if (h_tile < total_tiles) {
            if (h_new) {
              int mt, ntile, var;
              split(h_tile, mt, ntile, var);
              tile_coords(a, mt, h_img, h_y0, h_x0);
              h_new = false;
            }
            if (__all_sync(0xffffffffu, mbar_test(bar_hempty + 8u * rh.idx, rh.phase ^ 1u))) {
              if (elect_one()) {
                mbar_arrive_expect_tx(bar_hfull + 8u * rh.idx, kHaloBytes);
                tma_load_4d(sH + (uint32_t)rh.idx * kHaloBytes, &tmap_in, h_c * kBK, h_x0 - 1, h_y0 - 1, h_img, bar_hfull + 8u * rh.idx);
              }
              __syncwarp();
              progress = true;
              rh.advance(1);
              if (++h_c == a.nchunks) { h_c = 0; h_tile += step; h_new = true; }
            }
          }
          if (b_tile < total_tiles) {
            if (b_new) {
              int mt, ntile, var;
              split(b_tile, mt, ntile, var);
              b_bytes = (uint32_t)a.nt.rows[ntile] * 128u;
              b_src = wbase + (size_t)a.nt.rows_before[ntile] * a.w_kblocks * 128 + (size_t)(a.v_wrow[0][0] * a.nchunks) * b_bytes;
              b_new = false;
            }
            if (__all_sync(0xffffffffu, mbar_test(bar_bempty + 8u * rb.idx, rb.phase ^ 1u))) {
              if (elect_one()) {
                mbar_arrive_expect_tx(bar_bfull + 8u * rb.idx, b_bytes);
                bulk_g2s(sB + (uint32_t)rb.idx * a.b_stage_bytes, b_src, b_bytes, bar_bfull + 8u * rb.idx);
              }
              __syncwarp();
              progress = true;
              b_src += b_bytes;
              rb.advance(1);
              if (++b_c == a.nchunks) { b_c = 0; b_tile += step; b_new = true; }
            }
          }
          // neither slot is free: sleep on a barrier (hardware-suspended, bounded) instead of spinning -- the producer shares its
          // scheduler with math warps that need every issue slot
          if (!progress) {
            if (h_tile < total_tiles) mbar_try_wait_ns(bar_hempty + 8u * rh.idx, rh.phase ^ 1u, 400u);
            else mbar_try_wait_ns(bar_bempty + 8u * rb.idx, rb.phase ^ 1u, 400u);
          }
        }
      }
      for (int tile = first; !kDw && tile < total_tiles; tile += step) {
        int mt, ntile, var;
        split(tile, mt, ntile, var);
        const int ntaps = a.v_ntaps[var];
        const uint32_t bytes = (uint32_t)a.nt.rows[ntile] * 128u;
        // packed weights: [n_tile][weight k-block = tap*nchunks + chunk][rows x 128 B, swizzled]
        const char* tbase = wbase + (size_t)a.nt.rows_before[ntile] * a.w_kblocks * 128;
        int n_img, y0, x0;
        tile_coords(a, mt, n_img, y0, x0);
        {
          const int xb = x0 * p.istride, yb = y0 * p.istride;
          if (a.b_res && tile == first && elect_one()) {   // first tile of this CTA: bring in every weight block, once
            mbar_arrive_expect_tx(bar_bfull, bytes * (uint32_t)(ntaps * a.nchunks));
            for (int t = 0; t < ntaps; ++t)
              for (int c = 0; c < a.nchunks; ++c)
                bulk_g2s(sB + (uint32_t)(t * a.nchunks + c) * a.b_stage_bytes, tbase + (size_t)(a.v_wrow[0][t] * a.nchunks + c) * bytes, bytes, bar_bfull);
          }
          uint32_t tmask = 0xffffffffu;            // pair mode: the union over both CTAs' tiles (the leader issues one MMA stream for both)
          if (a.skip_taps) tmask = kPair ? (tap_mask(a, var, mt & ~1) | tap_mask(a, var, mt | 1)) : tap_mask(a, var, mt);
          for (int t = 0; t < ntaps; ++t) {
            if (!((tmask >> t) & 1u)) continue;
            const char* wsrc = tbase + (size_t)(a.v_wrow[var][t] * a.nchunks) * bytes;
            const int xt = xb + a.v_dx[var][t], yt = yb + a.v_dy[var][t];
            for (int c = 0; c < a.nchunks; ++c) {
              mbar_wait(bar_aempty + 8u * ra.idx, ra.phase ^ 1u);
              const uint32_t bar = bar_afull + 8u * ra.idx;
              if (elect_one()) {
                if constexpr (kPair) {
                  // both CTAs' copies complete on the LEADER's full barrier; each CTA brings its own A tile and its half of B's rows
                  // (the weight box always has maxrows/2 rows; for a narrower last N tile the surplus rows are never read)
                  if (crank == 0) mbar_arrive_expect_tx(bar, 2u * ((uint32_t)(a.nt.maxrows >> 1) * 128u + (uint32_t)a.a_tile_bytes));
                  tma_load_4d_2sm(sA + (uint32_t)ra.idx * kAStageBytes, &tmap_in, c * kBK, xt, yt, n_img, bar);
                  const int half = a.nt.rows[ntile] >> 1;
                  tma_load_2d_2sm(sB + (uint32_t)ra.idx * a.b_stage_bytes, &tmap_w, 0,
                                  (a.b_row0[ntile] + (a.v_wrow[var][t] * a.nchunks + c) * a.nt.rows[ntile] + (int)crank * half) >> 2, bar);
                } else {
                  mbar_arrive_expect_tx(bar, (a.b_res ? 0u : bytes) + (uint32_t)a.a_tile_bytes);
                  tma_load_4d(sA + (uint32_t)ra.idx * kAStageBytes, &tmap_in, c * kBK, xt, yt, n_img, bar);
                  if (!a.b_res) bulk_g2s(sB + (uint32_t)ra.idx * a.b_stage_bytes, wsrc, bytes, bar);
                }
              }
              __syncwarp();
              wsrc += bytes;
              ra.advance(1);
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    Ring ra(0, SA), rb(0, SB);
    int tcount = 0;
    const bool issuer = !kPair || crank == 0;   // pair mode: the leader CTA issues the M = 256 MMAs for both
    const int last_ksteps = (p.Cin - (a.nchunks - 1) * kBK + 15) >> 4;   // K=16 steps of the last (possibly partial) chunk
    const uint32_t a_lo0 = sdesc_lo(sA), b_lo0 = sdesc_lo(sB), b_step = (uint32_t)a.b_stage_bytes >> 4;   // descriptor low words of stage 0
    for (int tile = first; tile < total_tiles; tile += step, ++tcount) {
      int mt_mma, ntile, var;
      split(tile, mt_mma, ntile, var);
      int kblocks = a.v_ntaps[var] * a.nchunks;
      if (!kDw && a.skip_taps) {
        const uint32_t tmask = kPair ? (tap_mask(a, var, mt_mma & ~1) | tap_mask(a, var, mt_mma | 1)) : tap_mask(a, var, mt_mma);
        kblocks = __popc(tmask) * a.nchunks;
      }
      const uint32_t n = (uint32_t)a.nt.rows[ntile];
      // instruction descriptor: D = F32 (bits 4-5), A/B format (bits 7-9 / 10-12), K-major A and B, N>>3 at 17-22, M>>4 at 24-28
      const uint32_t idesc = (1u << 4) | (Cv<T>::kFmt << 7) | (Cv<T>::kFmt << 10) | ((n >> 3) << 17) |
                             (((kPair ? 2 * kBM : kBM) >> 4) << 24);
      const int acc = tcount & 1;
      if (!issuer) continue;
      mbar_wait(bar_tempty + 8u * acc, (uint32_t)(((tcount >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.acc_stride);
      if (!kDw && a.b_res && tcount == 0) mbar_wait(bar_bfull, 0u);   // resident weights have landed
      int c = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int sa = ra.idx, sb = kDw ? rb.idx : (a.b_res ? kb : sa);
        mbar_wait(bar_afull + 8u * sa, ra.phase);
        if (kDw) mbar_wait(bar_bfull + 8u * sb, rb.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)sa * (kAStageBytes >> 4), b_lo = b_lo0 + (uint32_t)sb * b_step;
          const uint32_t accum = kb != 0 ? 1u : 0u;
          if (c != a.nchunks - 1 || last_ksteps == 4) umma_kblock<kPair, 4>(d_tmem, a_lo, b_lo, idesc, accum);
          else if (last_ksteps == 3) umma_kblock<kPair, 3>(d_tmem, a_lo, b_lo, idesc, accum);
          else if (last_ksteps == 2) umma_kblock<kPair, 2>(d_tmem, a_lo, b_lo, idesc, accum);
          else umma_kblock<kPair, 1>(d_tmem, a_lo, b_lo, idesc, accum);
          if constexpr (kPair) {
            umma_commit_2sm(bar_aempty + 8u * sa, (uint16_t)3);                         // frees the stage in BOTH CTAs
            if (kb == kblocks - 1) umma_commit_2sm(bar_tfull + 8u * acc, (uint16_t)3);  // both epilogues
          } else {
            umma_commit(bar_aempty + 8u * sa);                         // frees the stage(s) when these MMAs retire
            if (kDw) umma_commit(bar_bempty + 8u * sb);
            if (kb == kblocks - 1) umma_commit(bar_tfull + 8u * acc);  // accumulator complete
          }
        }
        __syncwarp();
        ra.advance(1);
        if (kDw) rb.advance(1);
        if (++c == a.nchunks) c = 0;
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (warp w reads TMEM lanes 32(w%4).. = pixel rows of the tile; with 8 warps, warps w and
    // w+4 take the two 32-column halves of every slab) =====================
    const int tid = threadIdx.x;
    const int row = (warp & 3) * 32 + lane;  // row of the tile
    const int my_half = warp >> 2;           // 8-warp mode: the half of each slab this warp converts
    const uint32_t row_off = (uint32_t)row * 128u;
    const int rsw = row & 7;
    // residual prefetch cursor (thread 0): slab sequence number -> (tile, slab)
    int pf_tile = first, pf_slab = 0, pf_buf = 0;
    auto prefetch_res = [&]() {
      if (pf_tile < total_tiles) {
        int mt, ntile, var;
        split(pf_tile, mt, ntile, var);
        int n_img, y0, x0;
        tile_coords(a, mt, n_img, y0, x0);
        mbar_arrive_expect_tx(bar_rfull + 8u * pf_buf, (uint32_t)a.a_tile_bytes);   // a 64-channel slab of the tile's pixels
        tma_load_4d(sO + (uint32_t)pf_buf * kSlabBytes, &tmap_res, a.nt.n0[ntile] + pf_slab * 64, x0, y0, n_img, bar_rfull + 8u * pf_buf);
        if (++pf_slab * 64 >= a.nt.rows[ntile]) { pf_slab = 0; pf_tile += step; }
      }
      if (++pf_buf == R) pf_buf = 0;
    };
    if (kRes && tid == 0)
      for (int i = 0; i < R - 1; ++i) prefetch_res();
    int tcount = 0;
    Ring rq(0, R);                           // output staging slab
    for (int tile = first; tile < total_tiles; tile += step, ++tcount) {
      int mt, ntile, var;
      split(tile, mt, ntile, var);
      const int n0 = a.nt.n0[ntile], n = a.nt.rows[ntile];
      int n_img, y0, x0;
      tile_coords(a, mt, n_img, y0, x0);
      const int acc = tcount & 1;
      mbar_wait(bar_tfull + 8u * acc, (uint32_t)((tcount >> 1) & 1));   // (one polling warp + a named barrier for the rest: measured, no difference)
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(acc * a.acc_stride);
      const int nslabs = (n + 63) >> 6;
      // this warp converts the 32-column half `my_half` of every 64-column slab; the accumulator read of slab j+1 is in
      // flight while slab j is converted and stored (the epilogue, not the MMA, paces the output-heavy layers)
      constexpr bool kMath = kDw;                // the kernel with depthwise math warps runs at 72 registers per thread: no room for a second buffer
      constexpr int kVB = kMath ? 1 : 2;
      uint32_t v[kVB][32];
      const int cbase = my_half * 32;
      if (!kMath && cbase < n) tmem_ld32(taddr + (uint32_t)cbase, v[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= nslabs) break;
        const int buf = rq.idx;
        uint8_t* srow = g_stage + (size_t)buf * kSlabBytes + row_off;
        const int c0 = j * 64 + cbase;           // column within the N tile
        if (kMath && c0 < n) tmem_ld32(taddr + (uint32_t)c0, v[0]);
        tmem_ld_wait();
        if (!kMath && j + 1 < nslabs && c0 + 64 < n) tmem_ld32(taddr + (uint32_t)(c0 + 64), v[(j + 1) % kVB]);
        if (j == nslabs - 1) {                   // this warp's last read of the accumulator is done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (kPair) mbar_arrive_leader(bar_tempty + 8u * acc); else mbar_arrive(bar_tempty + 8u * acc); }
        }
        if (kRes) mbar_wait(bar_rfull + 8u * buf, rq.phase);
        if (c0 < n) {
          if (p.relu6) epi_half<T, kRes, true>(v[j % kVB], s_scale + n0 + c0, s_shift + n0 + c0, srow, my_half, rsw);
          else epi_half<T, kRes, false>(v[j % kVB], s_scale + n0 + c0, s_shift + n0 + c0, srow, my_half, rsw);
        }
        fence_proxy_async();               // make this thread's slab writes visible to the TMA engine
        if (tid == 0) {                    // the slab the NEXT iteration writes must have been read out by its last store
          if (R == 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) {
          tma_store_4d(&tmaps_out.m[var], sO + (uint32_t)buf * kSlabBytes, n0 + j * 64, x0, y0, n_img);
          bulk_commit();
          if (kRes) {                      // refill the slab stored one iteration ago with the residual of the slab R-1 ahead
            bulk_wait_read<1>();
            prefetch_res();
          }
        }
        rq.advance(1);
      }
    }
    if (tid == 0) bulk_wait<0>();
  } else if (kDw) {
    // ===================== depthwise math warps: halo -> 3x3 window in registers -> swizzled A stage =====================
    // Two groups of 8 warps take alternate (tile, chunk) items.  Thread = 4 channels x 1 pixel column x 8 rows:
    // 16 lanes read one 128-byte halo pixel per LDS.64, the window slides down the column in registers.
    const int tm = threadIdx.x - kBaseThreads;           // 0..511
    const int grp = tm >> 8, tg = tm & 255;
    const int items = ((total_tiles - first + step - 1) / step) * a.nchunks;   // this CTA's (tile, chunk) items
    Ring rh(grp, SH), ra(grp, SA);
    int c = grp % a.nchunks;
    int cur_c = -1;
    if (a.dw_cols) {
      // Thread = 2 channels x 4 pixel columns x 4 rows.  A warp's 32 lanes are the 32 channel pairs of one halo pixel, so every
      // LDS.32 / STS.32 is one conflict-free 128-byte wavefront, and a loaded + unpacked halo value feeds up to 9 outputs of this
      // thread: 36 loads and 72 unpacks per 32 outputs, against 30 two-wavefront LDS.64 and 120 unpacks with one column per
      // thread.  Row-accumulate form: halo row j adds into output rows j-2 .. j, so at most three output rows are live.
      const int cp = tg & 31, sub = tg >> 5;
      const int x0 = (sub & 3) * 4, y0 = (sub >> 2) * 4;
      // three per-thread constants carry all the addressing (the compiler otherwise re-derived five offsets from threadIdx for
      // every item): the halo offset, the A-stage offset of output column 0 and the swizzle term -- column i of a row sits at
      // a_base + 128 i + (kx ^ 16 i), i.e. one LOP3 + one add per column and per item, the row offsets are immediates
      uint32_t h_off = (uint32_t)((y0 * kHaloW + x0) * (kBK * 2) + cp * 4);
      uint32_t a_base = (uint32_t)(y0 * kTW * 128 + x0 * 128 + (cp & 3) * 4) | ((uint32_t)(((cp >> 2) ^ (x0 & 4)) << 4) << 20);   // kx in bits 24..26
      asm volatile("" : "+r"(h_off), "+r"(a_base));       // opaque: keep them in registers instead of re-deriving them from threadIdx per item
      const uint32_t kx = a_base >> 20;
      const uint32_t w_off = smem_u32(s_dww) + (uint32_t)cp * 8u;
      float2 w[9];
      for (int it = grp; it < items; it += 2) {
        if (c != cur_c) {                              // this chunk's 9 x 2 weights: from the [chunk][tap][64] image in shared memory, or global
          cur_c = c;
          if (a.dww_bytes) {
            const uint32_t wp = w_off + (uint32_t)c * (9u * kBK * 4u);
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = lds_f2(wp + (uint32_t)t * (kBK * 4));
          } else {
            const int ch = c * kBK + cp * 2;
            const bool ch_ok = ch < p.Cin;
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = ch_ok ? __ldg(reinterpret_cast<const float2*>(a.dw_w + t * p.Cin + ch)) : make_float2(0.f, 0.f);
          }
        }
        mbar_wait(bar_hfull + 8u * rh.idx, rh.phase);
        mbar_wait(bar_aempty + 8u * ra.idx, ra.phase ^ 1u);
        const uint32_t hb = sH + (uint32_t)rh.idx * kHaloBytes + h_off;
        const uint32_t ab0 = sA + (uint32_t)ra.idx * kAStageBytes + (a_base & 0xfffffu);
        uint32_t ab[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ab[i] = ab0 + (uint32_t)i * 128u + (kx ^ ((uint32_t)i << 4));
        float2 acc[3][4];
#pragma unroll
        for (int j = 0; j < 6; ++j) {                  // halo rows y0 - 1 + j
          float2 x[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) x[i] = Cv<T>::up(lds_u32(hb + (uint32_t)((j * kHaloW + i) * (kBK * 2))));
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {             // this halo row is tap row ky of output row j - ky
            const int r = j - ky;
            if (r < 0 || r > 3) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2& d = acc[r % 3][i];
              d = ky == 0 ? fmul2(x[i], w[0]) : ffma2(x[i], w[ky * 3], d);
              d = ffma2(x[i + 1], w[ky * 3 + 1], d);
              d = ffma2(x[i + 2], w[ky * 3 + 2], d);
            }
          }
          if (j >= 2) {                                // output row j - 2 is complete
            const int r = j - 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) sts_u32(ab[i] + (uint32_t)(r * kTW * 128), Cv<T>::pack(acc[r % 3][i].x, acc[r % 3][i].y));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_afull + 8u * ra.idx);
          mbar_arrive(bar_hempty + 8u * rh.idx);
        }
        rh.advance(2); ra.advance(2);
        c += 2;
        if (c >= a.nchunks) c -= a.nchunks;
        if (c >= a.nchunks) c -= a.nchunks;
      }
    } else {
    const int cq = tg & 15, col = tg >> 4;
    const uint32_t a_thread = (uint32_t)col * 128u + (uint32_t)((((cq >> 1) ^ (col & 7)) << 4) + (cq & 1) * 8);
    const uint32_t h_thread = (uint32_t)col * (kBK * 2) + (uint32_t)cq * 8;
    float2 w[9][2];
    for (int it = grp; it < items; it += 2) {
      if (c != cur_c) {
        cur_c = c;
        const int ch = c * kBK + cq * 4;
        const bool ch_ok = ch < p.Cin;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch_ok) w0 = __ldg(reinterpret_cast<const float4*>(a.dw_w + t * p.Cin + ch));
          w[t][0] = make_float2(w0.x, w0.y); w[t][1] = make_float2(w0.z, w0.w);
        }
      }
      mbar_wait(bar_hfull + 8u * rh.idx, rh.phase);
      mbar_wait(bar_aempty + 8u * ra.idx, ra.phase ^ 1u);
      const uint8_t* hb = g_halo + (size_t)rh.idx * kHaloBytes + h_thread;
      uint8_t* ab = smem + (size_t)ra.idx * kAStageBytes + a_thread;
      float2 win[3][3][2];
      auto load_row = [&](int hy, float2 (&dst)[3][2]) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint2 u = *reinterpret_cast<const uint2*>(hb + (hy * kHaloW + kx) * (kBK * 2));
          dst[kx][0] = Cv<T>::up(u.x); dst[kx][1] = Cv<T>::up(u.y);
        }
      };
      load_row(0, win[0]);
      load_row(1, win[1]);
#pragma unroll
      for (int i = 0; i < kTH; ++i) {
        load_row(i + 2, win[(i + 2) % 3]);
        float2 acc0 = fmul2(win[i % 3][0][0], w[0][0]), acc1 = fmul2(win[i % 3][0][1], w[0][1]);
#pragma unroll
        for (int t = 1; t < 9; ++t) {
          acc0 = ffma2(win[(i + t / 3) % 3][t % 3][0], w[t][0], acc0);
          acc1 = ffma2(win[(i + t / 3) % 3][t % 3][1], w[t][1], acc1);
        }
        *reinterpret_cast<uint2*>(ab + i * kTW * 128) = make_uint2(Cv<T>::pack(acc0.x, acc0.y), Cv<T>::pack(acc1.x, acc1.y));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_afull + 8u * ra.idx);
        mbar_arrive(bar_hempty + 8u * rh.idx);
      }
      rh.advance(2); ra.advance(2);
      c += 2;
      if (c >= a.nchunks) c -= a.nchunks;
      if (c >= a.nchunks) c -= a.nchunks;
    }
    }
  }

  // teardown: everyone done with TMEM, then the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();      // neither CTA retires while the other may still signal its barriers or read its smem
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_2sm(tmem_base, (uint32_t)a.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

size_t fused_smem_bytes(const FusedArgs& a) {
  return 1024 + (size_t)a.SA * kAStageBytes + (size_t)a.SB * a.b_stage_bytes + (size_t)a.ring * kSlabBytes +
         (size_t)a.SH * kHaloBytes + 2 * kMaxC * sizeof(float) + (size_t)a.dww_bytes + (6 * kMaxStages + 4 + kMaxRing) * 8 + 16;
}

template <typename T, bool kDw, bool kRes, bool kPair>
cudaError_t launch_t(const FusedArgs& a, const CUtensorMap& tin, const OutMaps& tout, const CUtensorMap& tres, const CUtensorMap& tw,
                     int grid, size_t smem, cudaStream_t s) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t r = cudaFuncSetAttribute(fused_conv_kernel<T, kDw, kRes, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (r != cudaSuccess) return r;
    attr_dev = dev;
  }
  const int block = kDw ? base_threads(true) + kMathThreads : base_threads(false);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kPair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, fused_conv_kernel<T, kDw, kRes, kPair>, a, tin, tout, tres, tw);
}

}  // namespace

// Tile geometry of an output grid of N maps of MH x MW pixels: 16 x 8 pixel blocks of one image where the map tiles that way
// (the shape the depthwise producer and the fused trunk work on); else whole rows -- bh rows from each of bn consecutive images,
// chosen to fill the 128-row M tile best (6 x 6 maps of 96 x 96 crops: 3 rows x 7 images = 126 pixels; 24 x 24: 1 x 5 = 120).
static bool pick_tile(int MH, int MW, int N, int* bw, int* bh, int* bn) {
  if (MH % kTH == 0 && MW % kTW == 0) { *bw = kTW; *bh = kTH; *bn = 1; return true; }
  if (MW > kBM || MW < 1) return false;
  double best = 0;
  for (int h = 1; h <= MH && h * MW <= kBM; ++h)
    for (int n = 1; n * h * MW <= kBM && n <= 16 && n <= (N > 1 ? N : 1); ++n) {
      const double tiles = (double)((N + n - 1) / n) * ((MH + h - 1) / h);
      const double eff = (double)N * MH * MW / (tiles * kBM);
      if (eff > best + 1e-9) { best = eff; *bw = MW; *bh = h; *bn = n; }
    }
  return best > 0;
}

// dw != nullptr: the GEMM's A operand is the depthwise 3x3 (stride 1, rate 1, SAME) of p.in with weights dw [9][Cin]
bool fused_supported(const ConvParams& p, int et, const float* dw) {
  if (!tuning().fused) return false;
  if (et != ET_BF16 && et != ET_F16) return false;
  if (!p.w16 || p.in_f32 || p.out_f32 || p.clip01) return false;
  if (p.Cout < 8 || p.Cout > kMaxC || (p.Cout & 7) || (p.Cin & 7)) return false;
  if ((p.in.pitch & 7) || (p.in.coff & 7) || (p.out.pitch & 7) || (p.out.coff & 7)) return false;
  if (p.res.ptr && ((p.res.pitch & 7) || (p.res.coff & 7))) return false;
  int bw, bh, bn;
  if (!pick_tile(p.MH, p.MW, p.N, &bw, &bh, &bn)) return false;
  if (bw * p.istride > 256 || bh * p.istride > 256) return false;
  if (dw && (bw != kTW || bh != kTH || bn != 1)) return false;   // the depthwise producer works on 8 x 16 pixel tiles
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

// The nvar (1 or 4) convolutions differ only in their tap lists and output offsets (the sub-pixel phases of one transposed conv)
bool fused_multi_supported(const ConvParams* ps, int nvar, int et) {
  if (nvar != 4) return false;
  for (int v = 0; v < nvar; ++v) {
    const ConvParams& q = ps[v];
    if (!fused_supported(q, et, nullptr) || q.res.ptr || q.istride != 1) return false;
    if (q.in.ptr != ps[0].in.ptr || q.out.ptr != ps[0].out.ptr || q.w16 != ps[0].w16 || q.Cin != ps[0].Cin || q.Cout != ps[0].Cout ||
        q.MH != ps[0].MH || q.MW != ps[0].MW || q.N != ps[0].N || q.ostride != ps[0].ostride || q.relu6 != ps[0].relu6)
      return false;
  }
  return true;
}

static cudaError_t launch_impl(const ConvParams* ps, int nvar, int et, const float* dw, int num_sms, cudaStream_t s);

cudaError_t launch_conv_fused(const ConvParams& p, int et, const float* dw, int num_sms, cudaStream_t s) {
  return launch_impl(&p, 1, et, dw, num_sms, s);
}
cudaError_t launch_conv_fused_multi(const ConvParams* ps, int nvar, int et, int num_sms, cudaStream_t s) {
  return launch_impl(ps, nvar, et, nullptr, num_sms, s);
}

// CTA pairs (cta_group::2) for wide N tiles -- the GEMM is bound by operand bytes into the SM, and a pair holds half a B stage
// per CTA -- when there are enough pair items to fill the GPU
static bool want_pair(int m_tiles, const NTiling& nt, int nvar, int num_sms) {
  const Tuning& tn = tuning();
  const int items_pair = (m_tiles >> 1) * nt.nt * nvar;
  const int min_items = tn.pair_min_items >= 0 ? tn.pair_min_items : num_sms;   // fewer pair items than SMs: single CTAs fill the GPU better
  if (!(tn.pair && nt.maxrows >= tn.pair_min_rows && !(m_tiles & 1) && items_pair >= 1 && items_pair >= min_items)) return false;
  for (int i = 0; i < nt.nt; ++i)
    if (nt.rows[i] & 31) return false;                   // N and N/2 stay multiples of 16
  return true;
}

static cudaError_t launch_impl(const ConvParams* ps, int nvar, int et, const float* dw, int num_sms, cudaStream_t s) {
  const ConvParams& p = ps[0];
  FusedArgs a;
  memset(&a, 0, sizeof a);
  a.p = p;
  a.nvar = nvar;
  for (int v = 0; v < nvar; ++v) {
    a.v_ntaps[v] = ps[v].ntaps;
    for (int t = 0; t < ps[v].ntaps && t < 9; ++t) { a.v_dy[v][t] = ps[v].dy[t]; a.v_dx[v][t] = ps[v].dx[t]; a.v_wrow[v][t] = ps[v].wrow[t]; }
  }
  a.nt = make_ntiling(p.Cout);
  a.dw_mode = dw ? 1 : 0;
  a.dw_cols = tuning().dw_cols;
  a.dw_w = dw;
  a.has_res = p.res.ptr ? 1 : 0;
  a.nchunks = (p.Cin + kBK - 1) / kBK;
  a.dww_bytes = 0;
  a.w_kblocks = p.wtaps * a.nchunks;
  if (!pick_tile(p.MH, p.MW, p.N, &a.bw, &a.bh, &a.bn)) return cudaErrorInvalidValue;
  a.a_tile_bytes = a.bw * a.bh * a.bn * 128;
  a.tiles_x = p.MW / a.bw;
  a.tiles_per_img = ((p.MH + a.bh - 1) / a.bh) * a.tiles_x;
  a.m_tiles = ((p.N + a.bn - 1) / a.bn) * a.tiles_per_img;
  a.b_stage_bytes = ((a.nt.maxrows * 128) + 1023) & ~1023;
  a.d_nt = make_fastdiv((uint32_t)a.nt.nt); a.d_tpi = make_fastdiv((uint32_t)a.tiles_per_img);
  a.d_tx = make_fastdiv((uint32_t)a.tiles_x); a.d_chunks = make_fastdiv((uint32_t)a.nchunks);
  int accs = 32;
  while (accs < a.nt.maxrows) accs <<= 1;
  a.acc_stride = accs;
  a.tmem_cols = 2 * accs;
  // stage counts from the shared-memory budget
  a.ring = kMaxRing;
  if (a.dw_mode) {
    // Each of the two math groups holds one halo and one A stage at a time, so A needs one spare stage and the halo ring
    // wants everything that is left: the halo tiles are the HBM stream, and their prefetch distance is SH - 2 items.
    // The output ring only needs its third slab where a residual is prefetched into it.
    a.SA = 3; a.SB = 2; a.SH = 6;      // two weight stages: a third buys nothing, its bytes are better spent on halo stages (A/B: -1.3 % step time)
    a.ring = a.has_res ? kMaxRing : 2;
    const Tuning& tn = tuning();
    if (tn.dw_sa) a.SA = tn.dw_sa;
    if (tn.dw_sb) a.SB = tn.dw_sb;
    if (tn.dw_sh) a.SH = tn.dw_sh;
    if (tn.dw_ring) a.ring = tn.dw_ring;
    while (fused_smem_bytes(a) > (size_t)kSmemLimit) {
      if (a.SH > 4) { --a.SH; continue; }
      if (a.ring == 3 && a.nt.maxrows > 128) { a.ring = 2; continue; }
      if (a.SB > 2) { --a.SB; continue; }
      if (a.SH > 2) { --a.SH; continue; }
      if (a.SA > 2) { --a.SA; continue; }
      if (a.ring == 3) { a.ring = 2; continue; }
      return cudaErrorInvalidValue;
    }
    // the depthwise-weight image in shared memory only where it costs no pipeline stage (layers with several chunks per group
    // switch chunks every item: 9 LDS.64 with immediate offsets instead of 9 global loads and their pointer arithmetic)
    // (measured, deconv1_0, 384 -> 128: a third weight stage given up for the image is a gain, 0.66 -> 0.60 ms; a halo stage is
    // a loss, deconv1_1 0.30 -> 0.40 ms)
    a.dww_bytes = (a.dw_cols && a.nchunks > 2) ? a.nchunks * 9 * kBK * (int)sizeof(float) : 0;
    if (a.dww_bytes && fused_smem_bytes(a) > (size_t)kSmemLimit) {
      if (a.SB > 2) { --a.SB; if (fused_smem_bytes(a) > (size_t)kSmemLimit) { ++a.SB; a.dww_bytes = 0; } }
      else a.dww_bytes = 0;
    }
  } else {
    a.pair = want_pair(a.m_tiles, a.nt, nvar, num_sms) ? 1 : 0;
    if (a.pair) a.b_stage_bytes = (((a.nt.maxrows >> 1) * 128) + 1023) & ~1023;
    a.SA = kMaxStages; a.SH = 0;
    a.SB = a.SA;
    while (fused_smem_bytes(a) > (size_t)kSmemLimit) {
      if (a.ring == 3 && a.SA <= 4) { a.ring = 2; continue; }
      if (a.SA > 2) { --a.SA; a.SB = a.SA; continue; }
      return cudaErrorInvalidValue;
    }
    // resident weights: when every B block of the launch fits next to >= 3 A stages, load them once per CTA instead of
    // once per tile -- the per-tile traffic into the SM drops to the A tiles alone
    const int nblocks = p.ntaps * a.nchunks;
    if (a.nt.nt == 1 && nvar == 1 && !a.pair && nblocks <= 24) {
      FusedArgs b = a;
      b.b_res = 1; b.SB = nblocks; b.SA = kMaxStages; b.ring = kMaxRing;
      while (fused_smem_bytes(b) > (size_t)kSmemLimit && (b.SA > 3 || b.ring > 2)) {
        if (b.SA > 3) --b.SA; else --b.ring;
      }
      if (fused_smem_bytes(b) <= (size_t)kSmemLimit) a = b;
    }
  }
  // per-tile tap skipping only where a tap CAN fall wholly outside the image for some tile (a shift of at least one tile extent:
  // the rate-12 / rate-18 ASPP branches); elsewhere the mask arithmetic is pure overhead for the single issuing warps (measured
  // on the transposed convs: 0.63 -> 0.83 ms)
  a.skip_taps = 0;
  if (!a.dw_mode && !a.b_res && tuning().skip_taps)
    for (int v = 0; v < nvar; ++v)
      for (int t = 0; t < ps[v].ntaps; ++t)
        if (abs(ps[v].dy[t]) >= a.bh * p.istride || abs(ps[v].dx[t]) >= a.bw * p.istride) a.skip_taps = 1;
  const bool bf16 = et == ET_BF16;
  CUtensorMap tin, tres, tw;
  OutMaps tout;
  memset(&tw, 0, sizeof tw);
  if (a.pair) {
    size_t rows_total = 0;
    for (int i = 0; i < a.nt.nt; ++i) { a.b_row0[i] = (int)rows_total; rows_total += (size_t)a.nt.rows[i] * a.w_kblocks; }
    if (!tma_encode_linear512(&tw, bf16, p.w16, rows_total * 128, (a.nt.maxrows >> 1) * 128)) return cudaErrorInvalidValue;
  }
  memset(&tres, 0, sizeof tres);
  memset(&tout, 0, sizeof tout);
  void* in_base = reinterpret_cast<char*>(p.in.ptr) + (size_t)p.in.coff * 2;
  if (a.dw_mode) {
    if (!tma_encode_nhwc(&tin, bf16, in_base, p.Cin, p.in.W, p.in.H, p.N, p.in.pitch, kBK, kHaloW, kHaloH, 1, false))
      return cudaErrorInvalidValue;
  } else {
    if (!tma_encode_nhwc(&tin, bf16, in_base, p.Cin, p.in.W, p.in.H, p.N, p.in.pitch, kBK, a.bw * p.istride, a.bh * p.istride,
                         p.istride, true, a.bn))
      return cudaErrorInvalidValue;
  }
  for (int v = 0; v < nvar; ++v) {  // output view on the virtual grid: pixel (my, mx) -> out pixel (my*ostride + oy0, mx*ostride + ox0)
    const ConvParams& q = ps[v];
    const size_t sx = (size_t)q.ostride * q.out.pitch, sy = (size_t)q.ostride * q.out.W * q.out.pitch,
                 sn = (size_t)q.out.H * q.out.W * q.out.pitch;
    void* ob = reinterpret_cast<char*>(q.out.ptr) + (((size_t)q.oy0 * q.out.W + q.ox0) * q.out.pitch + q.out.coff) * 2;
    if (!tma_encode_view(&tout.m[v], bf16, ob, q.Cout, q.MW, q.MH, q.N, sx, sy, sn, 64, a.bw, a.bh, true, a.bn)) return cudaErrorInvalidValue;
  }
  if (a.has_res) {
    const size_t rx = (size_t)p.ostride * p.res.pitch, ry = (size_t)p.ostride * p.res.W * p.res.pitch,
                 rn = (size_t)p.res.H * p.res.W * p.res.pitch;
    void* rb = reinterpret_cast<char*>(p.res.ptr) + (((size_t)p.oy0 * p.res.W + p.ox0) * p.res.pitch + p.res.coff) * 2;
    if (!tma_encode_view(&tres, bf16, rb, p.Cout, p.MW, p.MH, p.N, rx, ry, rn, 64, a.bw, a.bh, true, a.bn)) return cudaErrorInvalidValue;
  }
  const size_t smem = fused_smem_bytes(a);
  const int total_tiles = a.m_tiles * a.nt.nt * nvar;
  int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (a.pair) {                          // clusters of 2 walk the pair-item list; an odd cluster count keeps the 4 phases balanced
    int ncl = num_sms >> 1;
    const int items_pair = (a.m_tiles >> 1) * a.nt.nt * nvar;
    if (ncl > items_pair) ncl = items_pair;
    if (nvar == 4 && ncl > 1 && !(ncl & 1)) --ncl;
    grid = 2 * ncl;
  } else if (nvar == 4 && grid > 1 && !(grid & 1)) --grid;   // a CTA strides the item list by the grid size: keep it odd so every CTA sees all 4 phases (8/4/4/2 k-blocks)
#define EMD_DISPATCH(TT)                                                                                                   \
  (a.dw_mode ? (a.has_res ? launch_t<TT, true, true, false>(a, tin, tout, tres, tw, grid, smem, s)                         \
                          : launch_t<TT, true, false, false>(a, tin, tout, tres, tw, grid, smem, s))                       \
   : a.pair  ? (a.has_res ? launch_t<TT, false, true, true>(a, tin, tout, tres, tw, grid, smem, s)                         \
                          : launch_t<TT, false, false, true>(a, tin, tout, tres, tw, grid, smem, s))                       \
             : (a.has_res ? launch_t<TT, false, true, false>(a, tin, tout, tres, tw, grid, smem, s)                        \
                          : launch_t<TT, false, false, false>(a, tin, tout, tres, tw, grid, smem, s)))
  last_launch_kind() = a.dw_mode ? LK_FUSED_DW : (a.pair ? LK_FUSED_PAIR : LK_FUSED_TAPS);
  if (bf16) return EMD_DISPATCH(__nv_bfloat16);
  return EMD_DISPATCH(__half);
#undef EMD_DISPATCH
}

}  // namespace emd
