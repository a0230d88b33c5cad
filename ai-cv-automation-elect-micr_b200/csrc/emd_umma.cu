// emd_umma.cu -- tcgen05 / TMEM implicit-GEMM convolution for sm_100a (16-bit operands, FP32 accumulate).
//
// One persistent, warp-specialised kernel covers every GEMM-class layer of the denoiser graph:
// pointwise halves of strided_conv_block (DMG:250-276), residual_conv 1x1 stride 2 (DMG:363-373),
// the ASPP 1x1 / dilated 3x3 / pellet convolutions (DMG:291-361), conv_block_not_sep 1x1
// (DMG:225-238) and the four sub-pixel phases of deconv_block (DMG:278-289).
//
//   D[m, co] = sum_{tap t} sum_{ci}  X[pix(m,t), ci] * W[t][ci][co]         (ConvParams, emd_kernels.h)
//
// CTA tile: 128 output pixels x n_tile (<=256) output channels; K advances in blocks of 64 channels
// of one tap.  Roles (10 warps):
//   warps 0-3  epilogue   tcgen05.ld accumulator rows -> folded BN scale/shift -> ReLU6 -> (+residual)
//                         -> 16-bit NHWC store (channel slice of a concat buffer if asked)
//   warp  4    MMA        one elected thread issues tcgen05.mma (M=128, N=n_tile, K=16) from the
//                         128B-swizzled smem stages into a double-buffered TMEM accumulator
//   warp  5    B loader   cp.async.bulk (TMA bulk copy) of the pre-swizzled weight block of the stage
//   warps 6-9  A loaders  im2col gather: 16-byte cp.async per (pixel, 8 channels) with zero fill for
//                         SAME padding / image borders / K tail, written straight into the swizzled
//                         layout; generic-proxy writes are fenced (fence.proxy.async) before the
//                         stage's mbarrier is signalled
// Pipelines: smem ring full/empty mbarriers (loaders <-> MMA), TMEM full/empty mbarriers
// (MMA <-> epilogue); the epilogue of tile i overlaps the MMAs of tile i+1.
#include "emd_kernels.h"
#include "emd_tma.h"

#include <cstring>

namespace emd {

namespace {

using namespace ptx;

constexpr int kBM = 128;            // output pixels per tile (UMMA M)
constexpr int kBK = 64;             // channels per K block (128 bytes = one swizzle row)
constexpr int kAStageBytes = kBM * kBK * 2;
constexpr int kThreads = 320;
constexpr int kLag = 3;             // A-loader look-ahead in stages
constexpr int kMaxCout = 1024;
constexpr int kSmemLimit = 227 * 1024;

// ---------------------------------------------------------------------------------------------
// PTX wrappers: the shared ones live in emd_tma.h (namespace ptx); only what this first-generation kernel alone uses is here
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
  static constexpr uint32_t kFmt = 1;  // UMMA F16F32Format::BF16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
  static __device__ __forceinline__ __nv_bfloat16 one(float a) { return __float2bfloat16_rn(a); }
};
template <> struct Cvt<__half> {
  static constexpr uint32_t kFmt = 0;  // UMMA F16F32Format::F16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static __device__ __forceinline__ __half one(float a) { return __float2half_rn(a); }
};

struct UmmaArgs {
  ConvParams p;
  NTiling nt;
  int stages, b_stage_bytes, nchunks, m_tiles, w_kblocks;
  int tma_mode;            // 1: A tiles come through the tensor map (block mode only); 0: cp.async gather
  int block_mode;          // 1: a tile is an 8 x 16 pixel block of one image; 0: 128 consecutive pixels
  int tiles_x, tiles_per_img;
  unsigned M, hw;          // output-grid pixels in total / per image
};

constexpr int kTH = 8, kTW = 16;   // block-mode tile shape (kTH * kTW == kBM)

// row r of M-tile mt -> (image, y, x) on the virtual output grid; false past the end (linear mode tail)
__device__ __forceinline__ bool row_coords(const UmmaArgs& a, int mt, int r, int& n_img, int& y, int& x) {
  if (a.block_mode) {
    n_img = mt / a.tiles_per_img;
    const int rem = mt - n_img * a.tiles_per_img;
    const int by = rem / a.tiles_x;
    y = by * kTH + (r >> 4);
    x = (rem - by * a.tiles_x) * kTW + (r & 15);
    return true;
  }
  const unsigned m = (unsigned)mt * kBM + (unsigned)r;
  if (m >= a.M) return false;
  n_img = (int)(m / a.hw);
  const unsigned rem = m - (unsigned)n_img * a.hw;
  y = (int)(rem / (unsigned)a.p.MW);
  x = (int)(rem - (unsigned)y * a.p.MW);
  return true;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs a, const __grid_constant__ CUtensorMap tmap_a) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = a.p;
  // carve (1024-byte aligned for the 128B swizzle):
  //   [stages x A][stages x B][epilogue staging 4 x 4 KB][scale][shift][barriers][tmem slot]
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_u32);
  const int S = a.stages;
  const uint32_t sA = smem_base, sB = smem_base + (uint32_t)S * kAStageBytes;
  uint8_t* s_stage = smem + (size_t)S * (kAStageBytes + a.b_stage_bytes);
  float* s_scale = reinterpret_cast<float*>(s_stage + 4 * 4096);
  float* s_shift = s_scale + kMaxCout;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + kMaxCout);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8u * S, bar_tfull = bar_empty + 8u * S,
                 bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform roles (emd_fused.cu)
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int kblocks = p.ntaps * a.nchunks;

  for (int i = threadIdx.x; i < kMaxCout; i += kThreads) {
    s_scale[i] = i < p.Cout ? p.scale[i] : 0.f;
    s_shift[i] = i < p.Cout ? p.shift[i] : 0.f;
  }
  if (threadIdx.x == 0) {
    const uint32_t full_count = a.tma_mode ? 1u : 128u + 1u;  // TMA: one expect_tx arrive; gather: 128 loaders + B
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8u * s, full_count); mbar_init(bar_empty + 8u * s, 1); }
    if (a.tma_mode) prefetch_tmap(&tmap_a);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 6) {
    if (!a.tma_mode) {
    // ===================== A loaders: im2col gather into the swizzled stage =====================
    const int ta = threadIdx.x - 192;        // 0..127
    const int q = ta & 7, rg = ta >> 3;      // 16-byte chunk within the 128-byte row; rows rg + 16 i
    const uint32_t dst_thread = (uint32_t)rg * 128u + (uint32_t)((q ^ (rg & 7)) << 4);
    const char* in_base = reinterpret_cast<const char*>(p.in.ptr);
    const int img_px = p.in.H * p.in.W;
    int it = 0;  // flat k-block counter across tiles (ring position)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / a.nt.nt;
      int iy0[8], ix0[8], img0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n_img, y, x;
        if (row_coords(a, mt, rg + 16 * i, n_img, y, x)) {
          iy0[i] = y * p.istride; ix0[i] = x * p.istride; img0[i] = n_img * img_px;
        } else {
          iy0[i] = -(1 << 28); ix0[i] = 0; img0[i] = 0;
        }
      }
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int t = kb / a.nchunks, c = kb - t * a.nchunks;
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        const int ch = c * kBK + q * 8;
        const bool ch_ok = ch < p.Cin;
        const int dy = p.dy[t], dx = p.dx[t];
        const uint32_t dst = sA + (uint32_t)s * kAStageBytes + dst_thread;
        const int choff = p.in.coff + ch;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int iy = iy0[i] + dy, ix = ix0[i] + dx;
          const bool ok = ch_ok && iy >= 0 && iy < p.in.H && ix >= 0 && ix < p.in.W;
          const char* src = ok ? in_base + ((size_t)(img0[i] + iy * p.in.W + ix) * p.in.pitch + choff) * 2 : in_base;
          cp_async_16_zfill(dst + (uint32_t)i * 16u * 128u, src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it >= kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async();
          mbar_arrive(bar_full + 8u * ((it - kLag) % S));
        }
      }
    }
    // drain the look-ahead
    cp_async_wait<0>();
    fence_proxy_async();
    for (int j = (it > kLag ? it - kLag : 0); j < it; ++j) mbar_arrive(bar_full + 8u * (j % S));
    }
  } else if (warp == 5) {
    // ===================== B loader: one bulk copy of the pre-swizzled weight block per stage =====================
    if (lane == 0) {
      const char* wbase = reinterpret_cast<const char*>(p.w16);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ntile = tile % a.nt.nt;
        const uint32_t rows = (uint32_t)a.nt.rows[ntile];
        const uint32_t bytes = rows * 128u;
        // packed layout: [n_tile][weight k-block = tap*nchunks + chunk][rows x 128 B, swizzled]
        const char* tbase = wbase + (size_t)a.nt.rows_before[ntile] * a.w_kblocks * 128;
        int n_img = 0, y0 = 0, x0 = 0;
        if (a.tma_mode) row_coords(a, tile / a.nt.nt, 0, n_img, y0, x0);   // top-left pixel of the 8 x 16 block
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int t = kb / a.nchunks, c = kb - t * a.nchunks;
          const int s = it % S;
          const uint32_t ph = (uint32_t)((it / S) & 1);
          mbar_wait(bar_empty + 8u * s, ph ^ 1u);
          const uint32_t bar = bar_full + 8u * s;
          mbar_arrive_expect_tx(bar, bytes + (a.tma_mode ? (uint32_t)kAStageBytes : 0u));
          if (a.tma_mode)
            tma_load_4d(sA + (uint32_t)s * kAStageBytes, &tmap_a, c * kBK, x0 * p.istride + p.dx[t], y0 * p.istride + p.dy[t],
                        n_img, bar);
          bulk_g2s(sB + (uint32_t)s * a.b_stage_bytes, tbase + (size_t)(p.wrow[t] * a.nchunks + c) * bytes, bytes, bar);
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int ntile = tile % a.nt.nt;
      const uint32_t n = (uint32_t)a.nt.rows[ntile];
      // instruction descriptor: D=F32 (bits 4-5), A/B format (bits 7-9 / 10-12), K-major A and B,
      // N>>3 at bits 17-22, M>>4 at bits 24-28
      const uint32_t idesc = (1u << 4) | (Cvt<T>::kFmt << 7) | (Cvt<T>::kFmt << 10) | ((n >> 3) << 17) | ((kBM >> 4) << 24);
      const int acc = tcount & 1;
      const uint32_t acc_ph = (uint32_t)((tcount >> 1) & 1);
      mbar_wait(bar_tempty + 8u * acc, acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int c = kb % a.nchunks;
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        mbar_wait(bar_full + 8u * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const int kvalid = min(kBK, p.Cin - c * kBK);
          const int ksteps = (kvalid + 15) >> 4;
          const uint64_t adesc = make_sdesc(sA + (uint32_t)s * kAStageBytes);
          const uint64_t bdesc = make_sdesc(sB + (uint32_t)s * a.b_stage_bytes);
          for (int ks = 0; ks < ksteps; ++ks)  // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
            umma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
          umma_commit(bar_empty + 8u * s);                       // frees the smem stage when these MMAs retire
          if (kb == kblocks - 1) umma_commit(bar_tfull + 8u * acc);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 0-3; warp w owns TMEM lanes 32w..32w+31) =====================
    // TMEM -> registers (one pixel row per lane) -> folded BN/ReLU6 -> FP32 transpose through a
    // warp-private 4 KB smem tile -> (+ residual) -> 16-bit stores where 4 lanes cover 64 contiguous bytes.
    const int row = warp * 32 + lane;
    float* stg = reinterpret_cast<float*>(s_stage + warp * 4096);
    const bool narrow = p.out_f32 || (p.Cout & 7);
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / a.nt.nt, ntile = tile - mt * a.nt.nt;
      const int n0 = a.nt.n0[ntile], n = a.nt.rows[ntile];
      int n_img, y, x;
      const bool valid = row_coords(a, mt, row, n_img, y, x);
      unsigned pix = 0;
      if (valid) pix = (unsigned)((n_img * p.out.H + y * p.ostride + p.oy0) * p.out.W + x * p.ostride + p.ox0);
      // coalesced-phase geometry: lane -> (pixel 8*itr + lane/4, channels 8*(lane%4) .. +8) of each 32-column group
      const int sub = lane & 3;
      unsigned ppix[4];
      bool pvalid[4];
#pragma unroll
      for (int itr = 0; itr < 4; ++itr) {
        const int pr = itr * 8 + (lane >> 2);
        ppix[itr] = __shfl_sync(0xffffffffu, pix, pr);
        pvalid[itr] = __shfl_sync(0xffffffffu, (int)valid, pr) != 0;
      }
      // residual operand: prefetched one 32-column group ahead (the first group before the accumulator is awaited)
      const T* resp = reinterpret_cast<const T*>(p.res.ptr);
      uint4 rnext[4];
      auto load_res = [&](int c0, uint4 (&r)[4]) {
        const int cgl = n0 + c0 + 8 * sub;
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          r[itr] = make_uint4(0u, 0u, 0u, 0u);
          if (resp && pvalid[itr] && c0 + 8 * sub < n && cgl < p.Cout)
            r[itr] = *reinterpret_cast<const uint4*>(resp + (size_t)ppix[itr] * p.res.pitch + p.res.coff + cgl);
        }
      };
      if (!narrow) load_res(0, rnext);
      const int acc = tcount & 1;
      const uint32_t acc_ph = (uint32_t)((tcount >> 1) & 1);
      mbar_wait(bar_tfull + 8u * acc, acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)acc * 256u;
      if (narrow) {
        // narrow / FP32 outputs (the final 64->1 conv, DMG:531-538): scalar stores straight from registers
        for (int c0 = 0; c0 < n; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          const int cg = n0 + c0;
          if (valid) {
            for (int j = 0; j < 16 && cg + j < p.Cout; ++j) {
              float xv = fmaf(__uint_as_float(v[j]), s_scale[cg + j], s_shift[cg + j]);
              if (p.relu6) xv = fminf(fmaxf(xv, 0.f), 6.f);
              if (p.clip01) xv = fminf(fmaxf(xv, 0.f), 1.f);
              const size_t o = (size_t)pix * p.out.pitch + p.out.coff + cg + j;
              if (p.out_f32) reinterpret_cast<float*>(p.out.ptr)[o] = xv;
              else reinterpret_cast<T*>(p.out.ptr)[o] = Cvt<T>::one(xv);
            }
          }
        }
      } else {
        for (int c0 = 0; c0 < n; c0 += 32) {
          const int ncol = min(32, n - c0);   // 16 or 32 (n is a multiple of 16)
          uint4 rcur[4];
#pragma unroll
          for (int itr = 0; itr < 4; ++itr) rcur[itr] = rnext[itr];
          if (c0 + 32 < n) load_res(c0 + 32, rnext);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 16 < ncol) {
              uint32_t v[16];
              tmem_ld16(taddr + (uint32_t)(c0 + 16 * h), v);
              tmem_ld_wait();
              const int cg = n0 + c0 + 16 * h;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                float4 o;
                float* po = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float xv = fmaf(__uint_as_float(v[4 * j4 + j]), s_scale[cg + 4 * j4 + j], s_shift[cg + 4 * j4 + j]);
                  if (p.relu6) xv = fminf(fmaxf(xv, 0.f), 6.f);
                  if (p.clip01) xv = fminf(fmaxf(xv, 0.f), 1.f);
                  po[j] = xv;
                }
                // row `lane`, 16-byte chunk (4h + j4), XOR-swizzled by the row
                *reinterpret_cast<float4*>(stg + lane * 32 + (((4 * h + j4) ^ (lane & 7)) << 2)) = o;
              }
            }
          }
          __syncwarp();
          const int cg = n0 + c0 + 8 * sub;
#pragma unroll
          for (int itr = 0; itr < 4; ++itr) {
            const int pr = itr * 8 + (lane >> 2);
            if (pvalid[itr] && 8 * sub < ncol && cg < p.Cout) {
              const float4 f0 = *reinterpret_cast<const float4*>(stg + pr * 32 + (((2 * sub) ^ (pr & 7)) << 2));
              const float4 f1 = *reinterpret_cast<const float4*>(stg + pr * 32 + (((2 * sub + 1) ^ (pr & 7)) << 2));
              float yv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
              if (resp) {
                const uint32_t rw[4] = {rcur[itr].x, rcur[itr].y, rcur[itr].z, rcur[itr].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = Cvt<T>::unpack(rw[j]);
                  yv[2 * j] += f.x; yv[2 * j + 1] += f.y;
                }
              }
              uint4 o;
              o.x = Cvt<T>::pack(yv[0], yv[1]); o.y = Cvt<T>::pack(yv[2], yv[3]);
              o.z = Cvt<T>::pack(yv[4], yv[5]); o.w = Cvt<T>::pack(yv[6], yv[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.out.ptr) + (size_t)ppix[itr] * p.out.pitch + p.out.coff + cg) = o;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * acc);
    }
  }

  // teardown: everyone done with TMEM, then the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

size_t smem_bytes_for(int stages, int b_stage_bytes) {
  return 1024 + (size_t)stages * (kAStageBytes + b_stage_bytes) + 4 * 4096 + 2 * kMaxCout * sizeof(float) +
         (2 * stages + 4) * 8 + 16;
}

}  // namespace

bool umma_supported(const ConvParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (!p.w16 || p.in_f32) return false;
  if (p.Cout < 1 || p.Cout > kMaxCout) return false;
  if ((p.Cout & 7) && p.res.ptr) return false;
  if ((p.Cin & 7) || (p.in.pitch & 7) || (p.in.coff & 7)) return false;
  if (!(p.out_f32 || (p.Cout & 7)) && ((p.out.pitch & 7) || (p.out.coff & 7))) return false;
  if (p.res.ptr && ((p.res.pitch & 7) || (p.res.coff & 7))) return false;
  if (make_ntiling(p.Cout).nt > kMaxNTiles) return false;
  if ((long long)p.N * p.MH * p.MW >= (1ll << 31) || (long long)p.N * p.in.H * p.in.W >= (1ll << 31) ||
      (long long)p.N * p.out.H * p.out.W >= (1ll << 31))
    return false;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    cudaGetLastError();
  }
  return fn;
}

// Tensor map over the input activation view: dims (C, W, H, N), box = 64 channels x 16 x 8 pixels (times the
// input stride: a 1x1 stride-2 conv walks the map with element stride 2), 128B swizzle, zero OOB fill.
static bool encode_a_map(const ConvParams& p, int et, CUtensorMap* map) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.in.W, (cuuint64_t)p.in.H, (cuuint64_t)p.N};
  const cuuint64_t strides[3] = {(cuuint64_t)p.in.pitch * 2, (cuuint64_t)p.in.W * p.in.pitch * 2,
                                 (cuuint64_t)p.in.H * p.in.W * p.in.pitch * 2};
  const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(kTW * p.istride), (cuuint32_t)(kTH * p.istride), 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)p.istride, (cuuint32_t)p.istride, 1};
  void* base = reinterpret_cast<char*>(p.in.ptr) + (size_t)p.in.coff * 2;
  CUresult r = enc(map, et == ET_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


cudaError_t launch_conv_umma(const ConvParams& p, int et, int num_sms, cudaStream_t s) {
  UmmaArgs a;
  a.p = p;
  a.nt = make_ntiling(p.Cout);
  a.nchunks = (p.Cin + kBK - 1) / kBK;
  a.w_kblocks = p.wtaps * a.nchunks;
  a.M = (unsigned)p.N * p.MH * p.MW;
  a.hw = (unsigned)p.MH * p.MW;
  a.block_mode = (p.MH % kTH == 0 && p.MW % kTW == 0) ? 1 : 0;
  a.tiles_x = p.MW / kTW;
  a.tiles_per_img = a.block_mode ? (p.MH / kTH) * a.tiles_x : 0;
  a.m_tiles = a.block_mode ? p.N * a.tiles_per_img : (int)((a.M + kBM - 1) / kBM);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  a.tma_mode = (tuning().tma && a.block_mode && kTW * p.istride <= 256 && encode_a_map(p, et, &tmap)) ? 1 : 0;
  a.b_stage_bytes = ((a.nt.maxrows * 128) + 1023) & ~1023;
  int stages = 8;
  while (stages > 4 && smem_bytes_for(stages, a.b_stage_bytes) > (size_t)kSmemLimit) --stages;
  a.stages = stages;
  const size_t smem = smem_bytes_for(stages, a.b_stage_bytes);
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  static thread_local int attr_dev[2] = {-1, -1};  // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (et == ET_BF16) {
    if (attr_dev[0] != dev) {
      cudaError_t r = cudaFuncSetAttribute(conv_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (r != cudaSuccess) return r;
      attr_dev[0] = dev;
    }
    conv_umma_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(a, tmap);
  } else {
    if (attr_dev[1] != dev) {
      cudaError_t r = cudaFuncSetAttribute(conv_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (r != cudaSuccess) return r;
      attr_dev[1] = dev;
    }
    conv_umma_kernel<__half><<<grid, kThreads, smem, s>>>(a, tmap);
  }
  return cudaGetLastError();
}

// Host-side packing of the B operand: FP32 [wtaps*Cin][Cout] -> 16-bit blocks
// [n_tile][tap*nchunks + chunk][rows x 64 channels], each row 128 bytes with its 16-byte chunks
// XOR-swizzled by (row & 7) -- exactly the image the UMMA descriptor (SWIZZLE_128B, K-major) reads, so
// one linear cp.async.bulk per stage brings it in.  Rows >= Cout and channels >= Cin are zero.
size_t umma_pack_weights(const float* w, int wtaps, int Cin, int Cout, int et, void* dst_host) {
  if ((et != ET_BF16 && et != ET_F16) || (Cin & 7) || Cout < 1 || Cout > kMaxCout) return 0;
  const NTiling nt = make_ntiling(Cout);
  if (nt.nt > kMaxNTiles) return 0;
  const int nchunks = (Cin + kBK - 1) / kBK;
  const size_t wkb = (size_t)wtaps * nchunks;
  size_t total = 0;
  for (int j = 0; j < nt.nt; ++j) total += (size_t)nt.rows[j] * wkb * 128;
  if (!dst_host) return total;
  uint16_t* out = reinterpret_cast<uint16_t*>(dst_host);
  memset(out, 0, total);
  for (int j = 0; j < nt.nt; ++j) {
    const size_t tbase = (size_t)nt.rows_before[j] * wkb * 64;  // in 16-bit elements
    for (int t = 0; t < wtaps; ++t)
      for (int c = 0; c < nchunks; ++c) {
        const size_t bbase = tbase + ((size_t)t * nchunks + c) * nt.rows[j] * 64;
        for (int r = 0; r < nt.rows[j]; ++r) {
          const int co = nt.n0[j] + r;
          if (co >= Cout) continue;
          for (int qq = 0; qq < 8; ++qq)
            for (int e = 0; e < 8; ++e) {
              const int ch = c * kBK + qq * 8 + e;
              if (ch >= Cin) continue;
              const float v = w[((size_t)t * Cin + ch) * Cout + co];
              uint16_t bits;
              if (et == ET_BF16) { __nv_bfloat16 h = __float2bfloat16_rn(v); memcpy(&bits, &h, 2); }
              else { __half h = __float2half_rn(v); memcpy(&bits, &h, 2); }
              out[bbase + (size_t)r * 64 + (size_t)((qq ^ (r & 7)) * 8) + e] = bits;
            }
        }
      }
  }
  return total;
}

}  // namespace emd
