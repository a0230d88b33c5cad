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

#include <cstring>

namespace emd {

namespace {

constexpr int kBM = 128;            // output pixels per tile (UMMA M)
constexpr int kBK = 64;             // channels per K block (128 bytes = one swizzle row)
constexpr int kAStageBytes = kBM * kBK * 2;
constexpr int kThreads = 320;
constexpr int kLag = 3;             // A-loader look-ahead in stages
constexpr int kMaxNTiles = 4;
constexpr int kMaxCout = 1024;
constexpr int kSmemLimit = 227 * 1024;

struct NTiling { int nt; int n0[kMaxNTiles]; int rows[kMaxNTiles]; int rows_before[kMaxNTiles]; int maxrows; };

inline NTiling make_ntiling(int Cout) {
  NTiling t;
  const int cpad = (Cout + 15) & ~15;
  t.nt = (cpad + 255) / 256;
  const int base = (((cpad + t.nt - 1) / t.nt) + 15) & ~15;
  t.maxrows = 0;
  int acc = 0;
  for (int i = 0; i < kMaxNTiles; ++i) { t.n0[i] = 0; t.rows[i] = 0; t.rows_before[i] = 0; }
  for (int i = 0; i < t.nt; ++i) {
    t.n0[i] = i * base;
    t.rows[i] = (cpad - i * base) < base ? (cpad - i * base) : base;
    t.rows_before[i] = acc;
    acc += t.rows[i];
    if (t.rows[i] > t.maxrows) t.maxrows = t.rows[i];
  }
  return t;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (BF16 or FP16 operands, FP32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >>4 | LBO (ignored for swizzled K-major) | SBO = 1024 B between 8-row groups |
// version 1 (sm_100) | layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
  static constexpr uint32_t kFmt = 1;  // UMMA F16F32Format::BF16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
};
template <> struct Cvt<__half> {
  static constexpr uint32_t kFmt = 0;  // UMMA F16F32Format::F16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};

struct UmmaArgs {
  ConvParams p;
  NTiling nt;
  int stages, b_stage_bytes, nchunks, m_tiles, w_kblocks;
  long long M;
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = a.p;
  // carve (1024-byte aligned for the 128B swizzle): [stages x A][stages x B][scale][shift][barriers][tmem slot]
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_u32);
  const int S = a.stages;
  const uint32_t sA = smem_base, sB = smem_base + (uint32_t)S * kAStageBytes;
  float* s_scale = reinterpret_cast<float*>(smem + (size_t)S * (kAStageBytes + a.b_stage_bytes));
  float* s_shift = s_scale + kMaxCout;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + kMaxCout);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8u * S, bar_tfull = bar_empty + 8u * S,
                 bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int kblocks = p.ntaps * a.nchunks;

  for (int i = threadIdx.x; i < kMaxCout; i += kThreads) {
    s_scale[i] = i < p.Cout ? p.scale[i] : 0.f;
    s_shift[i] = i < p.Cout ? p.shift[i] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8u * s, 128 + 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 6) {
    // ===================== A loaders: im2col gather into the swizzled stage =====================
    const int ta = threadIdx.x - 192;        // 0..127
    const int q = ta & 7, rg = ta >> 3;      // 16-byte chunk within the 128-byte row; row group (rows rg + 16 i)
    const uint32_t dst_thread = (uint32_t)rg * 128u + (uint32_t)((q ^ (rg & 7)) << 4);
    const char* in_base = reinterpret_cast<const char*>(p.in.ptr);
    const long long hw = (long long)p.MH * p.MW;
    int it = 0;  // flat k-block counter across tiles (ring position)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / a.nt.nt;
      const long long m0 = (long long)mt * kBM;
      int iy0[8], ix0[8];
      long long img0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long m = m0 + rg + 16 * i;
        if (m < a.M) {
          const int n_img = (int)(m / hw);
          const int rem = (int)(m - (long long)n_img * hw);
          const int my = rem / p.MW;
          iy0[i] = my * p.istride; ix0[i] = (rem - my * p.MW) * p.istride;
          img0[i] = (long long)n_img * p.in.H * p.in.W;
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
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int iy = iy0[i] + dy, ix = ix0[i] + dx;
          const bool ok = ch_ok && iy >= 0 && iy < p.in.H && ix >= 0 && ix < p.in.W;
          const char* src = ok ? in_base + ((img0[i] + (long long)iy * p.in.W + ix) * p.in.pitch + p.in.coff + ch) * 2 : in_base;
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
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int t = kb / a.nchunks, c = kb - t * a.nchunks;
          const int s = it % S;
          const uint32_t ph = (uint32_t)((it / S) & 1);
          mbar_wait(bar_empty + 8u * s, ph ^ 1u);
          const uint32_t bar = bar_full + 8u * s;
          mbar_arrive_expect_tx(bar, bytes);
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
        if (lane == 0) {
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
    const long long hw = (long long)p.MH * p.MW;
    const int row = warp * 32 + lane;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / a.nt.nt, ntile = tile - mt * a.nt.nt;
      const int n0 = a.nt.n0[ntile], n = a.nt.rows[ntile];
      const long long m = (long long)mt * kBM + row;
      const bool valid = m < a.M;
      size_t pix = 0;
      if (valid) {
        const int n_img = (int)(m / hw);
        const int rem = (int)(m - (long long)n_img * hw);
        const int my = rem / p.MW;
        const int oy = my * p.ostride + p.oy0, ox = (rem - my * p.MW) * p.ostride + p.ox0;
        pix = ((size_t)n_img * p.out.H + oy) * p.out.W + ox;
      }
      T* orow = reinterpret_cast<T*>(p.out.ptr) + pix * p.out.pitch + p.out.coff;
      const T* rrow = p.res.ptr ? reinterpret_cast<const T*>(p.res.ptr) + pix * p.res.pitch + p.res.coff : nullptr;
      const int acc = tcount & 1;
      const uint32_t acc_ph = (uint32_t)((tcount >> 1) & 1);
      mbar_wait(bar_tfull + 8u * acc, acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)acc * 256u;
      for (int c0 = 0; c0 < n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        const int cg = n0 + c0;  // global output channel of v[0]
        if (valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {          // two 8-channel halves = two 16-byte stores
            const int ch = cg + 8 * h;
            if (ch >= p.Cout) break;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = fmaf(__uint_as_float(v[8 * h + j]), s_scale[ch + j], s_shift[ch + j]);
              if (p.relu6) x = fminf(fmaxf(x, 0.f), 6.f);
              y[j] = x;
            }
            if (rrow) {
              const uint4 r = *reinterpret_cast<const uint4*>(rrow + ch);
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = Cvt<T>::unpack(rw[j]);
                y[2 * j] += f.x; y[2 * j + 1] += f.y;
              }
            }
            uint4 o;
            o.x = Cvt<T>::pack(y[0], y[1]); o.y = Cvt<T>::pack(y[2], y[3]);
            o.z = Cvt<T>::pack(y[4], y[5]); o.w = Cvt<T>::pack(y[6], y[7]);
            *reinterpret_cast<uint4*>(orow + ch) = o;
          }
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
  return 1024 + (size_t)stages * (kAStageBytes + b_stage_bytes) + 2 * kMaxCout * sizeof(float) + (2 * stages + 4) * 8 + 16;
}

}  // namespace

bool umma_supported(const ConvParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (!p.w16 || p.in_f32 || p.out_f32 || p.clip01) return false;
  if (p.Cout < 16 || p.Cout > kMaxCout || (p.Cout & 7)) return false;
  if ((p.Cin & 7) || (p.in.pitch & 7) || (p.in.coff & 7) || (p.out.pitch & 7) || (p.out.coff & 7)) return false;
  if (p.res.ptr && ((p.res.pitch & 7) || (p.res.coff & 7))) return false;
  if (make_ntiling(p.Cout).nt > kMaxNTiles) return false;
  return true;
}

cudaError_t launch_conv_umma(const ConvParams& p, int et, int num_sms, cudaStream_t s) {
  UmmaArgs a;
  a.p = p;
  a.nt = make_ntiling(p.Cout);
  a.nchunks = (p.Cin + kBK - 1) / kBK;
  a.w_kblocks = p.wtaps * a.nchunks;
  a.M = (long long)p.N * p.MH * p.MW;
  a.m_tiles = (int)((a.M + kBM - 1) / kBM);
  a.b_stage_bytes = ((a.nt.maxrows * 128) + 1023) & ~1023;
  int stages = 8;
  while (stages > 4 && smem_bytes_for(stages, a.b_stage_bytes) > (size_t)kSmemLimit) --stages;
  a.stages = stages;
  const size_t smem = smem_bytes_for(stages, a.b_stage_bytes);
  const int total_tiles = a.m_tiles * a.nt.nt;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  static bool attr_done[2] = {false, false};
  if (et == ET_BF16) {
    if (!attr_done[0]) {
      cudaError_t r = cudaFuncSetAttribute(conv_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (r != cudaSuccess) return r;
      attr_done[0] = true;
    }
    conv_umma_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(a);
  } else {
    if (!attr_done[1]) {
      cudaError_t r = cudaFuncSetAttribute(conv_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (r != cudaSuccess) return r;
      attr_done[1] = true;
    }
    conv_umma_kernel<__half><<<grid, kThreads, smem, s>>>(a);
  }
  return cudaGetLastError();
}

// Host-side packing of the B operand: FP32 [wtaps*Cin][Cout] -> 16-bit blocks
// [n_tile][tap*nchunks + chunk][rows x 64 channels], each row 128 bytes with its 16-byte chunks
// XOR-swizzled by (row & 7) -- exactly the image the UMMA descriptor (SWIZZLE_128B, K-major) reads, so
// one linear cp.async.bulk per stage brings it in.  Rows >= Cout and channels >= Cin are zero.
size_t umma_pack_weights(const float* w, int wtaps, int Cin, int Cout, int et, void* dst_host) {
  if ((et != ET_BF16 && et != ET_F16) || (Cin & 7) || Cout < 16 || (Cout & 7) || Cout > kMaxCout) return 0;
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
