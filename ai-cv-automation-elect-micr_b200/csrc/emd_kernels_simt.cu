// emd_kernels_simt.cu -- CUDA-core kernels.
//
//  * conv_simt_kernel: FP32-accumulate implicit-GEMM convolution on CUDA cores.  It is the FP32
//    validation mode of every GEMM-class layer (strided_conv_block pointwise DMG:250-276,
//    residual_conv DMG:363-373, ASPP DMG:291-361, conv_block_not_sep DMG:225-238, deconv_block
//    DMG:278-289 as 4 sub-pixel phases) and the fallback of the 16-bit modes for shapes the
//    tcgen05 kernel does not take.
//  * dw3x3_kernel: depthwise 3x3 (stride / rate in the depthwise stage, App. A.2), memory-bound.
//  * resize / avgpool / cast: memory-bound helpers (DMG:331-345, 494).
#include "emd_kernels.h"
#include "emd_tma.h"

namespace emd {

// ---------------------------------------------------------------------------------------------
// element helpers
// ---------------------------------------------------------------------------------------------
template <typename T> struct ElemOf;
template <> struct ElemOf<float> { static constexpr int et = ET_F32; };
template <> struct ElemOf<__nv_bfloat16> { static constexpr int et = ET_BF16; };
template <> struct ElemOf<__half> { static constexpr int et = ET_F16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// two 16-bit values in one 32-bit word <-> two floats
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }

template <typename T, int V> struct VecIO;
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIO<float, 8> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    VecIO<float, 4>::ld(p, v); VecIO<float, 4>::ld(p + 4, v + 4);
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    VecIO<float, 4>::st(p, v); VecIO<float, 4>::st(p + 4, v + 4);
  }
};
template <typename T> struct Pair;
template <> struct Pair<__nv_bfloat16> {
  using type = __nv_bfloat162;
  static __device__ __forceinline__ float2 up(type v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ type down(float a, float b) { return __floats2bfloat162_rn(a, b); }
};
template <> struct Pair<__half> {
  using type = __half2;
  static __device__ __forceinline__ float2 up(type v) { return __half22float2(v); }
  static __device__ __forceinline__ type down(float a, float b) { return __floats2half2_rn(a, b); }
};
template <typename T> struct VecIO<T, 4> {  // 16-bit types, 8 bytes
  using P = Pair<T>;
  static __device__ __forceinline__ void ld(const T* p, float* v) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = P::up(*reinterpret_cast<typename P::type*>(&u.x));
    float2 b = P::up(*reinterpret_cast<typename P::type*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(T* p, const float* v) {
    uint2 u;
    typename P::type a = P::down(v[0], v[1]), b = P::down(v[2], v[3]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <typename T> struct VecIO<T, 8> {  // 16-bit types, 16 bytes
  using P = Pair<T>;
  static __device__ __forceinline__ void ld(const T* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t t = w[i];
      float2 a = P::up(*reinterpret_cast<typename P::type*>(&t));
      v[2 * i] = a.x; v[2 * i + 1] = a.y;
    }
  }
  static __device__ __forceinline__ void st(T* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      typename P::type a = P::down(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&a);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <typename T> struct VecIO<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, float* v) { v[0] = to_f(*p); }
  static __device__ __forceinline__ void st(T* p, const float* v) { *p = from_f<T>(v[0]); }
};

// ---------------------------------------------------------------------------------------------
// implicit-GEMM convolution, CUDA cores, FP32 accumulate.  64x64x16 tiles, 4x4 per thread.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long M = (long long)p.N * p.MH * p.MW;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load role: row lr, channels kq*4..kq*4+3 of the K chunk
  const int lr = tid >> 2, kq = tid & 3;
  const long long mrow = m0 + lr;
  const bool mvalid = mrow < M;
  int n_img = 0, my = 0, mx = 0;
  if (mvalid) {
    n_img = (int)(mrow / ((long long)p.MH * p.MW));
    int rem = (int)(mrow - (long long)n_img * p.MH * p.MW);
    my = rem / p.MW; mx = rem - my * p.MW;
  }
  const bool a_vec = ((p.in.pitch | p.in.coff | p.Cin) & 3) == 0;
  // B-load role: k row bk, columns bn..bn+3
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  const bool b_vec = (p.Cout & 3) == 0;
  // compute role
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.ntaps; ++t) {
    const int iy = my * p.istride + p.dy[t], ix = mx * p.istride + p.dx[t];
    const bool valid = mvalid && iy >= 0 && iy < p.in.H && ix >= 0 && ix < p.in.W;
    const size_t src_off = valid ? (((size_t)n_img * p.in.H + iy) * p.in.W + ix) * p.in.pitch + p.in.coff : 0;
    const size_t wbase = (size_t)p.wrow[t] * p.Cin;
    for (int c0 = 0; c0 < p.Cin; c0 += BK) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      const int c = c0 + kq * 4;
      if (valid && c < p.Cin) {
        if (p.in_f32) {
          const float* src = reinterpret_cast<const float*>(p.in.ptr) + src_off + c;
          if (a_vec) VecIO<float, 4>::ld(src, a);
          else
            for (int j = 0; j < 4; ++j) if (c + j < p.Cin) a[j] = src[j];
        } else {
          const T* src = reinterpret_cast<const T*>(p.in.ptr) + src_off + c;
          if (a_vec) VecIO<T, 4>::ld(src, a);
          else
            for (int j = 0; j < 4; ++j) if (c + j < p.Cin) a[j] = to_f(src[j]);
        }
      }
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (c0 + bk < p.Cin) {
        const float* wp = p.w + (wbase + c0 + bk) * p.Cout + n0 + bn;
        if (b_vec && n0 + bn + 3 < p.Cout) VecIO<float, 4>::ld(wp, b);
        else
          for (int j = 0; j < 4; ++j) if (n0 + bn + j < p.Cout) b[j] = wp[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kq * 4 + j][lr] = a[j];
      *reinterpret_cast<float4*>(&Bs[bk][bn]) = make_float4(b[0], b[1], b[2], b[3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: folded BN (+bias) -> ReLU6 -> clip -> + residual -> store
  const int cbase = n0 + tx * 4;
  if (cbase >= p.Cout) return;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = min(cbase + j, p.Cout - 1);
    sc[j] = p.scale[c]; sh[j] = p.shift[c];
  }
  const bool o_vec = ((p.out.pitch | p.out.coff | p.Cout) & 3) == 0;
  const bool r_vec = p.res.ptr && ((p.res.pitch | p.res.coff) & 3) == 0 && o_vec;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int ni = (int)(m / ((long long)p.MH * p.MW));
    const int rem = (int)(m - (long long)ni * p.MH * p.MW);
    const int oy = (rem / p.MW) * p.ostride + p.oy0, ox = (rem % p.MW) * p.ostride + p.ox0;
    const size_t pix = ((size_t)ni * p.out.H + oy) * p.out.W + ox;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = fmaf(acc[i][j], sc[j], sh[j]);
      if (p.relu6) x = fminf(fmaxf(x, 0.f), 6.f);
      if (p.clip01) x = fminf(fmaxf(x, 0.f), 1.f);
      v[j] = x;
    }
    if (p.res.ptr) {
      const T* rp = reinterpret_cast<const T*>(p.res.ptr) + pix * p.res.pitch + p.res.coff + cbase;
      float r[4] = {0.f, 0.f, 0.f, 0.f};
      if (r_vec) VecIO<T, 4>::ld(rp, r);
      else
        for (int j = 0; j < 4; ++j) if (cbase + j < p.Cout) r[j] = to_f(rp[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    if (p.out_f32) {
      float* op = reinterpret_cast<float*>(p.out.ptr) + pix * p.out.pitch + p.out.coff + cbase;
      if (o_vec) VecIO<float, 4>::st(op, v);
      else
        for (int j = 0; j < 4; ++j) if (cbase + j < p.Cout) op[j] = v[j];
    } else {
      T* op = reinterpret_cast<T*>(p.out.ptr) + pix * p.out.pitch + p.out.coff + cbase;
      if (o_vec) VecIO<T, 4>::st(op, v);
      else
        for (int j = 0; j < 4; ++j) if (cbase + j < p.Cout) op[j] = from_f<T>(v[j]);
    }
  }
}

cudaError_t launch_conv_simt(const ConvParams& p, int et, cudaStream_t s) {
  const long long M = (long long)p.N * p.MH * p.MW;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((p.Cout + 63) / 64));
  if (et == ET_F32) conv_simt_kernel<float><<<grid, 256, 0, s>>>(p);
  else if (et == ET_BF16) conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  else conv_simt_kernel<__half><<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 (DepthwiseConv2dNative of slim.separable_convolution2d, DMG:253-273)
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) dw3x3_kernel(const DwParams p) {
  const int C = p.in.C;
  const int cg = C / V;
  const long long total = (long long)p.N * p.OH * p.OW * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % cg);
  long long pix = idx / cg;
  const int ox = (int)(pix % p.OW); pix /= p.OW;
  const int oy = (int)(pix % p.OH);
  const int n = (int)(pix / p.OH);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  const TI* base = reinterpret_cast<const TI*>(p.in.ptr);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * p.stride - p.pad + ky * p.rate;
    if (iy < 0 || iy >= p.in.H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * p.stride - p.pad + kx * p.rate;
      if (ix < 0 || ix >= p.in.W) continue;
      float x[V], w[V];
      VecIO<TI, V>::ld(base + (((size_t)n * p.in.H + iy) * p.in.W + ix) * p.in.pitch + p.in.coff + g * V, x);
      VecIO<float, V>::ld(p.w + (ky * 3 + kx) * C + g * V, w);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(x[j], w[j], acc[j]);
    }
  }
  TO* op = reinterpret_cast<TO*>(p.out.ptr) +
           (((size_t)n * p.out.H + oy) * p.out.W + ox) * p.out.pitch + p.out.coff + g * V;
  VecIO<TO, V>::st(op, acc);
}

// Strip form for the 16-bit modes on maps that do not tile into the 8 x 16 pixel blocks of the TMA-fed kernel (emd_dw.cu) -- the
// 24^2 / 12^2 / 6^2 maps of 96 x 96 crops (small_scans shape), where the one-thread-per-pixel kernel above re-read the nine
// weights and the nine inputs of every output (21 % of the HBM peak).  A warp = the 32 channel pairs (64 channels) of one
// pixel column of one image; it slides the 3x3 window down a block of up to kStripRows rows: every LDG.32 / STG.32 of the warp
// is one 128-byte line, each input row is loaded once per column (3 loads, kAhead rows in flight) and feeds three output rows,
// the 9 x 2 weights stay in registers.  Stride 1, any rate (TF SAME: pad = rate).  Same tap order as the other depthwise kernels.
constexpr int kStripRows = 8;
template <typename T>
__global__ void __launch_bounds__(256) dw_strip_kernel(const DwParams p, int nchunks, int row_blocks, long long n_items) {
  constexpr int kAhead = 3;
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= n_items) return;
  const int lane = threadIdx.x & 31;
  long long q = item;
  const int c = (int)(q % nchunks); q /= nchunks;
  const int x = (int)(q % p.OW); q /= p.OW;
  const int rb = (int)(q % row_blocks);
  const int n = (int)(q / row_blocks);
  const int cw = c * 32 + lane;                              // channel pair
  if (2 * cw >= p.in.C) return;
  const int H = p.in.H, W = p.in.W, r = p.rate;
  const int y0 = rb * kStripRows;
  const int rows = min(kStripRows, H - y0);
  const uint32_t* gin = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.in.ptr) + p.in.coff);
  uint32_t* gout = reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.out.ptr) + p.out.coff);
  const int ipitch = p.in.pitch >> 1, opitch = p.out.pitch >> 1;
  float2 w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = __ldg(reinterpret_cast<const float2*>(p.w + t * p.in.C + 2 * cw));
  bool x_ok[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x_ok[i] = x + (i - 1) * r >= 0 && x + (i - 1) * r < W;
  const size_t ibase = (size_t)n * H * W, obase = ((size_t)n * p.OH + y0) * p.OW + x;
  // halo row j (j = 0 .. rows + 1) is image row y0 + (j - 1) * r for rate 1; for rate r the three tap rows of output row i are
  // y0 + i - r, y0 + i, y0 + i + r: walk output rows and load their three tap rows directly when r > 1
  if (r == 1) {
    // running pointers (three input columns, one output column) advanced by one image row per step: the per-load 64-bit index
    // arithmetic of the first version cost ~20 instructions per load and made the kernel issue-bound (ncu: 1059 instructions per
    // warp for 30 loads and 72 FMAs)
    const size_t irs = (size_t)W * ipitch, ors = (size_t)p.OW * opitch;
    const uint32_t* pin = gin + ((ibase + (size_t)y0 * W + x) * ipitch + cw);      // (y0, x); rows / columns outside the image are never dereferenced
    pin -= irs;                                                                   // halo row 0 = image row y0 - 1
    uint32_t* pout = gout + (obase * opitch + cw);
    uint32_t raw[kAhead][3];
    int yl = y0 - 1;                                                               // image row of the next row to load
    auto load_row = [&](int slot) {
      const bool y_ok = yl >= 0 && yl < H;
      raw[slot][0] = (y_ok && x_ok[0]) ? __ldg(pin - ipitch) : 0u;
      raw[slot][1] = y_ok ? __ldg(pin) : 0u;
      raw[slot][2] = (y_ok && x_ok[2]) ? __ldg(pin + ipitch) : 0u;
      pin += irs;
      ++yl;
    };
#pragma unroll
    for (int j = 0; j < kAhead; ++j) load_row(j);
    float2 acc[3];
#pragma unroll
    for (int j = 0; j < kStripRows + 2; ++j) {
      if (j >= rows + 2) break;
      float2 xv[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) xv[i] = unpack2<T>(raw[j % kAhead][i]);
      if (j + kAhead < kStripRows + 2 && j + kAhead < rows + 2) load_row(j % kAhead);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int o = j - ky;
        if (o < 0 || o >= kStripRows) continue;
        float2& d = acc[o % 3];
        if (ky == 0) d = make_float2(xv[0].x * w[0].x, xv[0].y * w[0].y);
        else { d.x = fmaf(xv[0].x, w[ky * 3].x, d.x); d.y = fmaf(xv[0].y, w[ky * 3].y, d.y); }
        d.x = fmaf(xv[1].x, w[ky * 3 + 1].x, d.x); d.y = fmaf(xv[1].y, w[ky * 3 + 1].y, d.y);
        d.x = fmaf(xv[2].x, w[ky * 3 + 2].x, d.x); d.y = fmaf(xv[2].y, w[ky * 3 + 2].y, d.y);
      }
      if (j >= 2 && j - 2 < rows) {
        *pout = pack2<T>(acc[(j - 2) % 3].x, acc[(j - 2) % 3].y);
        pout += ors;
      }
    }
  } else {
    for (int i = 0; i < rows; ++i) {
      float2 d = make_float2(0.f, 0.f);
      bool first = true;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y = y0 + i + (ky - 1) * r;
        const bool y_ok = y >= 0 && y < H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          uint32_t v = 0u;
          if (y_ok && x_ok[kx]) v = __ldg(gin + (ibase + (size_t)y * W + (x + (kx - 1) * r)) * ipitch + cw);
          const float2 xv = unpack2<T>(v);
          if (first) { d = make_float2(xv.x * w[0].x, xv.y * w[0].y); first = false; }
          else { d.x = fmaf(xv.x, w[ky * 3 + kx].x, d.x); d.y = fmaf(xv.y, w[ky * 3 + kx].y, d.y); }
        }
      }
      gout[(obase + (size_t)i * p.OW) * opitch + cw] = pack2<T>(d.x, d.y);
    }
  }
}

bool dw_strip_supported(const DwParams& p, int et) {
  if (et != ET_BF16 && et != ET_F16) return false;
  if (p.in_f32 || p.stride != 1 || p.pad != p.rate || p.in.H != p.OH || p.in.W != p.OW) return false;
  return !((p.in.C | p.in.pitch | p.in.coff | p.out.pitch | p.out.coff) & 1);
}

cudaError_t launch_dw_strip(const DwParams& p, int et, cudaStream_t s) {
  const int nchunks = (p.in.C + 63) / 64, row_blocks = (p.OH + kStripRows - 1) / kStripRows;
  const long long n_items = (long long)p.N * row_blocks * p.OW * nchunks;
  const unsigned blocks = (unsigned)((n_items * 32 + 255) / 256);
  if (et == ET_BF16) dw_strip_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(p, nchunks, row_blocks, n_items);
  else dw_strip_kernel<__half><<<blocks, 256, 0, s>>>(p, nchunks, row_blocks, n_items);
  return cudaGetLastError();
}

// Small maps (24^2 / 12^2 / 6^2 at 96 x 96 crops), stride 1, rate 1: a warp takes an (up to) 8 x 8 pixel tile of one image and one
// 64-channel chunk, brings its 10 x 10 pixel halo into shared memory with 4-byte cp.async (lane = channel pair: one 128-byte line
// per instruction, ~100 loads in flight per warp and no registers held -- the strip kernel's nine loads in flight per warp left
// it latency-bound at 0.3 of the HBM roofline, and re-read every column three times), then slides the 3x3 window over it in
// passes of four columns (the thread mapping of the fused kernel's depthwise producer).  Same tap order as the other kernels.
constexpr int kTileT = 8;
template <typename T>
__global__ void __launch_bounds__(256) dw_tile_kernel(const DwParams p, int nchunks, int tiles_x, int tiles_y, long long n_items, int hw, int warp_bytes) {
  extern __shared__ __align__(16) uint8_t tile_smem[];
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= n_items) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  long long q = item;
  const int c = (int)(q % nchunks); q /= nchunks;
  const int tx = (int)(q % tiles_x); q /= tiles_x;
  const int ty = (int)(q % tiles_y);
  const int n = (int)(q / tiles_y);
  const int cw = c * 32 + lane;
  const bool ch_ok = 2 * cw < p.in.C;
  const int H = p.in.H, W = p.in.W;
  const int x0 = tx * kTileT, y0 = ty * kTileT;
  const int tw = min(kTileT, W - x0), th = min(kTileT, H - y0);
  const uint32_t* gin = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.in.ptr) + p.in.coff);
  uint32_t* gout = reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.out.ptr) + p.out.coff);
  const int ipitch = p.in.pitch >> 1, opitch = p.out.pitch >> 1;
  uint32_t* sm = reinterpret_cast<uint32_t*>(tile_smem + (size_t)wib * warp_bytes) + lane;      // [halo pixel (row pitch hw)][32 lanes]
  const uint32_t sm_addr = (uint32_t)__cvta_generic_to_shared(sm);
  // halo in: one cp.async per in-bounds halo pixel, zeros elsewhere (TF SAME padding)
  const uint32_t* img = gin + ((size_t)n * H * W) * ipitch + cw;
  for (int hy = 0; hy < th + 2; ++hy) {
    const int y = y0 - 1 + hy;
    const bool y_ok = y >= 0 && y < H;
    for (int hx = 0; hx < tw + 2; ++hx) {
      const int x = x0 - 1 + hx;
      const int slot = hy * hw + hx;
      if (ch_ok && y_ok && x >= 0 && x < W)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sm_addr + (uint32_t)slot * 128u), "l"(img + ((size_t)y * W + x) * ipitch) : "memory");
      else
        sm[slot * 32] = 0u;
    }
  }
  float2 w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = ch_ok ? __ldg(reinterpret_cast<const float2*>(p.w + t * p.in.C + 2 * cw)) : make_float2(0.f, 0.f);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  if (!ch_ok) return;
  uint32_t* orow0 = gout + (((size_t)n * p.OH + y0) * p.OW + x0) * opitch + cw;
  for (int cx = 0; cx < tw; cx += 4) {              // passes of four columns
    const int ncol = min(4, tw - cx);
    float2 acc[3][4];
#pragma unroll
    for (int j = 0; j < kTileT + 2; ++j) {          // halo rows y0 - 1 + j
      if (j >= th + 2) break;
      float2 x[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) x[i] = (i < ncol + 2) ? unpack2<T>(sm[(j * hw + cx + i) * 32]) : make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int r = j - ky;
        if (r < 0 || r >= kTileT) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2& d = acc[r % 3][i];
          d = ky == 0 ? ptx::fmul2(x[i], w[0]) : ptx::ffma2(x[i], w[ky * 3], d);
          d = ptx::ffma2(x[i + 1], w[ky * 3 + 1], d);
          d = ptx::ffma2(x[i + 2], w[ky * 3 + 2], d);
        }
      }
      if (j >= 2 && j - 2 < th) {
        uint32_t* o = orow0 + ((size_t)(j - 2) * p.OW + cx) * opitch;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < ncol) o[(size_t)i * opitch] = pack2<T>(acc[(j - 2) % 3][i].x, acc[(j - 2) % 3][i].y);
      }
    }
  }
}

bool dw_tile_supported(const DwParams& p, int et) {
  return dw_strip_supported(p, et) && p.rate == 1;
}

cudaError_t launch_dw_tile(const DwParams& p, int et, cudaStream_t s) {
  const int nchunks = (p.in.C + 63) / 64, tiles_x = (p.OW + kTileT - 1) / kTileT, tiles_y = (p.OH + kTileT - 1) / kTileT;
  const long long n_items = (long long)p.N * tiles_y * tiles_x * nchunks;
  const unsigned blocks = (unsigned)((n_items + 7) / 8);
  // the halo buffer is sized for THIS map (6 x 6 maps: 8 x 8 pixels = 8 KB per warp, three blocks of eight warps per SM; 24 x 24:
  // 10 x 10 = 12.5 KB, two blocks): residency is what hides the load phase of one warp under the math of the others
  const int hw = (p.OW < kTileT ? p.OW : kTileT) + 2, hh = (p.OH < kTileT ? p.OH : kTileT) + 2;
  const int warp_bytes = hw * hh * 128;
  const size_t smem = 8 * (size_t)warp_bytes;
  static bool attr_done[2] = {false, false};
  const size_t smem_max = 8 * (size_t)(kTileT + 2) * (kTileT + 2) * 128;
  if (et == ET_BF16) {
    if (!attr_done[0]) { cudaFuncSetAttribute(dw_tile_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max); attr_done[0] = true; }
    dw_tile_kernel<__nv_bfloat16><<<blocks, 256, smem, s>>>(p, nchunks, tiles_x, tiles_y, n_items, hw, warp_bytes);
  } else {
    if (!attr_done[1]) { cudaFuncSetAttribute(dw_tile_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max); attr_done[1] = true; }
    dw_tile_kernel<__half><<<blocks, 256, smem, s>>>(p, nchunks, tiles_x, tiles_y, n_items, hw, warp_bytes);
  }
  return cudaGetLastError();
}

// Register form for the same small maps (and any stride-1, rate-1 map): a warp = the 32 channel pairs of a strip of kCW output columns of
// one image, walked down a block of up to kRegRows rows.  Nothing is staged: each halo row is kCW + 2 predicated LDG.32 (one 128-byte line
// per instruction, kAhead rows = 24-32 lines in flight per warp), lives in registers, and feeds the three output rows it touches.  Against
// the shared-memory tile kernel above there is no load-then-wait phase per warp, no 8 x 8 tile to waste on a 6 x 6 map (one strip is the
// whole map) and a third of the instructions.  Same tap order as the other depthwise kernels: bit-identical results.
constexpr int kRegRows = 12;
template <typename T, int kCW, int kAhead>
__global__ void __launch_bounds__(128) dw_reg_kernel(const DwParams p, int nchunks, int strips, int row_blocks, long long n_items) {
  constexpr int kLW = kCW + 2;
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= n_items) return;
  const int lane = threadIdx.x & 31;
  long long q = item;
  const int c = (int)(q % nchunks); q /= nchunks;
  const int st = (int)(q % strips); q /= strips;
  const int rb = (int)(q % row_blocks);
  const int n = (int)(q / row_blocks);
  const int cw = c * 32 + lane;                              // channel pair
  if (2 * cw >= p.in.C) return;
  const int H = p.in.H, W = p.in.W;
  const int x0 = st * kCW, y0 = rb * kRegRows;
  const int rows = min(kRegRows, H - y0);
  const uint32_t* gin = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.in.ptr) + p.in.coff);
  uint32_t* gout = reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.out.ptr) + p.out.coff);
  const int ipitch = p.in.pitch >> 1, opitch = p.out.pitch >> 1;
  float2 w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = __ldg(reinterpret_cast<const float2*>(p.w + t * p.in.C + 2 * cw));
  uint32_t x_ok = 0;                                         // bit i: halo column x0 - 1 + i is inside the image
#pragma unroll
  for (int i = 0; i < kLW; ++i) x_ok |= (uint32_t)(x0 - 1 + i >= 0 && x0 - 1 + i < W) << i;
  const int ncol = min(kCW, W - x0);
  const uint32_t upitch = (uint32_t)ipitch, uopitch = (uint32_t)opitch;
  const long long irs = (long long)W * ipitch, ors = (long long)p.OW * opitch;
  // running pointers: halo row 0 = image row y0 - 1, halo column 0 = image column x0 - 1 (never dereferenced outside the image)
  const uint32_t* pin = gin + (((long long)n * H + (y0 - 1)) * W + (x0 - 1)) * ipitch + cw;
  uint32_t* pout = gout + (((long long)n * p.OH + y0) * p.OW + x0) * opitch + cw;
  uint32_t raw[kAhead][kLW];
  int yl = y0 - 1;
  auto load_row = [&](int slot) {
    const uint32_t m = (yl >= 0 && yl < H) ? x_ok : 0u;
#pragma unroll
    for (int i = 0; i < kLW; ++i) raw[slot][i] = ((m >> i) & 1u) ? __ldg(pin + (uint32_t)i * upitch) : 0u;   // 32-bit offsets: one IMAD.WIDE.U32 per address
    pin += irs;
    ++yl;
  };
#pragma unroll
  for (int j = 0; j < kAhead; ++j)
    if (j < rows + 2) load_row(j);
  float2 acc[3][kCW];
#pragma unroll
  for (int j = 0; j < kRegRows + 2; ++j) {
    if (j >= rows + 2) break;
    float2 x[kLW];
#pragma unroll
    for (int i = 0; i < kLW; ++i) x[i] = unpack2<T>(raw[j % kAhead][i]);
    if (j + kAhead < rows + 2) load_row(j % kAhead);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int r = j - ky;
      if (r < 0 || r >= kRegRows) continue;
#pragma unroll
      for (int i = 0; i < kCW; ++i) {
        float2& d = acc[r % 3][i];
        d = ky == 0 ? ptx::fmul2(x[i], w[0]) : ptx::ffma2(x[i], w[ky * 3], d);
        d = ptx::ffma2(x[i + 1], w[ky * 3 + 1], d);
        d = ptx::ffma2(x[i + 2], w[ky * 3 + 2], d);
      }
    }
    if (j >= 2) {
#pragma unroll
      for (int i = 0; i < kCW; ++i)
        if (i < ncol) pout[(uint32_t)i * uopitch] = pack2<T>(acc[(j - 2) % 3][i].x, acc[(j - 2) % 3][i].y);
      pout += ors;
    }
  }
}

bool dw_reg_supported(const DwParams& p, int et) {
  return dw_strip_supported(p, et) && p.rate == 1;
}

cudaError_t launch_dw_reg(const DwParams& p, int et, cudaStream_t s) {
  const int nchunks = (p.in.C + 63) / 64, row_blocks = (p.OH + kRegRows - 1) / kRegRows;
  const bool wide = (p.OW % 8 == 0) && (p.OW % 6 != 0);      // 8-column strips where they tile the row and 6-column ones do not
  const int cw = wide ? 8 : 6, strips = (p.OW + cw - 1) / cw;
  const long long n_items = (long long)p.N * row_blocks * strips * nchunks;
  const unsigned blocks = (unsigned)((n_items + 3) / 4);
  if (et == ET_BF16) {
    if (wide) dw_reg_kernel<__nv_bfloat16, 8, 3><<<blocks, 128, 0, s>>>(p, nchunks, strips, row_blocks, n_items);
    else dw_reg_kernel<__nv_bfloat16, 6, 4><<<blocks, 128, 0, s>>>(p, nchunks, strips, row_blocks, n_items);
  } else {
    if (wide) dw_reg_kernel<__half, 8, 3><<<blocks, 128, 0, s>>>(p, nchunks, strips, row_blocks, n_items);
    else dw_reg_kernel<__half, 6, 4><<<blocks, 128, 0, s>>>(p, nchunks, strips, row_blocks, n_items);
  }
  return cudaGetLastError();
}

template <typename TI, typename TO>
static cudaError_t launch_dw_t(const DwParams& p, cudaStream_t s) {
  const int C = p.in.C;
  const bool al8 = ((C | p.in.pitch | p.in.coff | p.out.pitch | p.out.coff) & 7) == 0;
  const bool al4 = ((C | p.in.pitch | p.in.coff | p.out.pitch | p.out.coff) & 3) == 0;
  const long long px = (long long)p.N * p.OH * p.OW;
  if (sizeof(TI) == 2 && sizeof(TO) == 2 && al8) {
    long long total = px * (C / 8);
    dw3x3_kernel<TI, TO, 8><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  } else if (al4) {
    long long total = px * (C / 4);
    dw3x3_kernel<TI, TO, 4><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  } else {
    long long total = px * C;
    dw3x3_kernel<TI, TO, 1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_dw3x3(const DwParams& p, int et, cudaStream_t s) {
  if (et == ET_F32) return launch_dw_t<float, float>(p, s);
  if (et == ET_BF16)
    return p.in_f32 ? launch_dw_t<float, __nv_bfloat16>(p, s) : launch_dw_t<__nv_bfloat16, __nv_bfloat16>(p, s);
  return p.in_f32 ? launch_dw_t<float, __half>(p, s) : launch_dw_t<__half, __half>(p, s);
}

// ---------------------------------------------------------------------------------------------
// TF1 legacy bilinear resize (align_corners=False, no half-pixel centres; DMG:344, 494) with an
// optional per-channel affine + ReLU6 (the BN/ReLU6 that follows the image-level branch, DMG:345)
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) resize_kernel(const ResizeParams p, int fx, int fy) {
  // thread = (fx consecutive output pixels of fy consecutive rows, V channels); blockIdx.y = output row group, blockIdx.z = image.
  // fy = out.H / in.H when that is a power of two (the x4 decoder upsample: the 16 outputs of a source cell come from four loads).
  // fx = out.W / in.W when that is an integer (x4 decoder upsample, x2 image-level branch, x1 BN-only steps): the fx
  // outputs share their two source columns, so each source vector is fetched once per thread instead of once per output.
  const int C = p.in.C, cg = C / V;
  const int oy = blockIdx.y * fy, n = blockIdx.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int strips = p.out.W / fx;
  if (idx >= strips * cg) return;
  const int sx_i = idx / cg, g = idx - sx_i * cg;
  const float ry = (float)p.in.H / (float)p.out.H, rx = (float)p.in.W / (float)p.out.W;
  const float sy = oy * ry;
  const int y0 = (int)floorf(sy);
  const int y1 = min(y0 + 1, p.in.H - 1);
  const int ox0 = sx_i * fx;
  const int x0 = (int)floorf(ox0 * rx);
  const int x1 = min(x0 + 1, p.in.W - 1);
  const T* base = reinterpret_cast<const T*>(p.in.ptr) + (size_t)n * p.in.H * p.in.W * p.in.pitch + p.in.coff + g * V;
  float tl[V], tr[V], bl[V], br[V], sc[V], sh[V];
  VecIO<T, V>::ld(base + ((size_t)y0 * p.in.W + x0) * p.in.pitch, tl);
  VecIO<T, V>::ld(base + ((size_t)y0 * p.in.W + x1) * p.in.pitch, tr);
  VecIO<T, V>::ld(base + ((size_t)y1 * p.in.W + x0) * p.in.pitch, bl);
  VecIO<T, V>::ld(base + ((size_t)y1 * p.in.W + x1) * p.in.pitch, br);
  if (p.scale) {
#pragma unroll
    for (int j = 0; j < V; ++j) { sc[j] = p.scale[g * V + j]; sh[j] = p.shift[g * V + j]; }
  }
  T* op = reinterpret_cast<T*>(p.out.ptr) + (((size_t)n * p.out.H + oy) * p.out.W + ox0) * p.out.pitch + p.out.coff + g * V;
  for (int q = 0; q < fy; ++q) {
    const float wy = (oy + q) * ry - (float)y0;   // same y0 for the whole row group (fy is a power of two: the products are exact)
    for (int r = 0; r < fx; ++r) {
      const float sx = (ox0 + r) * rx;
      const float wx = sx - (float)x0;     // same x0 for the whole strip (fx divides the scale)
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float top = tl[j] + (tr[j] - tl[j]) * wx;
        const float bot = bl[j] + (br[j] - bl[j]) * wx;
        float v = top + (bot - top) * wy;
        if (p.scale) v = fmaf(v, sc[j], sh[j]);
        if (p.relu6) v = fminf(fmaxf(v, 0.f), 6.f);
        o[j] = v;
      }
      VecIO<T, V>::st(op + ((size_t)q * p.out.W + r) * p.out.pitch, o);
    }
  }
}

template <typename T>
static cudaError_t launch_resize_t(const ResizeParams& p, cudaStream_t s) {
  const int C = p.in.C;
  const bool al8 = ((C | p.in.pitch | p.in.coff | p.out.pitch | p.out.coff) & 7) == 0;
  const long long px = (long long)p.N * p.out.H * p.out.W;
  (void)px;
  if (p.N > 65535 || p.out.H > 65535) return cudaErrorInvalidValue;
  const int fx = (p.out.W % p.in.W == 0 && p.out.W / p.in.W <= 8) ? p.out.W / p.in.W : 1;
  const int ry = p.out.H % p.in.H == 0 ? p.out.H / p.in.H : 1;
  const int fy = (ry == 2 || ry == 4 || ry == 8) ? ry : 1;
  if (sizeof(T) == 2 && al8) {
    dim3 grid((unsigned)(((p.out.W / fx) * (C / 8) + 255) / 256), (unsigned)(p.out.H / fy), (unsigned)p.N);
    resize_kernel<T, 8><<<grid, 256, 0, s>>>(p, fx, fy);
  } else {
    dim3 grid((unsigned)(((p.out.W / fx) * (C / 4) + 255) / 256), (unsigned)(p.out.H / fy), (unsigned)p.N);
    resize_kernel<T, 4><<<grid, 256, 0, s>>>(p, fx, fy);
  }
  return cudaGetLastError();
}
cudaError_t launch_resize(const ResizeParams& p, int et, cudaStream_t s) {
  if (et == ET_F32) return launch_resize_t<float>(p, s);
  if (et == ET_BF16) return launch_resize_t<__nv_bfloat16>(p, s);
  return launch_resize_t<__half>(p, s);
}

// ---------------------------------------------------------------------------------------------
// 2x2 average pool, stride 2 (tf.nn.pool AVG SAME on even sizes, DMG:331-335)
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) avgpool_kernel(const PoolParams p) {
  const int C = p.in.C, cg = C / V;
  const long long total = (long long)p.N * p.out.H * p.out.W * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % cg);
  long long pix = idx / cg;
  const int ox = (int)(pix % p.out.W); pix /= p.out.W;
  const int oy = (int)(pix % p.out.H);
  const int n = (int)(pix / p.out.H);
  const T* base = reinterpret_cast<const T*>(p.in.ptr) + (size_t)n * p.in.H * p.in.W * p.in.pitch + p.in.coff + g * V;
  float a[V], b[V], c[V], d[V], o[V];
  VecIO<T, V>::ld(base + ((size_t)(2 * oy) * p.in.W + 2 * ox) * p.in.pitch, a);
  VecIO<T, V>::ld(base + ((size_t)(2 * oy) * p.in.W + 2 * ox + 1) * p.in.pitch, b);
  VecIO<T, V>::ld(base + ((size_t)(2 * oy + 1) * p.in.W + 2 * ox) * p.in.pitch, c);
  VecIO<T, V>::ld(base + ((size_t)(2 * oy + 1) * p.in.W + 2 * ox + 1) * p.in.pitch, d);
#pragma unroll
  for (int j = 0; j < V; ++j) o[j] = ((a[j] + b[j]) + (c[j] + d[j])) * 0.25f;
  T* op = reinterpret_cast<T*>(p.out.ptr) + (((size_t)n * p.out.H + oy) * p.out.W + ox) * p.out.pitch + p.out.coff + g * V;
  VecIO<T, V>::st(op, o);
}
template <typename T>
static cudaError_t launch_avgpool_t(const PoolParams& p, cudaStream_t s) {
  const int C = p.in.C;
  const bool al8 = ((C | p.in.pitch | p.in.coff | p.out.pitch | p.out.coff) & 7) == 0;
  const long long px = (long long)p.N * p.out.H * p.out.W;
  if (sizeof(T) == 2 && al8) {
    long long total = px * (C / 8);
    avgpool_kernel<T, 8><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  } else {
    long long total = px * (C / 4);
    avgpool_kernel<T, 4><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  }
  return cudaGetLastError();
}
cudaError_t launch_avgpool(const PoolParams& p, int et, cudaStream_t s) {
  if (et == ET_F32) return launch_avgpool_t<float>(p, s);
  if (et == ET_BF16) return launch_avgpool_t<__nv_bfloat16>(p, s);
  return launch_avgpool_t<__half>(p, s);
}

// ---------------------------------------------------------------------------------------------
// network stem: layers whose input is the 1-channel FP32 crop (cnn0 = depthwise + pointwise 1->64,
// DMG:396; residual0 = 1x1 stride-2 conv 1->128, DMG:407).  K = 1 "GEMMs" are outer products:
// one thread = (output pixel, 8 output channels); 8 consecutive lanes write 128 contiguous bytes.
// ---------------------------------------------------------------------------------------------
// Two phases per block of 256 consecutive output pixels: (1) one thread per pixel computes d once into shared
// memory; (2) the block streams the pixels x channels out as consecutive 16-byte chunks (a warp instruction
// writes 512 contiguous bytes), each thread keeping the weights / scale / shift of its fixed 8-channel group.
constexpr int kStemGroups = 4;   // 256-pixel groups per block: the 24 per-thread constants are loaded once per 1024 pixels
template <typename T>
__global__ void __launch_bounds__(256) stem_kernel(const StemParams p) {
  __shared__ float s_d[2][256];
  const int OW = p.out.W, OH = p.out.H;
  const long long npix = (long long)p.N * OH * OW;
  const int cg = p.out.C >> 3;                 // 8-channel groups per pixel (power of two: 8 or 16)
  const int g = threadIdx.x & (cg - 1);
  float ws[8], sh[8];                          // weight x folded BN scale, shift: one FFMA per output
  {
    float w[8], sc[8];
    VecIO<float, 8>::ld(p.w + g * 8, w);
    VecIO<float, 8>::ld(p.scale + g * 8, sc);
    VecIO<float, 8>::ld(p.shift + g * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) ws[j] = w[j] * sc[j];
  }
  float dwt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) dwt[t] = p.dw ? p.dw[t] : 0.f;
  const int ppi = 256 / cg;                    // pixels covered by one pass of the block
  T* obase = reinterpret_cast<T*>(p.out.ptr);
  for (int grp = 0; grp < kStemGroups; ++grp) {
    const long long pix0 = ((long long)blockIdx.x * kStemGroups + grp) * 256;
    if (pix0 >= npix) break;
    float* sd = s_d[grp & 1];
    {
      const long long pix = pix0 + threadIdx.x;
      float d = 0.f;
      if (pix < npix) {
        // 32-bit index arithmetic (the launcher guarantees npix < 2^31): three 64-bit divisions per thread were a visible part of
        // this write-bound kernel's instruction stream
        const unsigned upix = (unsigned)pix;
        const int ox = (int)(upix % (unsigned)OW);
        const unsigned r = upix / (unsigned)OW;
        const int oy = (int)(r % (unsigned)OH), n = (int)(r / (unsigned)OH);
        const float* img = p.in + (size_t)n * p.IH * p.IW;
        if (p.dw) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy - 1 + ky;
            if (iy < 0 || iy >= p.IH) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int ix = ox - 1 + kx;
              if (ix < 0 || ix >= p.IW) continue;
              d = fmaf(__ldg(img + (size_t)iy * p.IW + ix), dwt[ky * 3 + kx], d);
            }
          }
        } else {
          d = __ldg(img + (size_t)(oy * p.istride) * p.IW + ox * p.istride);
        }
        if (sizeof(T) == 2) d = to_f(from_f<T>(d));  // the GEMM A operand is 16-bit in the 16-bit modes
      }
      sd[threadIdx.x] = d;
    }
    __syncthreads();                           // two buffers: the next group's phase 1 cannot overtake this group's readers by more than one sync
    for (int k = 0; k < cg; ++k) {
      const int lp = k * ppi + (threadIdx.x / cg);
      const long long pix = pix0 + lp;
      if (pix >= npix) break;
      const float d = sd[lp];
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(d, ws[j], sh[j]);
        if (p.relu6) v = fminf(fmaxf(v, 0.f), 6.f);
        o[j] = v;
      }
      VecIO<T, 8>::st(obase + (size_t)pix * p.out.pitch + p.out.coff + g * 8, o);
    }
  }
}

cudaError_t launch_stem(const StemParams& p, int et, cudaStream_t s) {
  const long long npix = (long long)p.N * p.out.H * p.out.W;
  const int cg = p.out.C >> 3;
  if (cg < 1 || cg > 256 || (cg & (cg - 1)) || npix >= (1ll << 31)) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((npix + 256 * kStemGroups - 1) / (256 * kStemGroups));
  if (et == ET_F32) stem_kernel<float><<<grid, 256, 0, s>>>(p);
  else if (et == ET_BF16) stem_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  else stem_kernel<__half><<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// casts between the f32 I/O tensors and the activation element type
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}
template <typename T>
__global__ void uncast_kernel(const T* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = to_f(src[i]);
}
static unsigned grid_for(size_t n) { return (unsigned)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256); }
cudaError_t launch_cast(const float* src, void* dst, size_t n, int et, cudaStream_t s) {
  if (et == ET_F32) return cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s);
  if (et == ET_BF16) cast_kernel<<<grid_for(n), 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  else cast_kernel<<<grid_for(n), 256, 0, s>>>(src, reinterpret_cast<__half*>(dst), n);
  return cudaGetLastError();
}
cudaError_t launch_uncast(const void* src, float* dst, size_t n, int et, cudaStream_t s) {
  if (et == ET_F32) return cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s);
  if (et == ET_BF16) uncast_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  else uncast_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const __half*>(src), dst, n);
  return cudaGetLastError();
}

}  // namespace emd
