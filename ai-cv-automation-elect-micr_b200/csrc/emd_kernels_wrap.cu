// emd_kernels_wrap.cu -- whole-image wrapper kernels (all HBM-bound, CUDA cores):
//   normalise  = preprocess (DMG:853-858) + scale0to1 (DEN:684-695 == DMG:817-828)
//   gather     = the crop slicing of Denoiser.denoise (DEN:671-673)
//   stitch     = accumulate / contributions / divide / clip (DEN:658-659, 671-680; App. D-4 repair)
#include "emd_kernels.h"
#include <math.h>

namespace emd {

static constexpr int kMinMaxBlocks = 148 * 4;
size_t minmax_partial_bytes() { return sizeof(double) * 2 * kMinMaxBlocks; }

template <typename T>
__device__ __forceinline__ T fix_nonfinite(T v) {  // DMG:855-856: NaN -> 0.5, Inf -> 0.5
  return (isnan(v) || isinf(v)) ? (T)0.5 : v;
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ img, size_t n, double* __restrict__ partial) {
  T mn = (T)INFINITY, mx = (T)-INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const T v = fix_nonfinite(img[i]);
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  double dmn = (double)mn, dmx = (double)mx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmn = fmin(dmn, __shfl_xor_sync(0xffffffffu, dmn, o));
    dmx = fmax(dmx, __shfl_xor_sync(0xffffffffu, dmx, o));
  }
  __shared__ double smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = dmn; smx[threadIdx.x >> 5] = dmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { dmn = fmin(dmn, smn[w]); dmx = fmax(dmx, smx[w]); }
    partial[2 * blockIdx.x] = dmn;
    partial[2 * blockIdx.x + 1] = dmx;
  }
}

__global__ void minmax_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < nblocks; i += 32) { mn = fmin(mn, partial[2 * i]); mx = fmax(mx, partial[2 * i + 1]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
}

cudaError_t launch_minmax(const void* img, int in_f64, size_t n, double* d_minmax, void* d_partial, cudaStream_t s) {
  int blocks = (int)((n + 255) / 256 < (size_t)kMinMaxBlocks ? (n + 255) / 256 : kMinMaxBlocks);
  if (blocks < 1) blocks = 1;
  double* partial = reinterpret_cast<double*>(d_partial);
  if (in_f64) minmax_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<const double*>(img), n, partial);
  else minmax_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(img), n, partial);
  minmax_final_kernel<<<1, 32, 0, s>>>(partial, blocks, d_minmax);
  return cudaGetLastError();
}

// (x - min) / (max - min) in the input precision with IEEE round-to-nearest sub/div (App. A.10)
__global__ void __launch_bounds__(256) normalise_f32_kernel(const float* __restrict__ img, size_t n,
                                                            const double* __restrict__ mm, float* __restrict__ out) {
  const float mn = (float)mm[0], mx = (float)mm[1];
  const bool constant = (mn == mx);
  const float den = __fsub_rn(mx, mn);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fix_nonfinite(img[i]);
    out[i] = constant ? 0.5f : __fdiv_rn(__fsub_rn(v, mn), den);
  }
}
__global__ void __launch_bounds__(256) normalise_f64_kernel(const double* __restrict__ img, size_t n,
                                                            const double* __restrict__ mm, float* __restrict__ out) {
  const double mn = mm[0], mx = mm[1];
  const bool constant = (mn == mx);
  const double den = __dsub_rn(mx, mn);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = fix_nonfinite(img[i]);
    out[i] = constant ? 0.5f : __double2float_rn(__ddiv_rn(__dsub_rn(v, mn), den));
  }
}

cudaError_t launch_normalise_apply(const void* img, int in_f64, size_t n, const double* d_minmax, float* out,
                                   cudaStream_t s) {
  unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)(148 * 8) ? (n + 255) / 256 : 148 * 8);
  if (blocks < 1) blocks = 1;
  if (in_f64) normalise_f64_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const double*>(img), n, d_minmax, out);
  else normalise_f32_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(img), n, d_minmax, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Denoiser.preprocess (DEN:632-643), the single-crop path: cv2.resize to S x S (INTER_LINEAR: half-pixel centres, edge
// replicate, float coefficients, horizontal pass then vertical), scale0to1, NaN -> 0.5, Inf -> 0.5, scale0to1.  The FIRST
// min-max runs before the NaN/Inf replacement: numpy's min/max propagate NaN, so one NaN flattens the crop.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_linear_kernel(const float* __restrict__ img, int H, int W, int S, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int oy = idx / S, ox = idx - oy * S;
  if (H == S && W == S) { out[idx] = img[idx]; return; }            // cv2 copies when the size does not change
  const double fy = ((double)oy + 0.5) * ((double)H / S) - 0.5, fx = ((double)ox + 0.5) * ((double)W / S) - 0.5;
  const int y0 = (int)floor(fy), x0 = (int)floor(fx);
  const float wy = (float)(fy - y0), wx = (float)(fx - x0);
  const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1), x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
  const float a0 = __fsub_rn(1.0f, wx), b0 = __fsub_rn(1.0f, wy);
  const float r0 = __fadd_rn(__fmul_rn(img[(size_t)y0c * W + x0c], a0), __fmul_rn(img[(size_t)y0c * W + x1c], wx));
  const float r1 = __fadd_rn(__fmul_rn(img[(size_t)y1c * W + x0c], a0), __fmul_rn(img[(size_t)y1c * W + x1c], wx));
  out[idx] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, wy));
}

// numpy-style min / max of a small array (one block): result NaN if any element is NaN.  mm = {min, max}
__global__ void __launch_bounds__(1024) minmax_numpy_kernel(const float* __restrict__ x, int n, float* __restrict__ mm) {
  float mn = INFINITY, mx = -INFINITY;
  int nan = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i];
    if (isnan(v)) nan = 1;
    else { mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  __shared__ float smn[32], smx[32];
  __shared__ int snan;
  if (threadIdx.x == 0) snan = 0;
  __syncthreads();
  if (nan) atomicOr(&snan, 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    mm[0] = snan ? NAN : mn;
    mm[1] = snan ? NAN : mx;
  }
}

// scale0to1 as numpy computes it on float32 (NaN / Inf propagate through the IEEE sub / div), optionally followed by the
// NaN -> 0.5, Inf -> 0.5 replacement of DEN:638-639
__global__ void __launch_bounds__(256) scale0to1_kernel(const float* __restrict__ x, int n, const float* __restrict__ mm, int fix, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mn = mm[0], mx = mm[1];
  float v = (mn == mx) ? 0.5f : __fdiv_rn(__fsub_rn(x[i], mn), __fsub_rn(mx, mn));     // NaN == NaN is false: falls to the division, like numpy
  if (fix && (isnan(v) || isinf(v))) v = 0.5f;
  out[i] = v;
}

cudaError_t launch_preprocess_crop(const float* d_img, int H, int W, int S, float* d_tmp, float* d_mm, float* d_out, cudaStream_t s) {
  const int n = S * S, blocks = (n + 255) / 256;
  resize_linear_kernel<<<blocks, 256, 0, s>>>(d_img, H, W, S, d_out);
  minmax_numpy_kernel<<<1, 1024, 0, s>>>(d_out, n, d_mm);
  scale0to1_kernel<<<blocks, 256, 0, s>>>(d_out, n, d_mm, 1, d_tmp);
  minmax_numpy_kernel<<<1, 1024, 0, s>>>(d_tmp, n, d_mm + 2);
  scale0to1_kernel<<<blocks, 256, 0, s>>>(d_tmp, n, d_mm + 2, 0, d_out);
  return cudaGetLastError();
}

// crops[(i*nx+j), r, c] = img[ys[i]+r, xs[j]+c]; one thread per 4 output pixels (crop % 4 == 0)
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ img, int H, int W,
                                                     const int* __restrict__ ys, const int* __restrict__ xs, int ny,
                                                     int nx, int crop, float* __restrict__ crops) {
  const int q = crop >> 2;
  const size_t total = (size_t)ny * nx * crop * q;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % q);
    size_t t = idx / q;
    const int r = (int)(t % crop); t /= crop;
    const int j = (int)(t % nx), i = (int)(t / nx);
    const float* src = img + (size_t)(ys[i] + r) * W + xs[j] + c4 * 4;
    float4 v = make_float4(src[0], src[1], src[2], src[3]);
    reinterpret_cast<float4*>(crops)[idx] = v;
  }
}
cudaError_t launch_gather(const float* img, int H, int W, const int* d_ys, const int* d_xs, int ny, int nx, int crop,
                          float* crops, cudaStream_t s) {
  const size_t total = (size_t)ny * nx * crop * (crop / 4);
  unsigned blocks = (unsigned)((total + 255) / 256 < (size_t)(148 * 16) ? (total + 255) / 256 : 148 * 16);
  gather_kernel<<<blocks, 256, 0, s>>>(img, H, W, d_ys, d_xs, ny, nx, crop, crops);
  return cudaGetLastError();
}

// gather-form overlap average: each output pixel sums its covering tiles in the reference's loop
// order (i outer, j inner; DEN:666-675) in float64 and divides by the count -- deterministic.
template <typename TO>
__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ tiles, const int* __restrict__ ys,
                                                     const int* __restrict__ xs, int ny, int nx, int crop, int H, int W,
                                                     int clip, TO* __restrict__ out) {
  const size_t total = (size_t)H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % W), r = (int)(idx / W);
    double sum = 0.0, cnt = 0.0;
    for (int i = 0; i < ny; ++i) {
      const int dy = r - ys[i];
      if (dy < 0 || dy >= crop) continue;
      for (int j = 0; j < nx; ++j) {
        const int dx = c - xs[j];
        if (dx < 0 || dx >= crop) continue;
        sum += (double)tiles[((size_t)(i * nx + j) * crop + dy) * crop + dx];
        cnt += 1.0;
      }
    }
    double v = sum / cnt;
    if (clip) v = fmin(fmax(v, 0.0), 1.0);
    out[idx] = (TO)v;     // float output: the float64 average rounded once (round to nearest)
  }
}
cudaError_t launch_stitch(const float* tiles, const int* d_ys, const int* d_xs, int ny, int nx, int crop, int H, int W,
                          int clip, void* out, int out_f32, cudaStream_t s) {
  const size_t total = (size_t)H * W;
  unsigned blocks = (unsigned)((total + 255) / 256 < (size_t)(148 * 16) ? (total + 255) / 256 : 148 * 16);
  if (out_f32) stitch_kernel<float><<<blocks, 256, 0, s>>>(tiles, d_ys, d_xs, ny, nx, crop, H, W, clip, reinterpret_cast<float*>(out));
  else stitch_kernel<double><<<blocks, 256, 0, s>>>(tiles, d_ys, d_xs, ny, nx, crop, H, W, clip, reinterpret_cast<double*>(out));
  return cudaGetLastError();
}

}  // namespace emd
