// Image-quality metrics of the reference's training / evaluation code on the GPU:
//   * mean squared error and the trainer's Huberised form of it (misc_py/denoiser-multi-gpu.py:772-773),
//   * mean SSIM as `tf_ssim` computes it (misc_py/denoiser-multi-gpu.py:124-167): 11 x 11 Gaussian window (sigma 1.5,
//     normalised), VALID windows, K1 = 0.01, K2 = 0.03, L = 1.
// One block = a 32 x 32 tile of SSIM-map pixels: the 42 x 42 patches of both images go to shared memory, each thread folds the
// five windowed moments of its 4 pixels in FP32 (as TensorFlow does), block sums go out in FP64 and a second kernel adds them
// in a fixed order -- the result is deterministic.  The MSE rides in the same launch (each block also sums its share of
// squared differences over the WHOLE image, grid-stride).
#include <cuda_runtime.h>

#include "emd_kernels.h"

namespace emd {
namespace {

constexpr int kWin = 11, kTile = 32, kPatch = kTile + kWin - 1;   // 42
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

__constant__ float c_gauss[kWin * kWin];

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];   // fixed order
  return t;
}

// grid (tiles_x, tiles_y, n); partial[(img * blocks + block) * 2 + {0, 1}] = sum of SSIM-map values, sum of squared differences
__global__ void __launch_bounds__(256) quality_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                      double* __restrict__ partial) {
  __shared__ float sa[kPatch][kPatch + 1], sb[kPatch][kPatch + 1];
  __shared__ double red[8];
  const int img = blockIdx.z, blocks = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
  const float* pa = a + (size_t)img * H * W;
  const float* pb = b + (size_t)img * H * W;
  const int oh = H - kWin + 1, ow = W - kWin + 1;
  const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
  for (int i = threadIdx.x; i < kPatch * kPatch; i += 256) {
    const int py = i / kPatch, px = i - py * kPatch;
    const int y = y0 + py, x = x0 + px;
    const bool in = y < H && x < W;
    sa[py][px] = in ? pa[(size_t)y * W + x] : 0.f;
    sb[py][px] = in ? pb[(size_t)y * W + x] : 0.f;
  }
  __syncthreads();
  double ssim_sum = 0.0;
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  for (int r = 0; r < 4; ++r) {
    const int ty = ty0 + 8 * r;
    if (y0 + ty < oh && x0 + tx < ow) {
      float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
      for (int ky = 0; ky < kWin; ++ky)
#pragma unroll
        for (int kx = 0; kx < kWin; ++kx) {
          const float w = c_gauss[ky * kWin + kx], x = sa[ty + ky][tx + kx], y = sb[ty + ky][tx + kx];
          mu1 = fmaf(w, x, mu1); mu2 = fmaf(w, y, mu2);
          s11 = fmaf(w, x * x, s11); s22 = fmaf(w, y * y, s22); s12 = fmaf(w, x * y, s12);
        }
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float v1 = s11 - mu1_sq, v2 = s22 - mu2_sq, cov = s12 - mu12;
      ssim_sum += (double)(((2.f * mu12 + kC1) * (2.f * cov + kC2)) / ((mu1_sq + mu2_sq + kC1) * (v1 + v2 + kC2)));
    }
  }
  double sq = 0.0;
  const size_t n = (size_t)H * W;
  for (size_t i = (size_t)blk * 256 + threadIdx.x; i < n; i += (size_t)blocks * 256) {
    const float d = pa[i] - pb[i];
    sq += (double)(d * d);
  }
  const double t0 = block_sum(ssim_sum, red);
  const double t1 = block_sum(sq, red);
  if (threadIdx.x == 0) {
    partial[((size_t)img * blocks + blk) * 2 + 0] = t0;
    partial[((size_t)img * blocks + blk) * 2 + 1] = t1;
  }
}

// one warp per image: out[3 * img + {0, 1, 2}] = MSE, Huberised loss, mean SSIM
__global__ void quality_final_kernel(const double* __restrict__ partial, int blocks, int H, int W, double* __restrict__ out) {
  const int img = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 32) {
    s += partial[((size_t)img * blocks + i) * 2 + 0];
    q += partial[((size_t)img * blocks + i) * 2 + 1];
  }
  for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); q += __shfl_down_sync(0xffffffffu, q, o); }
  if (threadIdx.x == 0) {
    const double mse = q / ((double)H * W);
    out[3 * img + 0] = mse;
    out[3 * img + 1] = mse < 0.001 ? 1000.0 * mse : sqrt(1000.0 * mse);   // DMG:773
    out[3 * img + 2] = s / ((double)(H - kWin + 1) * (W - kWin + 1));
  }
}

}  // namespace

size_t quality_partial_bytes(int n, int H, int W) {
  const int tx = (W - kWin + 1 + kTile - 1) / kTile, ty = (H - kWin + 1 + kTile - 1) / kTile;
  return (size_t)n * tx * ty * 2 * sizeof(double);
}

cudaError_t launch_quality(const float* a, const float* b, int n, int H, int W, double* d_partial, double* d_out, cudaStream_t s) {
  static thread_local int window_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (window_dev != dev) {   // _tf_fspecial_gauss(11, 1.5): exp(-(x^2 + y^2) / (2 sigma^2)) in FP32, normalised by its sum
    float g[kWin * kWin], sum = 0.f;
    for (int y = 0; y < kWin; ++y)
      for (int x = 0; x < kWin; ++x) {
        const float fx = (float)(x - kWin / 2), fy = (float)(y - kWin / 2);
        g[y * kWin + x] = expf(-((fx * fx + fy * fy) / (2.0f * 1.5f * 1.5f)));
        sum += g[y * kWin + x];
      }
    for (float& v : g) v /= sum;
    cudaError_t r = cudaMemcpyToSymbolAsync(c_gauss, g, sizeof g, 0, cudaMemcpyHostToDevice, s);
    if (r != cudaSuccess) return r;
    r = cudaStreamSynchronize(s);   // g is a stack array
    if (r != cudaSuccess) return r;
    window_dev = dev;
  }
  const int tx = (W - kWin + 1 + kTile - 1) / kTile, ty = (H - kWin + 1 + kTile - 1) / kTile;
  quality_kernel<<<dim3(tx, ty, n), 256, 0, s>>>(a, b, H, W, d_partial);
  quality_final_kernel<<<n, 32, 0, s>>>(d_partial, tx * ty, H, W, d_out);
  return cudaGetLastError();
}

}  // namespace emd
