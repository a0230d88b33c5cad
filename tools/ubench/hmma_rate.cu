// Microbenchmark: mma.sync.m16n8k16 bf16 (legacy HMMA path) and ldmatrix.x4 issue throughput per SM on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) k_hmma(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b[2] = {0x3c003c00u, 0};
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512) k_ldsm(float* out, long long* cyc, int iters) {
  __shared__ __align__(128) uint8_t sm[32768];
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  __syncthreads();
  uint32_t acc = 0;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 128 + (threadIdx.x >> 5) * 16;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t r0, r1, r2, r3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + ((i * 4096 + it * 128) & 16383)));
      acc += r0 ^ r1 ^ r2 ^ r3;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int which = 0; which < 2; ++which)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (which == 0) k_hmma<<<148, warps * 32>>>(out, cyc, iters); else k_ldsm<<<148, warps * 32>>>(out, cyc, iters);
      }
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double n = (double)iters * 8 * warps;
      printf("%s warps/SM %2d: %.3f instr/cycle/SM (%.2f cycles per instr per SM)\n", which ? "LDSM.x4      " : "HMMA m16n8k16", warps, n / avg, avg / n);
    }
  return 0;
}
