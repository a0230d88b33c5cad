// Microbenchmark: issue rate of tcgen05.mma (M=128, K=16, kind::f16) as a function of N, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(long long* cyc, int N, int count, int distinct_cols) {
  extern __shared__ uint8_t raw[];
  const uint32_t r32 = smem_u32(raw);
  const uint32_t base = (r32 + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - r32);
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3c003c00u, 0, 0x3c003c00u, 0);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint64_t adesc = (uint64_t)((base & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t bs = base + 16384;
    const uint64_t bdesc = (uint64_t)((bs & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      const uint32_t dstep = distinct_cols ? 64u : 0u;
      for (int i = 0; i < count; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + dstep * u), "l"(adesc + (uint64_t)(u * 2)),
                       "l"(bdesc + (uint64_t)(u * 2)), "r"(idesc)
                       : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      long long t1 = clock64();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"((uint32_t)(rep & 1)) : "memory");
      long long t2 = clock64();
      cyc[rep * 2] = t1 - t0; cyc[rep * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int dc : {0, 1})
    for (int N : {16, 32, 64, 128, 256}) {
      const int count = 256;
      k<<<1, 128, smem>>>(d, N, count, dc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      long long h[6]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      printf("N=%3d %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor N/2 = %d)\n", N, dc ? "distinct D cols" : "same D cols    ",
             (double)h[4] / count, (double)h[5] / count, N / 2);
    }
  return 0;
}
