// Experiment: can a tcgen05 A-operand descriptor address a SHIFTED 128-row view of a 128B-swizzled pixel buffer?
// Buffer: HR rows x P pixels x 64 bf16 channels (128 B per pixel), written with the address-based 128B swizzle
// (16-byte chunk index ^= (byte_address >> 7) & 7, what TMA SWIZZLE_128B produces for a 1024-aligned box).
// View for tap (ky,kx): M row r = y*8 + x (16 x 8 tile) -> pixel (y+ky, x+kx): start = base + (ky*P + kx)*128,
// SBO = P*128.  Variants: P in {10,16}, descriptor base_offset in {0, (start>>7)&7}.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* halo /*[HR][P][64]*/, const __nv_bfloat16* wB /*[16][64]*/,
                                            float* out /*[9][128][16]*/, int P, int HR, int use_bo) {
  extern __shared__ uint8_t raw[];
  const uint32_t r32 = smem_u32(raw);
  const uint32_t base = (r32 + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - r32);
  uint8_t* sHalo = sm;                                  // HR*P*128 bytes
  const uint32_t halo_bytes = ((uint32_t)HR * P * 128 + 1023u) & ~1023u;
  uint8_t* sBm = sm + halo_bytes;                       // 16 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBm + 2048);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  // fill halo with address-based swizzle
  for (int i = tid; i < HR * P * 8; i += 128) {
    const int pix = i >> 3, ck = i & 7;
    const uint32_t addr = base + (uint32_t)pix * 128u;
    const int pc = ck ^ ((addr >> 7) & 7);
    *reinterpret_cast<uint4*>(sHalo + (size_t)pix * 128 + pc * 16) = *reinterpret_cast<const uint4*>(halo + (size_t)pix * 64 + ck * 8);
  }
  for (int i = tid; i < 16 * 8; i += 128) {
    const int row = i >> 3, ck = i & 7;
    *reinterpret_cast<uint4*>(sBm + row * 128 + ((ck ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(wB + row * 64 + ck * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    if (tid == 0) {
      const uint32_t start = base + (uint32_t)(ky * P + kx) * 128u;
      const uint32_t sbo = (uint32_t)P * 128u;
      const uint64_t bo = use_bo ? ((start >> 7) & 7u) : 0u;
      const uint64_t adesc = (uint64_t)((start & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
                             (bo << 49) | ((uint64_t)2 << 61);
      const uint32_t bs = smem_u32(sBm);
      const uint64_t bdesc = (uint64_t)((bs & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                             ((uint64_t)2 << 61);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t accum = ks ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(adesc + (uint64_t)(ks * 2)), "l"(bdesc + (uint64_t)(ks * 2)), "r"(idesc), "r"(accum)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
      }
    }
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[((size_t)tap * 128 + tid) * 16 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  const int HR = 18;
  for (int P : {10, 16}) {
    for (int use_bo : {0, 1}) {
      std::vector<__nv_bfloat16> halo((size_t)HR * P * 64), wB(16 * 64);
      std::vector<float> hf(halo.size()), wf(wB.size());
      srand(1);
      for (size_t i = 0; i < halo.size(); ++i) { hf[i] = (float)((rand() % 17) - 8) / 8.f; halo[i] = __float2bfloat16(hf[i]); }
      for (size_t i = 0; i < wB.size(); ++i) { wf[i] = (float)((rand() % 9) - 4) / 4.f; wB[i] = __float2bfloat16(wf[i]); }
      __nv_bfloat16 *dh, *dw; float* dout;
      cudaMalloc(&dh, halo.size() * 2); cudaMalloc(&dw, wB.size() * 2); cudaMalloc(&dout, 9 * 128 * 16 * 4);
      cudaMemcpy(dh, halo.data(), halo.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dw, wB.data(), wB.size() * 2, cudaMemcpyHostToDevice);
      cudaMemset(dout, 0, 9 * 128 * 16 * 4);
      const int smem = 1024 + HR * P * 128 + 1024 + 2048 + 64;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k<<<1, 128, smem>>>(dh, dw, dout, P, HR, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("P=%d bo=%d: CUDA error %s\n", P, use_bo, cudaGetErrorString(e)); return 1; }
      std::vector<float> out(9 * 128 * 16);
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      printf("P=%2d base_offset=%s:", P, use_bo ? "auto" : "0   ");
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        double maxerr = 0;
        for (int r = 0; r < 128; ++r) {
          const int y = r / 8, x = r % 8;
          const size_t pix = (size_t)(y + ky) * P + x + kx;
          for (int n = 0; n < 16; ++n) {
            double ref = 0;
            for (int c = 0; c < 64; ++c) ref += (double)hf[pix * 64 + c] * wf[n * 64 + c];
            maxerr = fmax(maxerr, fabs(ref - out[((size_t)tap * 128 + r) * 16 + n]));
          }
        }
        printf(" t%d:%s", tap, maxerr < 1e-3 ? "OK" : "BAD");
      }
      printf("\n");
      cudaFree(dh); cudaFree(dw); cudaFree(dout);
    }
  }
  return 0;
}
