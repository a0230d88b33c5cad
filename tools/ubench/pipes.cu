// Microbenchmark: issue throughput of the CUDA-core ops the depthwise producer uses (per SM, warp-instr / cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, long long* cycles, int iters, float seed) {
  float a[8], b[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; u[i] = __float_as_uint(a[i]); }
  const float w0 = seed * 1.0001f, w1 = seed * 0.9999f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (OP == 0) {  // scalar FFMA, 3 register operands, 8 independent chains
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], w0, b[i]);
      } else if (OP == 1) {  // FFMA2: 4 independent pair chains
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          unsigned long long x = ((unsigned long long)__float_as_uint(a[i + 1]) << 32) | __float_as_uint(a[i]);
          unsigned long long w = ((unsigned long long)__float_as_uint(w1) << 32) | __float_as_uint(w0);
          unsigned long long c = ((unsigned long long)__float_as_uint(b[i + 1]) << 32) | __float_as_uint(b[i]);
          unsigned long long d;
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(w), "l"(c));
          a[i] = __uint_as_float((uint32_t)d); a[i + 1] = __uint_as_float((uint32_t)(d >> 32));
        }
      } else if (OP == 2) {  // HFMA2.BF16
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(__float_as_uint(w0)), "r"(__float_as_uint(b[i])));
      } else if (OP == 3) {  // shift left 16 (compiler picks the op)
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = (u[i] << 16) ^ u[(i + 1) & 7];
      } else if (OP == 4) {  // PRMT
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __byte_perm(u[i], u[(i + 1) & 7], 0x1044);
      } else if (OP == 5) {  // LOP3 and-mask
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = (u[i] & 0xffff0000u) | u[(i + 3) & 7];
      } else if (OP == 7) {  // FHFMA.BF16: fp32 += bf16 * bf16 (mixed-precision fma, PTX 8.6, sm_100+), low halves
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned short xl, xh, wl, wh;
          asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(u[i]));
          asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(u[(i + 1) & 7]));
          asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(xl), "h"(wl));
        }
      } else if (OP == 8) {  // FHFMA.BF16 alternating low / high halves (the depthwise pattern)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned short xl, xh, wl, wh;
          asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(u[i & 3]));
          asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(u[4 + (i & 3)]));
          if (i & 4) asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(xh), "h"(wh));
          else asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(xl), "h"(wl));
        }
      } else if (OP == 9 || OP == 10 || OP == 11) {  // mixes: 4 FFMA2 + 4 of {FHFMA (9), LOP3 (10), IMAD.U32 shift (11)}
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          unsigned long long x = ((unsigned long long)__float_as_uint(a[i + 1]) << 32) | __float_as_uint(a[i]);
          unsigned long long w = ((unsigned long long)__float_as_uint(w1) << 32) | __float_as_uint(w0);
          unsigned long long c = ((unsigned long long)__float_as_uint(b[i + 1]) << 32) | __float_as_uint(b[i]);
          unsigned long long d;
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(w), "l"(c));
          a[i] = __uint_as_float((uint32_t)d); a[i + 1] = __uint_as_float((uint32_t)(d >> 32));
          if (OP == 9) {
            unsigned short xl, xh, wl, wh;
            asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(u[i]));
            asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(u[i + 1]));
            asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(b[i]) : "h"(xl), "h"(wh));
          } else if (OP == 10) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[i + 1]), "r"(0xffff0000u));
          } else {
            asm volatile("mul.lo.u32 %0, %1, 65536;" : "=r"(u[i]) : "r"(u[i + 1]));
          }
        }
      } else if (OP == 12 || OP == 13) {  // 4 FFMA2 + 4 of {cvt.f32.f16 (HADD2.F32) (12), shift + mask = integer FP16 unpack (13)}
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          unsigned long long x = ((unsigned long long)__float_as_uint(a[i + 1]) << 32) | __float_as_uint(a[i]);
          unsigned long long w = ((unsigned long long)__float_as_uint(w1) << 32) | __float_as_uint(w0);
          unsigned long long c = ((unsigned long long)__float_as_uint(b[i + 1]) << 32) | __float_as_uint(b[i]);
          unsigned long long d;
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(w), "l"(c));
          a[i] = __uint_as_float((uint32_t)d); a[i + 1] = __uint_as_float((uint32_t)(d >> 32));
          if (OP == 12) {
            unsigned short xl, xh;
            asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(u[i]));
            asm volatile("cvt.f32.f16 %0, %1;" : "=f"(b[i]) : "h"(xl));
          } else {
            uint32_t t;
            asm volatile("shr.u32 %0, %1, 3;" : "=r"(t) : "r"(u[i + 1]));
            asm volatile("and.b32 %0, %1, 0x0FFFE000;" : "=r"(u[i]) : "r"(t));
          }
        }
      } else if (OP == 14) {  // cvt.f32.f16 alone
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned short xl, xh;
          asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(u[i]));
          asm volatile("cvt.f32.f16 %0, %1;" : "=f"(a[i]) : "h"(i & 1 ? xh : xl));
          u[i] += __float_as_uint(a[i]);
        }
      } else if (OP == 6) {  // HFMA2 fp16
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(__float_as_uint(w0)), "r"(__float_as_uint(b[i])));
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    k<OP><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double winst = (double)iters * ops_per_iter * warps;
    printf("%-14s warps/SM %2d: %.3f warp-instr/cycle/SM (%.2f cycles per warp-instr per SMSP)\n", name, warps, winst / avg,
           avg / (winst / 4));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 64);
  run<1>("FFMA2", 32);
  run<7>("FHFMA.BF16", 64);
  run<8>("FHFMA.BF16 l/h", 64);
  run<9>("4FFMA2+4FHFMA", 64);
  run<10>("4FFMA2+4LOP3", 64);
  run<11>("4FFMA2+4IMAD", 64);
  run<12>("4FFMA2+4CVT.F16", 64);
  run<13>("4FFMA2+4(SHR,AND)", 96);
  run<14>("CVT.F32.F16+IADD", 128);
  run<2>("HFMA2.BF16", 64);
  run<6>("HFMA2.F16", 64);
  run<3>("SHL16+XOR", 64);
  run<4>("PRMT", 64);
  run<5>("LOP3", 64);
  return 0;
}
