// emd_tma.h -- host-side tensor-map encoding and the device-side PTX wrappers shared by the
// TMA-fed kernels (emd_umma.cu, emd_dw.cu).  libcuda is not linked: the encoder is fetched through
// cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace emd {

// unsigned division by a runtime constant (Granlund-Montgomery): q = (t + ((x - t) >> 1)) >> (l - 1), t = mulhi(m, x)
struct FastDiv { uint32_t d, m, l; };
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d; f.l = 0;
  while ((1ull << f.l) < d) ++f.l;
  f.m = (uint32_t)((((1ull << f.l) - d) << 32) / d + 1);
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) {
  if (f.d == 1) return x;
  const uint32_t t = __umulhi(f.m, x);
  return (t + ((x - t) >> 1)) >> (f.l - 1);
}
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tma_encoder() {
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

// 4-D map over an NHWC 16-bit activation view: dims (C, W, H, N); box (bc, bw, bh, 1); element strides
// (1, es, es, 1); zero fill out of range.  base = first channel of the view (16-byte aligned).
inline bool tma_encode_nhwc(CUtensorMap* map, bool bf16, void* base, int C, int W, int H, int N, int pitch, int bc, int bw,
                            int bh, int es, bool swizzle128, int bn = 1) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  memset(map, 0, sizeof *map);
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Same, with explicit strides in ELEMENTS between consecutive x, y and n indices (a sub-pixel phase of a
// transposed convolution's output is the view with sx = 2*pitch, sy = 2*W*pitch).
inline bool tma_encode_view(CUtensorMap* map, bool bf16, void* base, int C, int W, int H, int N, size_t sx, size_t sy, size_t sn,
                            int bc, int bw, int bh, bool swizzle128, int bn = 1) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)sn * 2};
  const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  memset(map, 0, sizeof *map);
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D map over the packed weight image, which is linear in memory: viewed as rows of 256 16-bit elements (512 B) so a
// block of `box_bytes` is box_bytes/512 row requests (TMA issues one request per box row; 128-byte rows would be 4x as many)
inline bool tma_encode_linear512(CUtensorMap* map, bool bf16, const void* base, size_t total_bytes, int box_bytes) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || (box_bytes & 511)) return false;
  const cuuint64_t dims[2] = {256, (cuuint64_t)((total_bytes + 511) / 512)};
  const cuuint64_t strides[1] = {512};
  const cuuint32_t box[2] = {256, (cuuint32_t)(box_bytes / 512)};
  const cuuint32_t estr[2] = {1, 1};
  memset(map, 0, sizeof *map);
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory accesses by 32-bit shared address (the compiler keeps base + immediate forms)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
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
  // (a stated suspend-time limit -- try_wait's 4th operand -- was measured: ptxas turns it into a PHASECHK + NANOSLEEP loop that
  // polls as often and executes more instructions; the plain form stays)
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
// non-blocking probe of a phase (true = that phase has completed)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// one potentially-blocking probe: the thread may be suspended by the hardware for up to ~ns nanoseconds
__device__ __forceinline__ bool mbar_try_wait_ns(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 4-D tiled TMA load (c, x, y, n); out-of-range coordinates are zero-filled = TF SAME padding
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst), "l"(map), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// 4-D tiled TMA store smem -> global (bulk async-group completion); out-of-range elements are not written
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c, int x, int y, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// ---- CTA pair (cta_group::2): both CTAs' loads signal the mbarrier of the even (leader) CTA of the pair ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader's copy
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst), "l"(map), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                   "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & kPeerBitMask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- tcgen05 / TMEM ----
// Programmatic dependent launch (a kernel launched with the programmatic-stream-serialization attribute may become resident
// while the kernel before it in the stream is still running): `griddep_launch` lets the NEXT kernel's CTAs start taking free
// SMs; `griddep_wait` blocks until the PREVIOUS kernel has completed and its writes are visible -- everything that reads or
// writes activations comes after it, only the prologue (barriers, TMEM, tensor maps, constant weights) before.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// one lane of a converged warp (elect.sync): the form the compiler recognises as a single-thread region of a uniform warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {   // same warp id, same dst offset in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// completion of the pair's MMAs arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// M = 256 across the CTA pair: each CTA supplies its 128 rows of A and half of B's N rows, issued by the leader only
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// One K block (up to four K=16 steps) in a single PTX block: the descriptors' high words (SBO 1024 B, version 1,
// SWIZZLE_128B) are constants, the low words advance by 2 (32 bytes >> 4) per step.  Keeping the address arithmetic out
// of the issuing thread's instruction stream matters: one thread issues every MMA of the CTA, and a loop of a dozen
// dependent instructions per tcgen05.mma makes the kernel issue-bound (measured ~250 cycles per MMA before).
template <bool kPair, int kSteps>
__device__ __forceinline__ void umma_kblock(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum_first) {
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // bits 32-45 SBO, 46 version, 61-63 layout
  static_assert(kSteps >= 1 && kSteps <= 4, "K block is at most 64 channels");
#define EMD_MMA(G, STEP, PRED)                                                                                   \
  "add.u32 al, %1, " #STEP ";\n\tadd.u32 bl, %2, " #STEP ";\n\tmov.b64 da, {al, %4};\n\tmov.b64 db, {bl, %4};\n\t" \
  "tcgen05.mma.cta_group::" #G ".kind::f16 [%0], da, db, %3, " PRED ";\n\t"
  if constexpr (!kPair) {
    if constexpr (kSteps == 4)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(1, 0, "p")
                       EMD_MMA(1, 2, "1") EMD_MMA(1, 4, "1") EMD_MMA(1, 6, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else if constexpr (kSteps == 3)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(1, 0, "p")
                       EMD_MMA(1, 2, "1") EMD_MMA(1, 4, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else if constexpr (kSteps == 2)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(1, 0, "p")
                       EMD_MMA(1, 2, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(1, 0, "p") "}" ::"r"(
                       d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
  } else {
    if constexpr (kSteps == 4)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(2, 0, "p")
                       EMD_MMA(2, 2, "1") EMD_MMA(2, 4, "1") EMD_MMA(2, 6, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else if constexpr (kSteps == 3)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(2, 0, "p")
                       EMD_MMA(2, 2, "1") EMD_MMA(2, 4, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else if constexpr (kSteps == 2)
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(2, 0, "p")
                       EMD_MMA(2, 2, "1") "}" ::"r"(d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\tsetp.ne.b32 p, %5, 0;\n\t" EMD_MMA(2, 0, "p") "}" ::"r"(
                       d_tmem),
                   "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kHi), "r"(accum_first)
                   : "memory");
  }
#undef EMD_MMA
}
// low word of a K-major SWIZZLE_128B descriptor: start address >> 4 (14 bits) | LBO field = 1
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
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
// 32 lanes x 32 consecutive 32-bit columns: thread (lane) gets its row's columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, 128-byte swizzle shared-memory matrix descriptor: start address >> 4 | LBO (ignored for
// swizzled K-major) | SBO = 1024 B between 8-row groups | version 1 (sm_100) | layout 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// packed FP32 pair FMA (sm_100 FFMA2): d = a * b + c element-wise on {x, y}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
}  // namespace ptx
#endif

}  // namespace emd
