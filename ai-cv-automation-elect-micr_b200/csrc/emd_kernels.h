// emd_kernels.h -- parameter blocks and launchers shared by the engine and the kernel files.
// All activations are NHWC; "pitch" is the channel count of the *allocated* tensor (a layer may
// read/write a channel slice [coff, coff+C) of a wider tensor -- that is how concats are formed
// without a copy, DMG:348-350 / 497-499 / 509-511).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace emd {

enum ElemType { ET_F32 = 0, ET_BF16 = 1, ET_F16 = 2 };
static inline size_t elem_size(int et) { return et == ET_F32 ? 4 : 2; }

// View of an NHWC activation (possibly a channel slice of a wider tensor).
struct View {
  void* ptr;      // base of the allocated tensor for image 0
  int H, W;       // spatial size per image
  int pitch;      // channels of the allocated tensor
  int coff;       // first channel of this view
  int C;          // channels of this view
};

// Implicit-GEMM convolution:  D[m, co] = sum_t sum_ci  X[pix(m, t), ci] * Wt[(wrow[t]*Cin + ci), co]
// over a virtual output grid m = (n, my, mx) in N x MH x MW;
//   input pixel  (my*istride + dy[t], mx*istride + dx[t])  (out of range -> 0, TF SAME padding)
//   output pixel (my*ostride + oy0,   mx*ostride + ox0)    (ostride 2 = one phase of a transposed conv)
// epilogue: v = acc*scale[co] + shift[co]; relu6?; clip01?; v += residual?; store.
struct ConvParams {
  View in, out, res;         // res.ptr == nullptr: no residual
  int N, MH, MW;
  int istride, ostride, oy0, ox0;
  int ntaps;
  int dy[9], dx[9], wrow[9];
  int wtaps;                 // taps in the weight tensor (k*k); wrow[t] < wtaps
  int Cin, Cout;
  const float* w;            // FP32 [ktaps*Cin][Cout] (TF kernel layout flattened)
  const void* w16;           // 16-bit operand copy, UMMA tile-swizzled (emd_umma.cu); may be null
  const float* scale;        // [Cout]
  const float* shift;        // [Cout]
  int relu6, clip01;
  int out_f32;               // output tensor is float regardless of the element type (final layer)
  int in_f32;                // input tensor is float regardless of the element type (network input)
};

struct DwParams {
  View in, out;
  int N, OH, OW;
  int stride, rate, pad;     // input row = oy*stride - pad + ky*rate
  const float* w;            // [9][C] FP32, tap-major
  int in_f32;
};

struct ResizeParams {        // TF1 legacy bilinear (DMG:344, 494) + optional affine/ReLU6 (DMG:345)
  View in, out;
  int N;
  const float* scale;        // may be null
  const float* shift;
  int relu6;
};

// 1-channel input layers (network stem): out[p][c] = relu6(scale[c] * (w[c] * d(p)) + shift[c]) with
// d(p) = depthwise 3x3 of the input at p (cnn0, DMG:396) or the input sampled at stride `istride`
// (residual0, DMG:407).  Input is always the FP32 network input.
struct StemParams {
  const float* in; int IH, IW;
  View out; int N;
  const float* dw;           // [9] or nullptr
  const float* w;            // [Cout]
  const float* scale; const float* shift;
  int istride, relu6;
};

struct PoolParams { View in, out; int N; };   // 2x2 average, stride 2 (DMG:331-335)

// launchers (emd_kernels_simt.cu); et = element type of the activations
cudaError_t launch_conv_simt(const ConvParams& p, int et, cudaStream_t s);
cudaError_t launch_dw3x3(const DwParams& p, int et, cudaStream_t s);
bool dw_strip_supported(const DwParams& p, int et);   // 16-bit, stride 1: warp = 64 channels of one pixel column, window slides down the column
cudaError_t launch_dw_strip(const DwParams& p, int et, cudaStream_t s);
bool dw_tile_supported(const DwParams& p, int et);    // 16-bit, stride 1, rate 1: warp = 8 x 8 pixel tile x 64 channels, halo staged in smem by cp.async
cudaError_t launch_dw_tile(const DwParams& p, int et, cudaStream_t s);
bool dw_reg_supported(const DwParams& p, int et);    // register-strip form (emd_kernels_simt.cu): stride 1, rate 1, any map size
cudaError_t launch_dw_reg(const DwParams& p, int et, cudaStream_t s);
cudaError_t launch_resize(const ResizeParams& p, int et, cudaStream_t s);
cudaError_t launch_avgpool(const PoolParams& p, int et, cudaStream_t s);
cudaError_t launch_stem(const StemParams& p, int et, cudaStream_t s);
cudaError_t launch_cast(const float* src, void* dst, size_t n, int et, cudaStream_t s);
cudaError_t launch_uncast(const void* src, float* dst, size_t n, int et, cudaStream_t s);

// emd_dw.cu: TMA-fed depthwise 3x3 (stride 1) for the 16-bit modes
bool dw_tma_supported(const DwParams& p, int et);
cudaError_t launch_dw_tma(const DwParams& p, int et, int num_sms, cudaStream_t s);
bool dw_s2_tma_supported(const DwParams& p, int et);   // stride 2 (cnnN_strided)
cudaError_t launch_dw_s2_tma(const DwParams& p, int et, int num_sms, cudaStream_t s);
bool final_tma_supported(const ConvParams& p, int et);
cudaError_t launch_final_tma(const ConvParams& p, float scale, float shift, int et, int num_sms, cudaStream_t s);
// same layer on the tensor cores (emd_final.cu): channel reduction as a per-tile GEMM over the halo, then a 9-tap gather
bool final_umma_supported(const ConvParams& p, int et);
cudaError_t launch_final_umma(const ConvParams& p, float scale, float shift, int et, int num_sms, cudaStream_t s);

// Split of the output channels into UMMA N tiles (<= 256 columns each, multiples of 16); the packed
// B operand (umma_pack_weights) and both tcgen05 kernels share it.
constexpr int kMaxNTiles = 4;
struct NTiling { int nt; int n0[kMaxNTiles]; int rows[kMaxNTiles]; int rows_before[kMaxNTiles]; int maxrows; };
inline NTiling make_ntiling(int Cout) {
  NTiling t;
  const int cpad = (Cout + 15) & ~15;
  t.nt = (cpad + 255) / 256;
  const int base = (((cpad + t.nt - 1) / t.nt) + 15) & ~15;
  t.maxrows = 0;
  int acc = 0;
  for (int i = 0; i < kMaxNTiles; ++i) { t.n0[i] = 0; t.rows[i] = 0; t.rows_before[i] = 0; }
  for (int i = 0; i < t.nt && i < kMaxNTiles; ++i) {
    t.n0[i] = i * base;
    t.rows[i] = (cpad - i * base) < base ? (cpad - i * base) : base;
    t.rows_before[i] = acc;
    acc += t.rows[i];
    if (t.rows[i] > t.maxrows) t.maxrows = t.rows[i];
  }
  return t;
}

// emd_umma.cu: tcgen05 implicit GEMM (16-bit element types only)
bool umma_supported(const ConvParams& p, int et);
cudaError_t launch_conv_umma(const ConvParams& p, int et, int num_sms, cudaStream_t s);
// packs FP32 [K][Cout] weights into the UMMA tile image; returns bytes needed when dst == nullptr
size_t umma_pack_weights(const float* w, int ntaps, int Cin, int Cout, int et, void* dst_host);

// emd_fused.cu: second-generation tcgen05 kernel for block-tiled output grids (TMA-store epilogue; optional
// depthwise-3x3 producer: dw = [9][Cin] FP32 weights, p.in = the depthwise input)
bool fused_supported(const ConvParams& p, int et, const float* dw);
cudaError_t launch_conv_fused(const ConvParams& p, int et, const float* dw, int num_sms, cudaStream_t s);
// the 4 sub-pixel phases of one transposed conv (same input, weights, epilogue; different taps / output offsets) as ONE launch
bool fused_multi_supported(const ConvParams* ps, int nvar, int et);
cudaError_t launch_conv_fused_multi(const ConvParams* ps, int nvar, int et, int num_sms, cudaStream_t s);

// Tuning and A/B switches: ONE struct, filled once per process from the EMD_* environment (emd_tuning.cu) and changed at run
// time only through emd_set_option (include/emd.h; the tests force kernel variants with it).  Every optimisation that
// replaced a simpler path keeps that path behind one of these, so a regression can be bisected on the GPU box without a rebuild.
struct Tuning {
  int umma = 1;            // EMD_DISABLE_UMMA=1 -> 0: GEMM-class layers of the 16-bit modes on the CUDA-core kernel
  int fused = 1;           // EMD_DISABLE_FUSED: first-generation tcgen05 kernel (emd_umma.cu) instead of emd_fused.cu
  int tma = 1;             // EMD_DISABLE_TMA: first-generation kernel without TMA tensor maps
  int pair = 1;            // EMD_DISABLE_PAIR: no cta_group::2 CTA pairs
  int final_umma = 1;      // EMD_DISABLE_FINAL_UMMA: final 3x3 conv on the CUDA-core reduction kernel
  int pdl = 1;             // EMD_DISABLE_PDL: no programmatic dependent launch
  int graphs = 1;          // EMD_DISABLE_GRAPH: no CUDA-graph replay of whole passes
  int sliced_io = 1;       // EMD_DISABLE_SLICED_IO: host-buffer passes as the two-chunk pipeline
  int halves = 1;          // EMD_DISABLE_HALVES: no half-batch head / tail in host-buffer passes
  int mid_graph = 1;       // EMD_DISABLE_MID_GRAPH: no graph replay of the whole-batch middle section
  int skip_taps = 1;       // EMD_DISABLE_SKIP_TAPS: dilated / transposed convs multiply every tap on every tile, padding or not
  int dw_tile = 1;         // EMD_DISABLE_DW_TILE: small-map depthwise on the strip kernel (no shared-memory staging)
  int dw_reg = 1;          // EMD_DISABLE_DW_REG: small-map depthwise on the shared-memory tile kernel instead of the register-strip kernel
  int dw_reg_all = 0;      // EMD_DW_REG_ALL=1: the register-strip kernel also where the TMA-fed depthwise kernel applies (A/B)
  int dw_strip = 1;        // EMD_DISABLE_DW_STRIP: small-map / dilated depthwise on the one-thread-per-pixel kernel
  int dw_cols = 1;         // EMD_DISABLE_DW_COLS: depthwise producer with one pixel column per thread (first form)
  int pad_pitch = 1;       // EMD_DISABLE_PAD_PITCH: 728-channel tensors dense (1456-byte pixels) instead of padded to 768 channels (read at emd_create)
  int fork_sms = 0;        // EMD_FORK_SMS=n: the decoder's 1x1 residual convs run beside the separable block that reads the same tensor, on n SMs of a second stream (0 = one kernel at a time)
  int poison = 0;          // EMD_POISON=1 (debugging): the activation workspace is filled with NaNs before every pass
  int strict = 0;          // EMD_STRICT=1: a GEMM-class layer of a 16-bit mode that would run on the CUDA-core kernel is an error
  int graph_max_n = 32;    // EMD_GRAPH_MAX_N: largest batch replayed from a graph
  int pair_min_rows = 128; // EMD_PAIR_MIN_ROWS: CTA pairs only for N tiles of at least this many columns
  int pair_min_items = -1; // EMD_PAIR_MIN_ITEMS: CTA pairs only from this many pair items on (-1 = the number of SMs)
  int io_slices = 0, io_parts = 0;              // EMD_IO_SLICES, EMD_IO_PARTS (0 = defaults)
  int dw_stages = 0;       // EMD_DW_STAGES: halo ring depth of the stand-alone depthwise kernel (even, 2..8; 0 = default 4)
  int dw_sa = 0, dw_sb = 0, dw_sh = 0, dw_ring = 0;   // EMD_DW_SA/SB/SH/RING: stage counts of the fused depthwise mode (0 = defaults)
};
Tuning& tuning();
bool tuning_set(const char* name, long long value);      // false: unknown name
bool tuning_get(const char* name, long long* value);
inline bool pdl_enabled() { return tuning().pdl != 0; }

// which kernel a conv launcher picked (read by the engine right after the call; per host thread)
enum LaunchKind { LK_NONE = 0, LK_SIMT, LK_UMMA_GEN1, LK_FUSED_TAPS, LK_FUSED_PAIR, LK_FUSED_DW, LK_FINAL_UMMA, LK_FINAL_TMA, LK_COUNT };
int& last_launch_kind();

// emd_quality.cu: MSE / Huberised loss / SSIM of image pairs (d_out = 3 doubles per pair, d_partial = quality_partial_bytes)
size_t quality_partial_bytes(int n, int H, int W);
cudaError_t launch_quality(const float* a, const float* b, int n, int H, int W, double* d_partial, double* d_out, cudaStream_t s);

// emd_kernels_wrap.cu: whole-image wrapper kernels
cudaError_t launch_minmax(const void* img, int in_f64, size_t n, double* d_minmax /*[2]*/, void* d_partial,
                          cudaStream_t s);
cudaError_t launch_normalise_apply(const void* img, int in_f64, size_t n, const double* d_minmax, float* out,
                                   cudaStream_t s);
cudaError_t launch_gather(const float* img, int H, int W, const int* d_ys, const int* d_xs, int ny, int nx,
                          int crop, float* crops, cudaStream_t s);
cudaError_t launch_stitch(const float* tiles, const int* d_ys, const int* d_xs, int ny, int nx, int crop, int H,
                          int W, int clip, void* out /* double*, or float* with out_f32 */, int out_f32, cudaStream_t s);
size_t minmax_partial_bytes();
// Denoiser.preprocess (DEN:632-643) on the device: resize to S x S, scale0to1, NaN/Inf -> 0.5, scale0to1; d_tmp [S*S], d_mm [4] floats
cudaError_t launch_preprocess_crop(const float* d_img, int H, int W, int S, float* d_tmp, float* d_mm, float* d_out, cudaStream_t s);

}  // namespace emd
