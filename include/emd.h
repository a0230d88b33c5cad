/*
 * emd.h -- C ABI of the B200-native electron-micrograph denoiser (libemd.so).
 *
 * This is the drop-in boundary for ONE path of Jeffrey-Ede/AI-CV-Automation-Elect-Micr:
 * inference through the atrous-convolutional Xception encoder-decoder denoiser and its
 * crop-tile / normalise / stitch wrapper.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree):
 *
 *   DEN = machine_learning/denoiser.py          (the Denoiser class, variant-B graph)
 *   DMG = misc_py/denoiser-multi-gpu.py         (canonical variant-A graph, normalise helpers)
 *   TMP = misc_py/denoiser_class_function-tmp.py (stand-alone copy of Denoiser.denoise)
 *
 * Conventions
 *   - plain C: opaque handle, plain pointers and sizes, no C++/torch types, no exceptions.
 *   - every function returns 0 on success, a negative EMD_E* code on failure;
 *     emd_last_error() gives the message.  A build without a usable CUDA device fails loudly
 *     (EMD_ECUDA) -- there is no CPU fallback.
 *   - image/crop pointers may be device pointers, pinned host pointers or pageable host pointers;
 *     the library classifies them with cudaPointerGetAttributes.  With a host output pointer the
 *     call returns after the result has landed; with device pointers it is asynchronous on
 *     `stream` (a cudaStream_t passed as void*; NULL = the engine's own stream).
 *   - a handle owns one device, its weights and its workspace; it is NOT thread-safe.  Handles on
 *     different GPUs are independent (micrographs/crops shard image-wise, no collective).
 */
#ifndef EMD_H_
#define EMD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMD_OK        0
#define EMD_EINVAL   -1   /* bad argument / shape */
#define EMD_ECUDA    -2   /* CUDA runtime error (message has the cudaError string) */
#define EMD_ESTATE   -3   /* call order (e.g. forward before load_weights) */
#define EMD_ENOMEM   -4

/* arithmetic modes of the network */
#define EMD_MODE_FP32 0   /* CUDA-core FP32 validation mode (parity gate 1e-5) */
#define EMD_MODE_BF16 1   /* tcgen05/TMEM tensor-core path, BF16 operands, FP32 accumulate */
#define EMD_MODE_FP16 2   /* same kernels with FP16 operands (all activations are ReLU6-bounded) */

/* graph variants */
#define EMD_VARIANT_A 0   /* DMG:200-540, dense dilated ASPP, in-graph clip */
#define EMD_VARIANT_B 1   /* DEN:58-398, the deployed class file: separable ASPP branches + extra BN/ReLU6, identity image branch, no in-graph clip */

/* emd_denoise_image flags */
#define EMD_FLAG_PREPROCESS   1   /* DEN:655-656 (repaired): NaN/Inf->0.5 then scale0to1 */
#define EMD_FLAG_POSTPROCESS  2   /* DEN:679-680: clip(0,1) */
#define EMD_FLAG_INPUT_F64    4   /* img is const double* (np.random.rand in DEN:708 is float64) */
#define EMD_FLAG_OUTPUT_F32   8   /* out is float* (the float64 overlap average rounded once); default double* like the reference's np.zeros accumulators (DEN:658) */

typedef struct emd_engine emd_engine;

int         emd_version(void);
/* last error message of this handle (or, with NULL, of the last failed emd_create on this thread) */
const char* emd_last_error(const emd_engine* e);

/* Replaces Denoiser.__init__ (DEN:587-630): pins one GPU (DEN:591 CUDA_VISIBLE_DEVICES), builds the
 * static layer schedule that stands in for get_model_fn/_tower_fn/architecture graph construction
 * (DEN:463-581, DMG:200-540) for crops of `cropsize` (multiple of 16; reference: 512, DMG:112), and
 * sizes the workspace for `max_batch` crops per pass (larger batches are processed in chunks). */
int emd_create(emd_engine** out, int device, int cropsize, int variant, int max_batch);
int emd_destroy(emd_engine* e);

/* Replaces tf.train.Saver().restore (DEN:626-627): takes the packed weight blob produced by the
 * host-side exporter (BN folded per DMG:210-223; layouts in DESIGN.md) and uploads FP32, BF16 and
 * FP16 copies. */
int emd_load_weights(emd_engine* e, const void* blob, size_t nbytes);

/* Replaces sess.run(self._tower_preds, feed_dict=...) (DEN:646-647) = architecture() forward
 * (DMG:392-540) for a batch: crops [n,S,S] f32 in -> out [n,S,S] f32 (variant A: clipped to [0,1]
 * in-graph, DMG:534-538). */
int emd_forward(emd_engine* e, const float* crops, int n, float* out, int mode, void* stream);

/* emd_forward with HOST buffers for a stream of batches (the reference calls sess.run once per crop and waits, DEN:646-647;
 * there is no counterpart): returns without waiting, the results have landed only after emd_synchronize.  Successive calls
 * pipeline -- the upload of batch i+1 and the download of batch i-1 run under batch i's network pass.  The caller keeps
 * `crops` and `out` untouched until emd_synchronize and gives every outstanding call its own `out`.  Same results, bit
 * for bit, as emd_forward.  (Batches of fewer than 16 crops, or an engine built for fewer, run synchronously.) */
int emd_forward_async(emd_engine* e, const float* crops, int n, float* out, int mode, void* stream);
int emd_synchronize(emd_engine* e, void* stream);

/* Tile plan of Denoiser.denoise (DEN:661-669 == TMP:11-19), repaired per SURVEY App. D-2/D-3
 * (integer origins, round-half-to-even, last tile clamped).  Pure integer host code.
 * ys/xs must hold at least H/(crop-overlap)+1 and W/(crop-overlap)+1 ints. */
int emd_plan_tiles(int H, int W, int crop, int overlap, int* ys, int* xs, int* ny, int* nx);

/* Replaces preprocess (DMG:853-858) + scale0to1 (DEN:684-695 == DMG:817-828) on a whole image:
 * NaN->0.5, Inf->0.5, then (x-min)/(max-min) in the input's precision (IEEE sub/div), constant
 * image -> 0.5, result cast to f32.  in_f64 != 0: img is const double*. */
int emd_normalise(emd_engine* e, const void* img, int in_f64, int H, int W, float* out, void* stream);

/* Replaces Denoiser.preprocess (DEN:632-643), the single-crop path of denoise_crop(preprocess=True): img [H,W] f32 (host or device)
 * -> out [S,S] f32: cv2.resize to the crop size (INTER_LINEAR arithmetic: half-pixel centres, edge replicate, float
 * coefficients), scale0to1, NaN -> 0.5, Inf -> 0.5, scale0to1 -- in the class file's order: the first min-max runs BEFORE the
 * NaN/Inf replacement and propagates NaN like numpy. */
int emd_preprocess_crop(emd_engine* e, const float* img, int H, int W, float* out, void* stream);

/* Replaces the crop slicing of DEN:671-673: img [H,W] f32 -> crops [ny*nx,crop,crop] f32, row-major
 * over (i,j) like the reference's double loop. */
int emd_gather_crops(emd_engine* e, const float* img, int H, int W, const int* ys, const int* xs,
                     int ny, int nx, int crop, float* crops, void* stream);

/* Replaces the accumulate / contributions / divide / clip of DEN:658-659, 671-680 (repaired per
 * App. D-4: +=): out[r,c] = sum of covering tiles / count, float64 like the reference's np.zeros
 * accumulators; gather form (no atomics), stitch weights exact in binary FP. */
int emd_stitch(emd_engine* e, const float* tiles, const int* ys, const int* xs, int ny, int nx,
               int crop, int H, int W, int clip, double* out, void* stream);

/* Replaces Denoiser.denoise (DEN:653-682 == TMP:3-32) end to end on one GPU: normalise -> tile ->
 * batched forward -> overlap-averaged stitch -> clip.  img [H,W] f32 (or f64 with
 * EMD_FLAG_INPUT_F64), out [H,W] f64 (f32 with EMD_FLAG_OUTPUT_F32). */
int emd_denoise_image(emd_engine* e, const void* img, int H, int W, int overlap, int flags,
                      int mode, void* out, void* stream);

/* A stream of `count` micrographs of one size (BASELINE.json configs[3]; the reference calls Denoiser.denoise once per
 * image, DEN:653): the same result as emd_denoise_image on each, bit for bit, but image i+1's upload + normalise + tile
 * gather and image i-1's stitch + download run on their own streams under image i's network pass (two staging slots).
 * imgs[i] / outs[i]: host (pinned for real overlap) or device pointers, formats as in emd_denoise_image. */
int emd_denoise_stream(emd_engine* e, const void* const* imgs, int count, int H, int W, int overlap,
                       int flags, int mode, void* const* outs, void* stream);

/* Image-quality metrics of the reference's training / evaluation code (misc_py/denoiser-multi-gpu.py) between n pairs of
 * [H,W] f32 images a[i], b[i] (host or device; H, W >= 11): out[3*i + 0] = mean squared error (DMG:772),
 * out[3*i + 1] = the trainer's Huberised loss of it, mse < 0.001 ? 1000*mse : sqrt(1000*mse) (DMG:773),
 * out[3*i + 2] = mean SSIM as tf_ssim computes it (DMG:142-167: 11x11 Gaussian window, sigma 1.5, VALID, K1 0.01, K2 0.03,
 * L 1).  out = 3*n doubles in HOST memory. */
int emd_quality(emd_engine* e, const float* a, const float* b, int n, int H, int W, double* out, void* stream);

/* ---- parity / measurement hooks (no reference counterpart; used by tests and bench.py) ---- */

/* keep != 0: every activation gets its own buffer (no reuse) so emd_get_activation works */
int emd_set_keep_activations(emd_engine* e, int keep);
/* copy a named activation of the last emd_forward to host as f32 NHWC; dims = {n,H,W,C} */
int emd_get_activation(emd_engine* e, const char* name, float* out, size_t cap_elems, int dims[4]);
/* run one fused layer on host NHWC f32 inputs (in2 = residual operand or NULL); out NHWC f32 */
int emd_run_layer(emd_engine* e, const char* name, const float* in, const float* in2, int n,
                  float* out, size_t out_cap_elems, int mode, int out_dims[4]);
/* number of kernels this engine has launched since creation; of those, tcgen05 (UMMA) kernels */
long long emd_kernel_launches(const emd_engine* e);
long long emd_tensor_core_launches(const emd_engine* e);
/* other counters by name: "launches", "tensor_core_launches", "graph_replays", "workspace_bytes" (activation workspace as
 * planned now), and conv launches by the kernel that ran them:
 * "conv_fused_pair" (emd_fused.cu, cta_group::2 CTA pairs), "conv_fused_taps", "conv_fused_dw" (depthwise computed inside the GEMM
 * kernel), "conv_tcgen05_gen1" (emd_umma.cu), "conv_cuda_core" (FP32 mode, or a 16-bit shape no tensor-core kernel supports),
 * "final_tcgen05", "final_cuda_core".  -1 = unknown name.  The parity tests assert with these that the kernel under test ran. */
long long emd_counter(const emd_engine* e, const char* name);
/* Tuning / A-B switches (csrc/emd_kernels.h, struct Tuning): process-wide, initialised once from the EMD_* environment, changed
 * by name here; `e` (may be NULL) drops its captured graphs so the change takes effect.  Names: umma, fused, tma, pair,
 * final_umma, pdl, graphs, sliced_io, halves, mid_graph, pad_pitch, poison, dw_cols, dw_reg, dw_reg_all, dw_tile, dw_strip, skip_taps, strict, graph_max_n, pair_min_rows, pair_min_items,
 * io_slices, io_parts, dw_stages, dw_sa, dw_sb, dw_sh, dw_ring.  strict = 1: a GEMM-class layer of a 16-bit mode that no tensor-core
 * kernel supports is an error (EMD_ESTATE) instead of a silent CUDA-core launch. */
int       emd_set_option(emd_engine* e, const char* name, long long value);
long long emd_get_option(const char* name);   /* -1 = unknown name */
/* passes replayed from a captured CUDA graph (batches <= 32 from their third pass on; EMD_DISABLE_GRAPH=1 turns it off, EMD_GRAPH_MAX_N changes the limit) */
long long emd_graph_replays(const emd_engine* e);
/* on = 0: the 16-bit modes run their GEMM-class layers on the CUDA-core kernel with the same
 * 16-bit operand values (A/B check of the tcgen05 kernel); default on */
int emd_set_tensor_cores(emd_engine* e, int on);
/* device time in ms of the kernels of the last emd_forward that belong to layer `name`
 * (needs emd_set_profile(e,1); serialises the stream) */
int emd_set_profile(emd_engine* e, int on);
int emd_num_steps(const emd_engine* e);
/* flops / bytes: ALGORITHMIC work per crop of the step as it ran last (a depthwise step computed inside the
 * next step's GEMM kernel reports 0 and the GEMM step reports depthwise input + its own output) */
int emd_step_info(const emd_engine* e, int idx, char* name, size_t name_cap, float* ms,
                  double* flops, double* bytes);
/* kernels launched for step idx in the last forward (a transposed conv is 4 sub-pixel phase launches) */
int emd_step_launches(const emd_engine* e, int idx);

#ifdef __cplusplus
}
#endif
#endif /* EMD_H_ */
