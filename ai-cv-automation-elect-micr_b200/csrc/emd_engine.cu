// emd_engine.cu -- the engine behind the C ABI of include/emd.h.
//
// The reference builds a TensorFlow graph (get_model_fn/_tower_fn/architecture, DEN:463-581,
// DMG:200-540) and runs it with one sess.run per 512x512 crop (DEN:646-647).  Here the graph is a
// static schedule of fused steps over NHWC activations in one workspace arena; a batch of crops
// goes through every step together.  Nothing in this file computes on the CPU: if CUDA is not
// usable every entry point fails with EMD_ECUDA.
#include "../../include/emd.h"
#include "emd_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace emd;

namespace {

thread_local std::string g_create_error;

struct Tensor {
  std::string name;
  int H, W, C;
  int pitch = 0;        // elements per pixel as laid out in the arena (>= C: odd channel counts are padded to whole 128-byte lines)
  bool external;        // network input / output: lives in caller (or staging) memory, always f32
  size_t offset = 0;    // bytes into the arena (image 0 of the batch)
  size_t bytes = 0;     // bytes reserved (max_batch images, 4-byte elements)
  int first = 1 << 30, last = -1;
};

struct Ref { int t = -1, coff = 0, C = 0; };

enum StepKind { SK_DW, SK_CONV, SK_DECONV, SK_RESIZE, SK_POOL };

struct Step {
  StepKind kind;
  std::string name;    // unique step name
  std::string layer;   // reference layer this step belongs to (emd_run_layer group)
  Ref in, out, res;
  int Cin = 0, Cout = 0, k = 1, stride = 1, rate = 1;
  bool relu6 = false, clip01 = false;
  std::string wname;   // blob prefix
  const float *w = nullptr, *scale = nullptr, *shift = nullptr, *dw = nullptr;
  void* w16[3] = {nullptr, nullptr, nullptr};  // indexed by ElemType: UMMA tile image
  const float* wr[3] = {nullptr, nullptr, nullptr};  // FP32 weights rounded to the 16-bit type (CUDA-core fallback of the 16-bit modes)
  float scale0 = 1.f, shift0 = 0.f;  // host copies of scale[0] / shift[0] (single-output-channel layers)
  float ms = 0.f;
  double flops = 0, bytes = 0;  // per crop: algorithmic FLOPs and (16-bit storage) HBM bytes
  double in_bytes = 0;          // the part of `bytes` that is the read of the step's input
  bool fused = false;           // last run: SK_DW computed inside the next step's GEMM kernel / SK_CONV that absorbed it
  int nlaunch = 0;              // kernels launched for this step in the last run
};

struct BlobEntry { uint32_t rows, cols; size_t offset; };

}  // namespace

// launch counters: all kernels, tcgen05 kernels, and conv launches by kernel kind (LaunchKind, emd_kernels.h)
struct Counts {
  long long launches = 0, umma = 0, kind[LK_COUNT] = {};
  Counts& operator+=(const Counts& o) { launches += o.launches; umma += o.umma; for (int i = 0; i < LK_COUNT; ++i) kind[i] += o.kind[i]; return *this; }
  Counts operator-(const Counts& o) const { Counts r = *this; r.launches -= o.launches; r.umma -= o.umma; for (int i = 0; i < LK_COUNT; ++i) r.kind[i] -= o.kind[i]; return r; }
};
struct GraphSlot { cudaGraphExec_t exec = nullptr; int calls = 0; bool failed = false; Counts cnt; };

struct emd_engine {
  int device = 0, S = 512, variant = 0, max_batch = 1, num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<Tensor> tensors;
  std::vector<Step> steps;
  std::map<std::string, Ref> named;      // activation name -> view
  int t_input = -1, t_output = -1;
  bool keep = false, profile = false, weights_loaded = false, use_umma = true;
  char* arena = nullptr; size_t arena_bytes = 0;
  char* d_blob = nullptr; size_t blob_bytes = 0;
  std::map<std::string, BlobEntry> entries;
  std::vector<void*> w16_allocs;
  Counts cnt;
  int last_n = 0, last_et = 0;
  // staging for host I/O of emd_forward
  float *d_stage_in = nullptr, *d_stage_out = nullptr;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;   // H2D / D2H streams of emd_forward with host buffers
  cudaEvent_t ev_in[2] = {}, ev_done[2] = {}, ev_out[2] = {}, ev_start = nullptr;
  bool chained = false;                                   // an emd_forward_async call is outstanding: the next one pipelines behind it
  cudaStream_t chain_stream = nullptr;                    // ... on this compute stream
  int last_input_step = 0;                                // last step that reads the network input
  int head_end = -1, tail_start = 1 << 30;                // half-batch phases of the host-buffer pass (plan_arena)
  static constexpr int kSlices = 8;                       // sliced first / last step of a pass (emd_forward, host buffers)
  cudaEvent_t ev_sl_in[kSlices] = {}, ev_sl_fin[kSlices] = {}, ev_in_free = nullptr, ev_out_free = nullptr;
  // whole-image pipeline buffers (grown on demand)
  void* d_img_raw = nullptr; size_t img_raw_bytes = 0;
  char *d_q_a = nullptr, *d_q_b = nullptr, *d_q_partial = nullptr; size_t q_a_bytes = 0, q_b_bytes = 0, q_partial_bytes = 0;   // emd_quality
  float* d_img = nullptr; size_t img_bytes = 0;
  float *d_crops = nullptr, *d_tiles = nullptr; size_t crops_bytes = 0, tiles_bytes = 0;
  double* d_sout = nullptr; size_t sout_bytes = 0;
  double* d_minmax = nullptr; void* d_partial = nullptr;
  int* d_origins = nullptr;  // ys then xs, 2*256 ints
  // whole-micrograph path (emd_denoise_image / emd_denoise_stream): two slots, so that image i+1's upload + normalise + tile
  // gather (pre stream) and image i-1's stitch + download (post stream) run under image i's network pass (compute stream)
  struct ImgSlot {
    char* d_raw = nullptr; size_t raw_bytes = 0;
    float* d_norm = nullptr; size_t norm_bytes = 0;
    float *d_crops = nullptr, *d_tiles = nullptr; size_t crops_bytes = 0, tiles_bytes = 0;
    char* d_sout = nullptr; size_t sout_bytes = 0;
    double* d_minmax = nullptr; char* d_partial = nullptr; size_t mm_bytes = 0, partial_bytes = 0;
    cudaEvent_t ev_pre = nullptr, ev_net = nullptr, ev_post = nullptr;
  } slots[2];
  cudaStream_t s_pre = nullptr, s_post = nullptr;
  cudaEvent_t ev_img_start = nullptr;
  std::vector<cudaEvent_t> events;
  // CUDA graphs of whole passes for small batches
  std::map<int, GraphSlot> graphs;
  // forked pairs (option fork_sms): fork_side[i] = j when step j (a 1x1 conv) reads the same tensor as the separable block whose
  // GEMM step is i and may run beside it on the other stream, each kernel on its share of the SMs
  std::vector<int> fork_side;
  cudaStream_t fork_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int fork_skip = -1;
  std::map<int, GraphSlot> mid_graphs;   // whole-batch middle section of the host-buffer pass (run_network_sliced)
  float *g_in = nullptr, *g_out = nullptr;
  long long graph_replays = 0;
};

namespace {

void drop_graphs(emd_engine* e) {
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
  for (auto& kv : e->mid_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->mid_graphs.clear();
}

int fail(emd_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CU(e, call)                                                                         \
  do {                                                                                      \
    cudaError_t _r = (call);                                                                \
    if (_r != cudaSuccess)                                                                  \
      return fail(e, EMD_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

// Outstanding emd_forward_async calls share the workspace and the staging buffers with everything else the engine does: any
// other entry point (or an async call on another stream) first waits for them.
int drain_chain(emd_engine* e) {
  if (!e->chained) return EMD_OK;
  CU(e, cudaStreamSynchronize(e->chain_stream));
  if (e->copy_out) CU(e, cudaStreamSynchronize(e->copy_out));
  e->chained = false;
  return EMD_OK;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------------------
// schedule construction (variant A: DMG:392-540)
// ---------------------------------------------------------------------------------------------
struct Builder {
  emd_engine* e;
  int add_tensor(const std::string& name, int H, int W, int C, bool external = false) {
    Tensor t; t.name = name; t.H = H; t.W = W; t.C = C; t.pitch = C; t.external = external;
    e->tensors.push_back(t);
    int id = (int)e->tensors.size() - 1;
    e->named[name] = Ref{id, 0, C};
    return id;
  }
  Ref whole(int t) { return Ref{t, 0, e->tensors[t].C}; }
  Ref slice(int t, int coff, int C, const std::string& alias) {
    Ref r{t, coff, C};
    if (!alias.empty()) e->named[alias] = r;
    return r;
  }
  // strided_conv_block (DMG:250-276): depthwise 3x3 (stride, rate) -> pointwise + folded BNx2 + ReLU6 (+ residual)
  void sep(const std::string& name, Ref in, Ref out, int Cin, int Cout, int stride, int rate, Ref res = Ref()) {
    const Tensor& ti = e->tensors[in.t];
    int OH = (ti.H + stride - 1) / stride, OW = (ti.W + stride - 1) / stride;
    int tmp = add_tensor(name + ":dw", OH, OW, Cin);
    Step d; d.kind = SK_DW; d.name = name + ":dw"; d.layer = name; d.in = in; d.out = whole(tmp);
    d.Cin = d.Cout = Cin; d.k = 3; d.stride = stride; d.rate = rate; d.wname = name;
    e->steps.push_back(d);
    Step c; c.kind = SK_CONV; c.name = name; c.layer = name; c.in = whole(tmp); c.out = out; c.res = res;
    c.Cin = Cin; c.Cout = Cout; c.k = 1; c.relu6 = true; c.wname = name;
    e->steps.push_back(c);
  }
  // conv_block_not_sep / residual_conv / ASPP convs: k x k conv + bias + folded BN + ReLU6
  void conv(const std::string& name, Ref in, Ref out, int Cin, int Cout, int k, int stride, int rate, bool relu6,
            Ref res = Ref(), bool clip01 = false, const std::string& layer = "") {
    Step c; c.kind = SK_CONV; c.name = name; c.layer = layer.empty() ? name : layer; c.in = in; c.out = out; c.res = res;
    c.Cin = Cin; c.Cout = Cout; c.k = k; c.stride = stride; c.rate = rate; c.relu6 = relu6; c.clip01 = clip01;
    c.wname = name;
    e->steps.push_back(c);
  }
  void deconv(const std::string& name, Ref in, Ref out, int Cin, int Cout) {  // deconv_block, DMG:278-289
    Step c; c.kind = SK_DECONV; c.name = name; c.layer = name; c.in = in; c.out = out;
    c.Cin = Cin; c.Cout = Cout; c.k = 3; c.stride = 2; c.relu6 = true; c.wname = name;
    e->steps.push_back(c);
  }
};

void build_schedule(emd_engine* e) {
  Builder b{e};
  const int S = e->S, f0 = 64, f1 = 128, f2 = 256, f3 = 728, f4 = 728, ao = 256;
  e->t_input = b.add_tensor("input", S, S, 1, true);
  e->t_output = b.add_tensor("output", S, S, 1, true);
  const int concat1 = b.add_tensor("concat1", S / 2, S / 2, f2 + f1);   // [deconv2to1, cnn0_strided] DMG:509-511
  const int concat2 = b.add_tensor("concat2", S / 4, S / 4, ao + f1);   // [upsampled aspp, cnn1_strided] DMG:497-499
  const int cat5 = b.add_tensor("aspp_concat", S / 16, S / 16, 5 * f4);  // DMG:348-350
  Ref in = b.whole(e->t_input);

  // encoding blocks 0-3 (DMG:395-453): sep, sep, sep(stride 2) + 1x1 stride-2 residual, added after ReLU6
  struct Enc { int cin, a, b, c; };
  const Enc enc[4] = {{1, f0, f0, f1}, {f1, f1, f1, f1}, {f1, f2, f2, f2}, {f2, f3, f3, f3}};
  Ref cur = in;
  int size = S;
  for (int i = 0; i < 4; ++i) {
    const std::string n = "cnn" + std::to_string(i);
    int ta = b.add_tensor(n, size, size, enc[i].a);
    int tb = b.add_tensor(n + "_last", size, size, enc[i].b);
    int tr = b.add_tensor("residual" + std::to_string(i), size / 2, size / 2, enc[i].c);
    Ref out;
    const std::string en = "enc" + std::to_string(i);
    if (i == 0) out = b.slice(concat1, f2, f1, en);
    else if (i == 1) out = b.slice(concat2, ao, f1, en);
    else out = b.whole(b.add_tensor(en, size / 2, size / 2, enc[i].c));
    b.sep(n, cur, b.whole(ta), enc[i].cin, enc[i].a, 1, 1);
    b.sep(n + "_last", b.whole(ta), b.whole(tb), enc[i].a, enc[i].b, 1, 1);
    b.conv("residual" + std::to_string(i), cur, b.whole(tr), enc[i].cin, enc[i].c, 1, 2, 1, true);
    b.sep(n + "_strided", b.whole(tb), out, enc[i].b, enc[i].c, 2, 1, b.whole(tr));
    cur = out;
    size /= 2;
  }
  // encoding block 4 (DMG:455-466) and the 11 middle blocks (DMG:468-469 -> 375-390)
  const int s16 = S / 16;
  Ref trunk = cur;
  for (int blk = -1; blk < 11; ++blk) {
    const std::string base = blk < 0 ? std::string("cnn4_") : "mid" + std::to_string(blk) + "_";
    int t0 = b.add_tensor(base + "0", s16, s16, f4);
    int t1 = b.add_tensor(base + "1", s16, s16, f4);
    int t2 = b.add_tensor(blk < 0 ? std::string("trunk4") : "trunk_mid" + std::to_string(blk), s16, s16, f4);
    b.sep(base + "0", trunk, b.whole(t0), f4, f4, 1, 1);
    b.sep(base + "1", b.whole(t0), b.whole(t1), f4, f4, 1, 1);
    b.sep(base + "2", b.whole(t1), b.whole(t2), f4, f4, 1, 1, trunk);
    trunk = b.whole(t2);
  }
  // ASPP: branches write straight into their slice of the 3640-channel concat
  b.conv("aspp_1x1", trunk, b.slice(cat5, 0, f4, "aspp_1x1"), f4, f4, 1, 1, 1, true);
  const int rates[3] = {6, 12, 18};
  auto bn_relu6 = [&](const std::string& name, Ref in, Ref out, int C) {   // stand-alone BatchNorm + ReLU6 = identity "resize" with an affine
    Step r; r.kind = SK_RESIZE; r.name = name; r.layer = name; r.in = in; r.out = out; r.Cin = r.Cout = C; r.relu6 = true; r.wname = name;
    e->steps.push_back(r);
  };
  if (e->variant == EMD_VARIANT_A) {   // DMG:291-361: dense dilated 3x3 branches; avg-pool -> 1x1 conv -> resize -> BN -> ReLU6
    for (int i = 0; i < 3; ++i) {
      const std::string n = "aspp_r" + std::to_string(rates[i]);
      b.conv(n, trunk, b.slice(cat5, (i + 1) * f4, f4, n), f4, f4, 3, 1, rates[i], true);
    }
    int tp = b.add_tensor("aspp_pool", s16 / 2, s16 / 2, f4);
    int ti = b.add_tensor("aspp_image_conv", s16 / 2, s16 / 2, f4);
    Step p; p.kind = SK_POOL; p.name = "aspp_pool"; p.layer = "aspp_image"; p.in = trunk; p.out = b.whole(tp);
    p.Cin = p.Cout = f4;
    e->steps.push_back(p);
    b.conv("aspp_image_conv", b.whole(tp), b.whole(ti), f4, f4, 1, 1, 1, false, Ref(), false, "aspp_image");
    e->steps.back().wname = "aspp_image";
    Step r; r.kind = SK_RESIZE; r.name = "aspp_image"; r.layer = "aspp_image"; r.in = b.whole(ti);
    r.out = b.slice(cat5, 4 * f4, f4, "aspp_image"); r.Cin = r.Cout = f4; r.relu6 = true; r.wname = "aspp_image:post";
    e->steps.push_back(r);
  } else {   // DEN:152-218: separable dilated branches (BN, BN, ReLU6) each followed by another BN -> ReLU6; the image-level
             // branch is resize_images(input, [32,32]) = identity -> BN -> ReLU6 (the pooled tensor is computed and discarded, DEN:185-200)
    for (int i = 0; i < 3; ++i) {
      const std::string n = "aspp_r" + std::to_string(rates[i]);
      int tb = b.add_tensor(n, s16, s16, f4);
      b.sep(n, trunk, b.whole(tb), f4, f4, 1, rates[i]);
      bn_relu6(n + "_post", b.whole(tb), b.slice(cat5, (i + 1) * f4, f4, n + "_post"), f4);
    }
    bn_relu6("aspp_image", trunk, b.slice(cat5, 4 * f4, f4, "aspp_image"), f4);
  }
  int taspp = b.add_tensor("aspp_pellet", s16, s16, ao);
  b.conv("aspp_pellet", b.whole(cat5), b.whole(taspp), 5 * f4, ao, 1, 1, 1, true);
  // decoder (DMG:494-531)
  {
    Step r; r.kind = SK_RESIZE; r.name = "upsample4"; r.layer = "upsample4"; r.in = b.whole(taspp);
    r.out = b.slice(concat2, 0, ao, "upsample4"); r.Cin = r.Cout = ao;
    e->steps.push_back(r);
  }
  struct Dec { const char* n; int cat; int cin, cout, size; };
  int d2a = b.add_tensor("deconv2_0", S / 4, S / 4, f2), d2r = b.add_tensor("residual2_d", S / 4, S / 4, f2);
  int dec2 = b.add_tensor("dec2", S / 4, S / 4, f2);
  b.sep("deconv2_0", b.whole(concat2), b.whole(d2a), ao + f1, f2, 1, 1);
  b.conv("residual2_d", b.whole(concat2), b.whole(d2r), ao + f1, f2, 1, 1, 1, true);
  b.sep("deconv2_1", b.whole(d2a), b.whole(dec2), f2, f2, 1, 1, b.whole(d2r));
  b.deconv("deconv2to1", b.whole(dec2), b.slice(concat1, 0, f2, "deconv2to1"), f2, f2);
  int d1a = b.add_tensor("deconv1_0", S / 2, S / 2, f1), d1r = b.add_tensor("residual1_d", S / 2, S / 2, f1);
  int dec1 = b.add_tensor("dec1", S / 2, S / 2, f1);
  b.sep("deconv1_0", b.whole(concat1), b.whole(d1a), f2 + f1, f1, 1, 1);
  b.conv("residual1_d", b.whole(concat1), b.whole(d1r), f2 + f1, f1, 1, 1, 1, true);
  b.sep("deconv1_1", b.whole(d1a), b.whole(dec1), f1, f1, 1, 1, b.whole(d1r));
  int d1to0 = b.add_tensor("deconv1to0", S, S, f1);
  b.deconv("deconv1to0", b.whole(dec1), b.whole(d1to0), f1, f1);
  int d0a = b.add_tensor("deconv0_0", S, S, f0), d0r = b.add_tensor("residual0_d", S, S, f0);
  int dec0 = b.add_tensor("dec0", S, S, f0);
  b.sep("deconv0_0", b.whole(d1to0), b.whole(d0a), f1, f0, 1, 1);
  b.conv("residual0_d", b.whole(d1to0), b.whole(d0r), f1, f0, 1, 1, 1, true);
  b.sep("deconv0_1", b.whole(d0a), b.whole(dec0), f0, f0, 1, 1, b.whole(d0r));
  // final 3x3 conv -> BN -> ReLU6 (DMG:531) + in-graph clip (DMG:534-538), written as f32
  // variant A clips in-graph (DMG:534-538); variant B returns the raw prediction (DEN:390-396), the wrapper clips (DEN:648-649)
  b.conv("final", b.whole(dec0), b.whole(e->t_output), f0, 1, 3, 1, 1, true, Ref(), e->variant == EMD_VARIANT_A);
}

// algorithmic work per crop of each step (16-bit storage), for roofline reporting
void annotate_work(emd_engine* e) {
  for (Step& s : e->steps) {
    const Tensor& ti = e->tensors[s.in.t];
    const Tensor& to = e->tensors[s.out.t];
    const double ipx = (double)ti.H * ti.W, opx = (double)to.H * to.W;
    const double in_b = ipx * s.in.C * (ti.external ? 4 : 2), out_b = opx * s.out.C * (to.external ? 4 : 2);
    const double res_b = s.res.t >= 0 ? opx * s.res.C * 2 : 0;
    switch (s.kind) {
      case SK_DW: s.flops = 2.0 * 9 * opx * s.Cin; s.bytes = in_b + out_b; s.in_bytes = in_b; break;
      case SK_CONV: {
        double taps = 0;  // in-bounds taps summed over output pixels
        const int half = s.k / 2;
        for (int ky = -half; ky <= half; ++ky)
          for (int kx = -half; kx <= half; ++kx) {
            const int dy = ky * s.rate, dx = kx * s.rate;
            const double vy = std::max(0, to.H - std::abs(dy)), vx = std::max(0, to.W - std::abs(dx));
            taps += (s.stride == 1) ? vy * vx : opx;
          }
        s.flops = 2.0 * taps * s.Cin * s.Cout;
        s.bytes = in_b / (s.stride * s.stride) + out_b + res_b + 2.0 * s.k * s.k * s.Cin * s.Cout / e->max_batch;
        s.in_bytes = in_b / (s.stride * s.stride);
        break;
      }
      case SK_DECONV:
        s.flops = 2.0 * ipx * 9 * s.Cin * s.Cout;
        s.bytes = in_b + out_b + 2.0 * 9 * s.Cin * s.Cout / e->max_batch;
        break;
      default: s.flops = 4.0 * opx * s.Cout; s.bytes = in_b + out_b; break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// arena: lifetime-aware placement, greedy by size
// ---------------------------------------------------------------------------------------------
int plan_arena(emd_engine* e) {
  drop_graphs(e);   // captured passes hold addresses of the old plan
  for (Tensor& t : e->tensors) { t.first = 1 << 30; t.last = -1; }
  for (int i = 0; i < (int)e->steps.size(); ++i) {
    const Step& s = e->steps[i];
    for (int t : {s.in.t, s.out.t, s.res.t}) {
      if (t < 0) continue;
      e->tensors[t].first = std::min(e->tensors[t].first, i);
      e->tensors[t].last = std::max(e->tensors[t].last, i);
    }
    // a depthwise step may be computed INSIDE the next step's GEMM kernel (emd_fused.cu, dw mode): its input is then read
    // while that step's output is being written, so the input must stay live (un-aliased) through the next step
    if (s.kind == SK_DW && i + 1 < (int)e->steps.size() && e->steps[i + 1].layer == s.layer)
      e->tensors[s.in.t].last = std::max(e->tensors[s.in.t].last, i + 1);
  }
  // Host-buffer passes (run_network_sliced) run the high-resolution head of the network (outputs >= S/2) and its
  // full-resolution tail once per HALF of the batch, so that the second half's upload and the first half's download hide
  // under the other half's compute.  Inside such a phase a late step of half 1 runs before an early step of half 2:
  // tensors touched by a phase must not share memory with each other.
  const int nsteps = (int)e->steps.size();
  e->head_end = -1; e->tail_start = nsteps;
  if (e->max_batch >= 16) {
    for (int i = 0; i < nsteps && e->tensors[e->steps[i].out.t].H * 2 >= e->S; ++i) e->head_end = i;
    for (int i = nsteps - 1; i >= 0 && e->tensors[e->steps[i].out.t].H >= e->S; --i) e->tail_start = i;
    if (e->head_end < 0 || e->tail_start >= nsteps || e->head_end + 1 >= e->tail_start) { e->head_end = -1; e->tail_start = nsteps; }
    for (Tensor& t : e->tensors) {
      if (t.last < 0) continue;
      if (t.first <= e->head_end) { t.first = 0; t.last = std::max(t.last, e->head_end); }
      if (t.last >= e->tail_start) { t.first = std::min(t.first, e->tail_start); t.last = nsteps - 1; }
    }
  }
  // forked pairs: the side conv runs WHILE the block before it runs, so its output must not share memory with anything that
  // block (or its depthwise step) still uses
  for (int i = 0; i < (int)e->fork_side.size(); ++i)
    if (e->fork_side[i] >= 0) {
      Tensor& t = e->tensors[e->steps[e->fork_side[i]].out.t];
      if (t.last >= 0) t.first = std::min(t.first, i - 1);
    }
  std::vector<int> order;
  for (int i = 0; i < (int)e->tensors.size(); ++i) {
    Tensor& t = e->tensors[i];
    if (t.external || t.last < 0) continue;
    // A pixel of a 728-channel tensor is 1456 bytes: every 64-channel chunk of it (the 128-byte row of a TMA box, a warp's
    // 128-byte load or store) straddles two 128-byte lines and starts on a half sector for odd pixels.  Laid out with a pitch of
    // 768 channels every chunk is one aligned line; the 40 padding channels are never read or written (the tensor maps and the
    // kernels see C = 728).  Only for tensors that are never addressed as channel slices (concat buffers keep their layout).
    t.pitch = t.C;
    if (tuning().pad_pitch && t.C > 64 && (t.C & 63)) {
      bool sliced = false;
      for (const Step& st : e->steps)
        for (const Ref* r : {&st.in, &st.out, &st.res})
          if (r->t == i && (r->coff != 0 || r->C != t.C)) sliced = true;
      if (!sliced) t.pitch = (t.C + 63) & ~63;
    }
    t.bytes = (((size_t)t.H * t.W * t.pitch * 4 * e->max_batch) + 1023) & ~(size_t)1023;
    if (e->keep) { t.first = 0; t.last = 1 << 30; }
    order.push_back(i);
  }
  std::sort(order.begin(), order.end(), [&](int a, int b) { return e->tensors[a].bytes > e->tensors[b].bytes; });
  std::vector<int> placed;
  size_t total = 0;
  for (int id : order) {
    Tensor& t = e->tensors[id];
    std::vector<std::pair<size_t, size_t>> busy;
    for (int pid : placed) {
      const Tensor& p = e->tensors[pid];
      if (p.first <= t.last && t.first <= p.last) busy.push_back({p.offset, p.offset + p.bytes});
    }
    std::sort(busy.begin(), busy.end());
    size_t off = 0;
    for (auto& iv : busy) {
      if (off + t.bytes <= iv.first) break;
      off = std::max(off, iv.second);
    }
    t.offset = off;
    total = std::max(total, off + t.bytes);
    placed.push_back(id);
  }
  if (total > e->arena_bytes) {
    if (e->arena) cudaFree(e->arena);
    e->arena = nullptr; e->arena_bytes = 0;
    cudaError_t r = cudaMalloc(&e->arena, total);
    if (r != cudaSuccess) return fail(e, EMD_ENOMEM, "workspace of %.2f GB: %s", total / 1e9, cudaGetErrorString(r));
    e->arena_bytes = total;
    // Every page of a fresh workspace is written once before a kernel sees it.  Measured (tools/keep_repro.py): the FIRST pass
    // over a newly allocated workspace of >= 28 GB, whose first accesses are TMA loads / stores, intermittently (1 pass in 4)
    // ended in "unspecified launch failure" or returned a few wrong tiles; later passes over the same memory never did, 21 GB
    // never did, the CUDA-core kernels never did, and clearing the allocation first removed it (0 of 20).
    r = cudaMemset(e->arena, 0, total);
    if (r == cudaSuccess) r = cudaDeviceSynchronize();
    if (r != cudaSuccess) return fail(e, EMD_ECUDA, "clearing the workspace: %s", cudaGetErrorString(r));
  }
  return EMD_OK;
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
struct BlobHeader { char magic[8]; uint32_t n_entries, variant, reserved[4]; };            // 32 bytes
struct BlobRecord { char name[48]; uint32_t rows, cols; uint64_t offset; };               // 64 bytes

const float* entry_ptr(emd_engine* e, const std::string& name, uint32_t rows, uint32_t cols, std::string* why) {
  auto it = e->entries.find(name);
  if (it == e->entries.end()) { *why = "missing blob entry " + name; return nullptr; }
  if (it->second.rows != rows || it->second.cols != cols) {
    char b[160];
    snprintf(b, sizeof b, "blob entry %s is %ux%u, expected %ux%u", name.c_str(), it->second.rows, it->second.cols, rows, cols);
    *why = b; return nullptr;
  }
  return reinterpret_cast<const float*>(e->d_blob + it->second.offset);
}

int bind_weights(emd_engine* e, const char* host_blob) {
  std::string why;
  for (Step& s : e->steps) {
    switch (s.kind) {
      case SK_DW:
        if (!(s.dw = entry_ptr(e, s.wname + "/dw", 9, s.Cin, &why))) return fail(e, EMD_EINVAL, "%s", why.c_str());
        break;
      case SK_CONV:
      case SK_DECONV: {
        const uint32_t K = (uint32_t)(s.k * s.k * s.Cin);
        if (!(s.w = entry_ptr(e, s.wname + "/w", K, s.Cout, &why)) ||
            !(s.scale = entry_ptr(e, s.wname + "/scale", 1, s.Cout, &why)) ||
            !(s.shift = entry_ptr(e, s.wname + "/shift", 1, s.Cout, &why)))
          return fail(e, EMD_EINVAL, "%s", why.c_str());
        if (s.name == "aspp_image_conv") {  // conv + bias only; BN/ReLU6 come after the resize (DMG:338-345)
          if (!(s.scale = entry_ptr(e, "aspp_image/one", 1, s.Cout, &why)) ||
              !(s.shift = entry_ptr(e, "aspp_image/bias", 1, s.Cout, &why)))
            return fail(e, EMD_EINVAL, "%s", why.c_str());
        }
        s.scale0 = *reinterpret_cast<const float*>(host_blob + e->entries[s.wname + "/scale"].offset);
        s.shift0 = *reinterpret_cast<const float*>(host_blob + e->entries[s.wname + "/shift"].offset);
        // 16-bit operand copies in the tcgen05 tile layout
        const float* hw = reinterpret_cast<const float*>(host_blob + e->entries[s.wname + "/w"].offset);
        for (int et : {ET_BF16, ET_F16}) {
          {  // operand-precision copy for the CUDA-core fallback, so both kernels see the same operand values
            const size_t nw = (size_t)K * s.Cout;
            std::vector<float> r(nw);
            for (size_t i = 0; i < nw; ++i)
              r[i] = et == ET_BF16 ? __bfloat162float(__float2bfloat16_rn(hw[i])) : __half2float(__float2half_rn(hw[i]));
            void* d = nullptr;
            CU(e, cudaMalloc(&d, nw * 4));
            e->w16_allocs.push_back(d);
            CU(e, cudaMemcpy(d, r.data(), nw * 4, cudaMemcpyHostToDevice));
            s.wr[et] = reinterpret_cast<const float*>(d);
          }
          size_t nb = umma_pack_weights(hw, s.k * s.k, s.Cin, s.Cout, et, nullptr);
          if (!nb) continue;
          std::vector<char> tmp(nb);
          umma_pack_weights(hw, s.k * s.k, s.Cin, s.Cout, et, tmp.data());
          void* d = nullptr;
          CU(e, cudaMalloc(&d, nb));
          e->w16_allocs.push_back(d);
          CU(e, cudaMemcpy(d, tmp.data(), nb, cudaMemcpyHostToDevice));
          s.w16[et] = d;
        }
        break;
      }
      case SK_RESIZE:
        if (s.wname.empty()) break;
        if (s.wname == "aspp_image:post") {   // variant A: BN after the resized image-level conv (DMG:345)
          if (!(s.scale = entry_ptr(e, "aspp_image/scale", 1, s.Cout, &why)) ||
              !(s.shift = entry_ptr(e, "aspp_image/bnshift", 1, s.Cout, &why)))
            return fail(e, EMD_EINVAL, "%s", why.c_str());
        } else if (!(s.scale = entry_ptr(e, s.wname + "/scale", 1, s.Cout, &why)) ||
                   !(s.shift = entry_ptr(e, s.wname + "/shift", 1, s.Cout, &why))) {
          return fail(e, EMD_EINVAL, "%s", why.c_str());
        }
        break;
      default: break;
    }
  }
  return EMD_OK;
}

// ---------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------
struct Override { int t; void* ptr; };

struct ExecCtx {
  emd_engine* e;
  int et, n;
  cudaStream_t s;
  const float* d_in;  // network input (device, f32)
  float* d_out;       // network output (device, f32)
  std::vector<Override> ov;
  int b0 = 0;         // first crop of the batch this step works on (a slice of the batch: every view starts b0 images in)
  int sms = 0;        // > 0: this launch may use only that many SMs (a forked pair of kernels shares the GPU); 0 = all
  bool fork_ok = false;   // whole-network passes only: a step may launch its forked partner (run_step)
};

inline int sms_of(const ExecCtx& c) { return c.sms > 0 ? c.sms : c.e->num_sms; }

View make_view(const ExecCtx& c, Ref r) {
  const Tensor& t = c.e->tensors[r.t];
  View v; v.H = t.H; v.W = t.W; v.C = r.C;
  for (const Override& o : c.ov)
    if (o.t == r.t) { v.ptr = o.ptr; v.pitch = r.C; v.coff = 0; return v; }
  v.pitch = t.pitch; v.coff = r.coff;
  size_t esz = 4;     // network input and output are FP32 images
  if (r.t == c.e->t_input) v.ptr = const_cast<float*>(c.d_in);
  else if (r.t == c.e->t_output) v.ptr = c.d_out;
  else { v.ptr = c.e->arena + t.offset; esz = c.et == ET_F32 ? 4 : 2; }
  if (c.b0) v.ptr = reinterpret_cast<char*>(v.ptr) + (size_t)c.b0 * t.H * t.W * t.pitch * esz;
  return v;
}

cudaError_t run_conv(ExecCtx& c, ConvParams& p, const Step& s) {
  emd_engine* e = c.e;
  e->cnt.launches++;
  auto tc = [&](int kind, cudaError_t r) { e->cnt.umma++; e->cnt.kind[kind]++; return r; };
  if (c.et != ET_F32 && e->use_umma && final_tma_supported(p, c.et)) {
    p.w = s.wr[c.et];
    if (final_umma_supported(p, c.et)) return tc(LK_FINAL_UMMA, launch_final_umma(p, s.scale0, s.shift0, c.et, sms_of(c), c.s));
    e->cnt.kind[LK_FINAL_TMA]++;
    return launch_final_tma(p, s.scale0, s.shift0, c.et, sms_of(c), c.s);
  }
  if (c.et != ET_F32 && e->use_umma && s.w16[c.et] && fused_supported(p, c.et, nullptr)) {
    const cudaError_t r = launch_conv_fused(p, c.et, nullptr, sms_of(c), c.s);
    return tc(last_launch_kind(), r);
  }
  if (c.et != ET_F32 && e->use_umma && s.w16[c.et] && umma_supported(p, c.et))
    return tc(LK_UMMA_GEN1, launch_conv_umma(p, c.et, sms_of(c), c.s));
  if (c.et != ET_F32 && e->use_umma && tuning().strict) {   // a silent CUDA-core fallback would pass every test and cost 10x
    e->err = "strict: no tensor-core kernel supports step " + s.name;
    return cudaErrorNotSupported;
  }
  if (c.et != ET_F32 && s.wr[c.et]) p.w = s.wr[c.et];
  e->cnt.kind[LK_SIMT]++;
  return launch_conv_simt(p, c.et, c.s);
}

// ConvParams of a plain (non-stem) SK_CONV step
void conv_params(const ExecCtx& c, const Step& s, ConvParams& p) {
  emd_engine* e = c.e;
  const Tensor& ti = e->tensors[s.in.t];
  const Tensor& to = e->tensors[s.out.t];
  p.in = make_view(c, s.in); p.out = make_view(c, s.out);
  if (s.res.t >= 0) p.res = make_view(c, s.res);
  p.N = c.n; p.MH = to.H; p.MW = to.W;
  p.istride = s.stride; p.ostride = 1; p.oy0 = p.ox0 = 0;
  p.Cin = s.Cin; p.Cout = s.Cout; p.w = s.w; p.w16 = s.w16[c.et]; p.scale = s.scale; p.shift = s.shift;
  p.relu6 = s.relu6; p.clip01 = s.clip01; p.out_f32 = to.external; p.in_f32 = ti.external;
  p.wtaps = s.k * s.k;
  // taps: TF SAME. stride 1: symmetric (k/2)*rate.  stride 2 only occurs with k = 1 here (DMG:365-370).
  const int half = s.k / 2;
  p.ntaps = 0;
  for (int ky = 0; ky < s.k; ++ky)
    for (int kx = 0; kx < s.k; ++kx) {
      const int dy = (ky - half) * s.rate, dx = (kx - half) * s.rate;
      if (std::abs(dy) >= ti.H || std::abs(dx) >= ti.W) continue;  // tap never in bounds (dilation >= map)
      p.dy[p.ntaps] = dy; p.dx[p.ntaps] = dx; p.wrow[p.ntaps] = ky * s.k + kx; p.ntaps++;
    }
}

// A stride-1 separable block whose pointwise GEMM can take the depthwise 3x3 as its A-operand producer
// (emd_fused.cu, dw mode): the SK_DW step is skipped and the SK_CONV step launches the fused kernel.
bool dw_fusable(const ExecCtx& c, int conv_idx, ConvParams* out) {
  emd_engine* e = c.e;
  if (c.et == ET_F32 || !e->use_umma || conv_idx < 1 || conv_idx >= (int)e->steps.size()) return false;
  const Step& s = e->steps[conv_idx];
  const Step& d = e->steps[conv_idx - 1];
  if (s.kind != SK_CONV || s.k != 1 || s.stride != 1 || !s.w16[c.et]) return false;
  if (d.kind != SK_DW || d.layer != s.layer || d.stride != 1 || d.rate != 1 || d.Cin == 1) return false;
  if (e->tensors[d.in.t].external) return false;
  ConvParams p{};
  conv_params(c, s, p);
  p.in = make_view(c, d.in);
  p.in_f32 = 0;
  if (!fused_supported(p, c.et, d.dw)) return false;
  if (out) *out = p;
  return true;
}

int step_error(emd_engine* e, int idx, cudaError_t r) {
  if (r == cudaErrorNotSupported && tuning().strict)
    return fail(e, EMD_ESTATE, "step %s: no tensor-core kernel supports this shape and EMD_STRICT forbids the CUDA-core fallback", e->steps[idx].name.c_str());
  return fail(e, EMD_ECUDA, "step %s: %s", e->steps[idx].name.c_str(), cudaGetErrorString(r));
}

cudaError_t run_step_impl(ExecCtx& c, int idx);
// A separable block whose depthwise runs inside its GEMM kernel is bound by that kernel's math warps and leaves a third of the
// HBM bandwidth unused; the 1x1 conv that follows it in the decoder (residualN_d) reads the same tensor and is bound by HBM alone.
// With fork_sms = n the two run side by side: the block on num_sms - n SMs of the pass's stream, the 1x1 conv on n SMs of a
// second stream, joined before the next step (which adds the two).
static int fork_side_of(const ExecCtx& c, int idx) {
  emd_engine* e = c.e;
  const int n = tuning().fork_sms;
  if (!c.fork_ok || n < 8 || n > e->num_sms - 32 || e->profile || c.et == ET_F32 || !e->use_umma || c.sms || !e->fork_stream) return -1;
  if (idx >= (int)e->fork_side.size() || e->fork_side[idx] < 0) return -1;
  if (!dw_fusable(c, idx, nullptr)) return -1;
  return e->fork_side[idx];
}

cudaError_t run_step(ExecCtx& c, int idx) {
  emd_engine* e = c.e;
  Step& s = e->steps[idx];
  if (e->fork_skip == idx) { e->fork_skip = -1; return cudaSuccess; }   // ran beside the step before it
  const long long before = e->cnt.launches;
  s.fused = false;
  const int side = fork_side_of(c, idx);
  if (side < 0) {
    cudaError_t r = run_step_impl(c, idx);
    s.nlaunch = (int)(e->cnt.launches - before);
    return r;
  }
  cudaError_t r = cudaEventRecord(e->ev_fork, c.s);                     // everything both kernels read is complete here
  if (r != cudaSuccess) return r;
  c.sms = e->num_sms - tuning().fork_sms;
  r = run_step_impl(c, idx);
  c.sms = 0;
  s.nlaunch = (int)(e->cnt.launches - before);
  if (r != cudaSuccess) return r;
  if ((r = cudaStreamWaitEvent(e->fork_stream, e->ev_fork, 0)) != cudaSuccess) return r;
  ExecCtx c2 = c;
  c2.s = e->fork_stream; c2.sms = tuning().fork_sms;
  Step& s2 = e->steps[side];
  const long long before2 = e->cnt.launches;
  s2.fused = false;
  r = run_step_impl(c2, side);
  s2.nlaunch = (int)(e->cnt.launches - before2);
  if (r != cudaSuccess) return r;
  if ((r = cudaEventRecord(e->ev_join, e->fork_stream)) != cudaSuccess) return r;
  if ((r = cudaStreamWaitEvent(c.s, e->ev_join, 0)) != cudaSuccess) return r;
  e->fork_skip = side;
  return cudaSuccess;
}

cudaError_t run_step_impl(ExecCtx& c, int idx) {
  emd_engine* e = c.e;
  Step& s = e->steps[idx];
  const Tensor& ti = e->tensors[s.in.t];
  const Tensor& to = e->tensors[s.out.t];
  switch (s.kind) {
    case SK_DW: {
      if (s.Cin == 1) { s.fused = true; return cudaSuccess; }  // 1-channel stem: the depthwise is fused into the pointwise kernel below
      if (dw_fusable(c, idx + 1, nullptr)) { s.fused = true; return cudaSuccess; }  // computed inside the pointwise GEMM's producer warps
      DwParams p{};
      p.in = make_view(c, s.in); p.out = make_view(c, s.out);
      p.N = c.n; p.OH = to.H; p.OW = to.W; p.stride = s.stride; p.rate = s.rate;
      p.pad = (s.stride == 1) ? s.rate : 0;  // TF SAME: symmetric `rate` at stride 1; 0 before / 1 after at stride 2 on even sizes
      p.w = s.dw; p.in_f32 = ti.external;
      e->cnt.launches++;
      if (tuning().dw_reg_all && dw_reg_supported(p, c.et)) return launch_dw_reg(p, c.et, c.s);
      if (e->use_umma && dw_tma_supported(p, c.et)) return launch_dw_tma(p, c.et, sms_of(c), c.s);
      if (e->use_umma && dw_s2_tma_supported(p, c.et)) return launch_dw_s2_tma(p, c.et, sms_of(c), c.s);
      if (tuning().dw_reg && dw_reg_supported(p, c.et)) return launch_dw_reg(p, c.et, c.s);
      if (tuning().dw_tile && dw_tile_supported(p, c.et)) return launch_dw_tile(p, c.et, c.s);
      if (tuning().dw_strip && dw_strip_supported(p, c.et)) return launch_dw_strip(p, c.et, c.s);
      return launch_dw3x3(p, c.et, c.s);
    }
    case SK_POOL: {
      PoolParams p{}; p.in = make_view(c, s.in); p.out = make_view(c, s.out); p.N = c.n;
      e->cnt.launches++;
      return launch_avgpool(p, c.et, c.s);
    }
    case SK_RESIZE: {
      ResizeParams p{}; p.in = make_view(c, s.in); p.out = make_view(c, s.out); p.N = c.n;
      p.scale = s.scale; p.shift = s.shift; p.relu6 = s.relu6;
      e->cnt.launches++;
      return launch_resize(p, c.et, c.s);
    }
    case SK_CONV: {
      if (s.Cin == 1 && s.k == 1) {  // network stem (cnn0 / residual0): outer-product kernel on the FP32 input
        const Step* dws = (idx > 0 && e->steps[idx - 1].kind == SK_DW && e->steps[idx - 1].layer == s.layer) ? &e->steps[idx - 1] : nullptr;
        const Ref rin = dws ? dws->in : s.in;
        const Tensor& tin = e->tensors[rin.t];
        if (tin.external && !(s.Cout & 7) && !((s.Cout >> 3) & ((s.Cout >> 3) - 1))) {
          StemParams p{};
          View vin = make_view(c, rin);
          p.in = reinterpret_cast<const float*>(vin.ptr); p.IH = tin.H; p.IW = tin.W;
          p.out = make_view(c, s.out); p.N = c.n;
          p.dw = dws ? dws->dw : nullptr;
          p.w = (c.et != ET_F32 && s.wr[c.et]) ? s.wr[c.et] : s.w;
          p.scale = s.scale; p.shift = s.shift; p.istride = s.stride; p.relu6 = s.relu6;
          e->cnt.launches++;
          s.fused = dws != nullptr;
          return launch_stem(p, c.et, c.s);
        }
      }
      ConvParams p{};
      if (dw_fusable(c, idx, &p)) {
        s.fused = true;
        e->cnt.launches++;
        e->cnt.umma++;
        e->cnt.kind[LK_FUSED_DW]++;
        return launch_conv_fused(p, c.et, e->steps[idx - 1].dw, sms_of(c), c.s);
      }
      p = ConvParams{};
      conv_params(c, s, p);
      return run_conv(c, p, s);
    }
    case SK_DECONV: {
      // conv2d_transpose 3x3 stride 2 SAME = 4 sub-pixel phases (App. A.4):
      //   out[2j+0] = in[j] w[0] + in[j-1] w[2];  out[2j+1] = in[j] w[1]   (per axis)
      ConvParams ph[4];
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          ConvParams& p = ph[py * 2 + px];
          p = ConvParams{};
          p.in = make_view(c, s.in); p.out = make_view(c, s.out);
          p.N = c.n; p.MH = ti.H; p.MW = ti.W;
          p.istride = 1; p.ostride = 2; p.oy0 = py; p.ox0 = px;
          p.Cin = s.Cin; p.Cout = s.Cout; p.w = s.w; p.w16 = s.w16[c.et]; p.scale = s.scale; p.shift = s.shift;
          p.relu6 = s.relu6; p.wtaps = 9;
          const int kys[2][2] = {{0, 2}, {1, -1}}, dys[2][2] = {{0, -1}, {0, 0}};
          p.ntaps = 0;
          for (int a = 0; a < 2; ++a) {
            if (kys[py][a] < 0) continue;
            for (int b2 = 0; b2 < 2; ++b2) {
              if (kys[px][b2] < 0) continue;
              p.dy[p.ntaps] = dys[py][a]; p.dx[p.ntaps] = dys[px][b2];
              p.wrow[p.ntaps] = kys[py][a] * 3 + kys[px][b2]; p.ntaps++;
            }
          }
        }
      if (c.et != ET_F32 && e->use_umma && s.w16[c.et] && fused_multi_supported(ph, 4, c.et)) {
        e->cnt.launches++;
        e->cnt.umma++;
        const cudaError_t r = launch_conv_fused_multi(ph, 4, c.et, sms_of(c), c.s);   // one launch: work items = (input tile, phase)
        e->cnt.kind[last_launch_kind()]++;
        return r;
      }
      for (int v = 0; v < 4; ++v) {
        cudaError_t r = run_conv(c, ph[v], s);
        if (r != cudaSuccess) return r;
      }
      return cudaSuccess;
    }
  }
  return cudaErrorInvalidValue;
}

// debugging aid (option `poison`): every byte of the activation workspace is set to 0xFF (NaN in every element type) before a
// pass, so a layer that reads what this pass has not written -- padding channels, a halo outside its tensor, a buffer whose
// producer has not finished -- turns the result into NaNs instead of silently reusing the previous pass's identical values
int poison_arena(emd_engine* e, cudaStream_t s) {
  if (tuning().poison && e->arena) CU(e, cudaMemsetAsync(e->arena, 0xFF, e->arena_bytes, s));
  return EMD_OK;
}

int run_network_direct(emd_engine* e, const float* d_in, float* d_out, int n, int mode, cudaStream_t s) {
  { int prc = poison_arena(e, s); if (prc) return prc; }
  ExecCtx c{e, mode == EMD_MODE_FP32 ? ET_F32 : (mode == EMD_MODE_BF16 ? ET_BF16 : ET_F16), n, s, d_in, d_out, {}};
  c.fork_ok = true; e->fork_skip = -1;
  if (e->profile && e->events.size() < e->steps.size() + 1) {
    e->events.resize(e->steps.size() + 1);
    for (auto& ev : e->events) CU(e, cudaEventCreate(&ev));
  }
  e->last_n = n; e->last_et = c.et;
  for (size_t i = 0; i < e->steps.size(); ++i) {
    if (e->profile) CU(e, cudaEventRecord(e->events[i], s));
    cudaError_t r = run_step(c, (int)i);
    if (r != cudaSuccess) return step_error(e, (int)i, r);
  }
  if (e->profile) {
    CU(e, cudaEventRecord(e->events[e->steps.size()], s));
    CU(e, cudaStreamSynchronize(s));
    for (size_t i = 0; i < e->steps.size(); ++i) cudaEventElapsedTime(&e->steps[i].ms, e->events[i], e->events[i + 1]);
  }
  return EMD_OK;
}

// One pass over a large batch with HOST input and / or output.  Only the first layer reads the crops and only the last one
// writes the result, so those two run slice by slice -- the first on a slice as soon as its H2D copy has landed, the D2H
// copy of a slice as soon as the last layer has written it.  Around them, the high-resolution head of the network and its
// full-resolution tail (plan_arena: head_end / tail_start) run once per HALF of the batch, so the second half's upload
// hides under the first half's head and the first half's download under the second half's tail; everything in between
// runs once on the whole batch and keeps its full-batch efficiency.  What stays exposed is one crop's copy each way.
// h_in / h_out null = that side is already on the device (d_in / d_out used as given).
int run_network_sliced(emd_engine* e, const float* h_in, float* h_out, const float* d_in, float* d_out, int n, int mode,
                       cudaStream_t s, bool first_pass) {
  { int prc = poison_arena(e, s); if (prc) return prc; }
  ExecCtx c{e, mode == EMD_MODE_FP32 ? ET_F32 : (mode == EMD_MODE_BF16 ? ET_BF16 : ET_F16), n, s, d_in, d_out, {}};
  c.fork_ok = true; e->fork_skip = -1;
  const size_t per = (size_t)e->S * e->S;
  const Tuning& tn = tuning();
  const int k_env = tn.io_slices;
  const bool no_halves = !tn.halves;
  const int K = k_env > 0 && k_env <= emd_engine::kSlices ? k_env : emd_engine::kSlices;
  const int last = (int)e->steps.size() - 1;
  int stem = 1;         // the steps of the first layer (its depthwise and pointwise halves)
  while (stem < last && e->steps[stem].layer == e->steps[0].layer) ++stem;
  // phases: [0, stem) sliced, [stem, head_end] per half, (head_end, tail_start) whole batch, [tail_start, last) per half, last sliced
  const bool halves = !no_halves && e->head_end >= stem && e->tail_start <= last && n >= 16;
  const int head_end = halves ? e->head_end : stem - 1, tail_start = halves ? e->tail_start : last;
  const int parts_env = tn.io_parts;   // 2 (default) or 4 parts
  const int nh = halves ? (parts_env == 4 && n >= 32 ? 4 : 2) : 1;
  auto part_lo = [&](int h) { return (int)((long long)n * h / nh); };
  e->last_n = n; e->last_et = c.et;
  auto step = [&](int i, int b0, int nb) -> int {
    c.b0 = b0; c.n = nb;
    cudaError_t r = run_step(c, i);
    c.b0 = 0; c.n = n;
    return r == cudaSuccess ? EMD_OK : step_error(e, i, r);
  };
  // boundaries of up to k slices of [lo, hi) into b[]; single = 1: the first slice is one crop (the exposed upload),
  // single = 2: the last one (the exposed download); returns the number of slices
  auto slices = [](int lo, int hi, int k, int single, int* b) -> int {
    const int m = hi - lo;
    k = std::max(1, std::min(k, m));
    int nb = 0;
    b[nb++] = lo;
    if (k > 1) {
      const int body_lo = lo + (single == 1 ? 1 : 0), body_hi = hi - (single == 2 ? 1 : 0), parts = (single ? k - 1 : k);
      if (single == 1) b[nb++] = body_lo;
      for (int i = 1; i <= parts; ++i) {
        const int v = body_lo + (int)((long long)(body_hi - body_lo) * i / parts);
        if (v > b[nb - 1]) b[nb++] = v;
      }
    }
    if (b[nb - 1] != hi) b[nb++] = hi;
    return nb - 1;
  };
  int last_input_reader = stem - 1;   // the staging copy of the crops is dead after this step (a fused depthwise reads at the next one)
  for (int i = stem; i < last; ++i)
    if (e->steps[i].in.t == e->t_input) last_input_reader = std::min(i + 1, last - 1);
  int rc, ev = 0;
  if (h_in && !first_pass) CU(e, cudaStreamWaitEvent(e->copy_in, e->ev_in_free, 0));   // the pass before has read the staging buffer
  int bounds[emd_engine::kSlices + 2];
  if (h_in) {           // all uploads are queued up front, in order; the compute stream waits slice by slice
    for (int h = 0; h < nh; ++h) {
      const int lo = part_lo(h), hi = part_lo(h + 1);
      const int ns = slices(lo, hi, K / nh, h == 0 ? 1 : 0, bounds);
      for (int q = 0; q < ns; ++q, ++ev) {
        const int b0 = bounds[q], nb = bounds[q + 1] - b0;
        CU(e, cudaMemcpyAsync(const_cast<float*>(d_in) + b0 * per, h_in + b0 * per, nb * per * 4, cudaMemcpyHostToDevice, e->copy_in));
        CU(e, cudaEventRecord(e->ev_sl_in[ev], e->copy_in));
      }
    }
  }
  ev = 0;
  for (int h = 0; h < nh; ++h) {
    const int lo = part_lo(h), hi = part_lo(h + 1);
    if (h_in) {
      const int ns = slices(lo, hi, K / nh, h == 0 ? 1 : 0, bounds);
      for (int q = 0; q < ns; ++q, ++ev) {
        const int b0 = bounds[q], nb = bounds[q + 1] - b0;
        CU(e, cudaStreamWaitEvent(s, e->ev_sl_in[ev], 0));
        for (int i = 0; i < stem; ++i)
          if ((rc = step(i, b0, nb))) return rc;
      }
    } else {
      for (int i = 0; i < stem; ++i)
        if ((rc = step(i, lo, hi - lo))) return rc;
    }
    for (int i = stem; i <= head_end; ++i) {
      if ((rc = step(i, lo, hi - lo))) return rc;
      if (h_in && h == nh - 1 && i == last_input_reader) CU(e, cudaEventRecord(e->ev_in_free, s));
    }
  }
  if (h_in && last_input_reader < stem) CU(e, cudaEventRecord(e->ev_in_free, s));
  // the whole-batch middle section (the 32x32-resolution trunk: ~90 small kernels) is replayed from a CUDA graph from its
  // second use on, like the device-resident pass; only when nothing but kernel launches happens inside it
  bool replayed = false;
  bool mid_ok = halves && tn.graphs && tn.mid_graph && last_input_reader <= head_end;   // mid_graph: +0.2..0.9 % end to end
  for (int i = head_end + 1; i < tail_start && mid_ok; ++i)          // no caller-owned buffer inside the captured range
    for (int t : {e->steps[i].in.t, e->steps[i].out.t, e->steps[i].res.t})
      if (t == e->t_input || t == e->t_output) mid_ok = false;
  if (mid_ok) {
    GraphSlot& g = e->mid_graphs[n * 4 + mode];
    if (g.calls++ >= 1 && !g.failed) {
      if (!g.exec) {
        cudaGraph_t graph = nullptr;
        cudaError_t r = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        int crc = EMD_OK;
        if (r == cudaSuccess) {
          const Counts c0 = e->cnt;
          for (int i = head_end + 1; i < tail_start && crc == EMD_OK; ++i) crc = step(i, 0, n);
          g.cnt = e->cnt - c0;
          e->cnt = c0;
          r = cudaStreamEndCapture(s, &graph);
        }
        if (r != cudaSuccess || crc != EMD_OK || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
          cudaGetLastError();
          g.failed = true; g.exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
      }
      if (g.exec) {
        CU(e, cudaGraphLaunch(g.exec, s));
        e->cnt += g.cnt;
        replayed = true;
      }
    }
  }
  for (int i = head_end + 1; i < tail_start && !replayed; ++i) {
    if ((rc = step(i, 0, n))) return rc;
    if (h_in && i == last_input_reader) CU(e, cudaEventRecord(e->ev_in_free, s));
  }
  if (h_out && !first_pass) CU(e, cudaStreamWaitEvent(s, e->ev_out_free, 0));         // the D2H copies of the pass before have drained
  ev = 0;
  for (int h = 0; h < nh; ++h) {
    const int lo = part_lo(h), hi = part_lo(h + 1);
    for (int i = tail_start; i < last; ++i) {
      if ((rc = step(i, lo, hi - lo))) return rc;
      if (h_in && i == last_input_reader) CU(e, cudaEventRecord(e->ev_in_free, s));   // (only if the tail still read the crops)
    }
    if (h_out) {
      const int nsl = slices(lo, hi, K / nh, h == nh - 1 ? 2 : 0, bounds);   // the very last download is one crop
      for (int q = 0; q < nsl; ++q, ++ev) {
        const int b0 = bounds[q], nb = bounds[q + 1] - b0;
        if ((rc = step(last, b0, nb))) return rc;
        CU(e, cudaEventRecord(e->ev_sl_fin[ev], s));
        CU(e, cudaStreamWaitEvent(e->copy_out, e->ev_sl_fin[ev], 0));
        CU(e, cudaMemcpyAsync(h_out + b0 * per, d_out + b0 * per, nb * per * 4, cudaMemcpyDeviceToHost, e->copy_out));
      }
    } else if ((rc = step(last, lo, hi - lo))) return rc;
  }
  if (h_out) CU(e, cudaEventRecord(e->ev_out_free, e->copy_out));
  return EMD_OK;
}

// The ~125-kernel chain of one pass is replayed from a CUDA graph for small batches, where it is launch-bound (a
// 512x512 crop at batch 1 is ~1.6 ms of mostly launch gaps).  A graph is captured per (batch, mode) on that shape's second
// pass (the first runs directly so every kernel's attributes are set outside capture); the network reads / writes fixed
// staging buffers inside the graph, with a device-to-device copy of the crops either side.
int run_network(emd_engine* e, const float* d_in, float* d_out, int n, int mode, cudaStream_t s) {
  if (!tuning().graphs || e->profile || e->keep || n > tuning().graph_max_n) return run_network_direct(e, d_in, d_out, n, mode, s);
  const int key = n * 4 + mode;
  GraphSlot& g = e->graphs[key];
  const size_t bytes = (size_t)n * e->S * e->S * sizeof(float);
  if (g.calls++ == 0 || g.failed) return run_network_direct(e, d_in, d_out, n, mode, s);
  if (!g.exec) {
    if (!e->g_in) {
      const size_t cap = (size_t)e->max_batch * e->S * e->S * sizeof(float);
      CU(e, cudaMalloc(&e->g_in, cap));
      CU(e, cudaMalloc(&e->g_out, cap));
    }
    cudaGraph_t graph = nullptr;
    cudaError_t r = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    int rc = EMD_OK;
    if (r == cudaSuccess) {
      const Counts c0 = e->cnt;
      rc = run_network_direct(e, e->g_in, e->g_out, n, mode, s);
      g.cnt = e->cnt - c0;
      e->cnt = c0;                                  // nothing ran yet: the replay below counts them
      r = cudaStreamEndCapture(s, &graph);
    }
    if (r != cudaSuccess || rc != EMD_OK || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      g.failed = true; g.exec = nullptr;
      if (graph) cudaGraphDestroy(graph);
      return run_network_direct(e, d_in, d_out, n, mode, s);
    }
    cudaGraphDestroy(graph);
  }
  { int prc = poison_arena(e, s); if (prc) return prc; }
  CU(e, cudaMemcpyAsync(e->g_in, d_in, bytes, cudaMemcpyDeviceToDevice, s));
  CU(e, cudaGraphLaunch(g.exec, s));
  CU(e, cudaMemcpyAsync(d_out, e->g_out, bytes, cudaMemcpyDeviceToDevice, s));
  e->last_n = n; e->last_et = mode == EMD_MODE_FP32 ? ET_F32 : (mode == EMD_MODE_BF16 ? ET_BF16 : ET_F16);
  e->cnt += g.cnt;
  e->graph_replays++;
  return EMD_OK;
}

int check_mode(emd_engine* e, int mode) {
  if (mode != EMD_MODE_FP32 && mode != EMD_MODE_BF16 && mode != EMD_MODE_FP16) return fail(e, EMD_EINVAL, "bad mode %d", mode);
  if (!e->weights_loaded) return fail(e, EMD_ESTATE, "emd_load_weights has not been called");
  return EMD_OK;
}

template <typename T>
int grow(emd_engine* e, T** p, size_t* have, size_t need) {
  if (*have >= need) return EMD_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  cudaError_t r = cudaMalloc(reinterpret_cast<void**>(p), need);
  if (r != cudaSuccess) return fail(e, EMD_ENOMEM, "cudaMalloc(%zu): %s", need, cudaGetErrorString(r));
  // written once before any TMA access (plan_arena); the memset runs on the legacy stream, which the engine's non-blocking
  // streams do not wait for: finish it here (allocation is rare)
  if ((r = cudaMemset(*p, 0, need)) != cudaSuccess || (r = cudaDeviceSynchronize()) != cudaSuccess)
    return fail(e, EMD_ECUDA, "clearing a new buffer: %s", cudaGetErrorString(r));
  *have = need;
  return EMD_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int emd_version(void) { return 100; }

const char* emd_last_error(const emd_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int emd_create(emd_engine** out, int device, int cropsize, int variant, int max_batch) {
  if (!out) return fail(nullptr, EMD_EINVAL, "out is NULL");
  *out = nullptr;
  if (cropsize < 32 || cropsize % 32) return fail(nullptr, EMD_EINVAL, "cropsize %d: must be a multiple of 32 (>= 32)", cropsize);
  if (variant != EMD_VARIANT_A && variant != EMD_VARIANT_B) return fail(nullptr, EMD_EINVAL, "variant %d (EMD_VARIANT_A or EMD_VARIANT_B)", variant);
  if (max_batch < 1) return fail(nullptr, EMD_EINVAL, "max_batch %d", max_batch);
  int ndev = 0;
  cudaError_t r = cudaGetDeviceCount(&ndev);
  if (r != cudaSuccess || ndev == 0)
    return fail(nullptr, EMD_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                r != cudaSuccess ? cudaGetErrorString(r) : "device count 0");
  if (device < 0 || device >= ndev) return fail(nullptr, EMD_EINVAL, "device %d of %d", device, ndev);
  emd_engine* e = new emd_engine();
  e->device = device; e->S = cropsize; e->variant = variant; e->max_batch = max_batch;
  e->use_umma = tuning().umma != 0;
#define CUC(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t _r = (call);                                                                            \
    if (_r != cudaSuccess) {                                                                            \
      int code = fail(nullptr, EMD_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_r));              \
      emd_destroy(e);                                                                                   \
      return code;                                                                                      \
    }                                                                                                   \
  } while (0)
  CUC(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    int code = fail(nullptr, EMD_ECUDA, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major, prop.minor);
    emd_destroy(e);
    return code;
  }
  e->num_sms = prop.multiProcessorCount;
  CUC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&e->fork_stream, cudaStreamNonBlocking));
  CUC(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  build_schedule(e);
  annotate_work(e);
  // forked pairs: [i-1] depthwise step, [i] its GEMM step, [i+1] a 1x1 conv without residual operand on the same input tensor
  e->fork_side.assign(e->steps.size(), -1);
  for (int i = 1; i + 1 < (int)e->steps.size(); ++i) {
    const Step &d = e->steps[i - 1], &m = e->steps[i], &sd = e->steps[i + 1];
    if (d.kind == SK_DW && m.kind == SK_CONV && d.layer == m.layer && m.res.t < 0 && sd.kind == SK_CONV && sd.k == 1 && sd.stride == 1 &&
        sd.res.t < 0 && sd.in.t == d.in.t && sd.out.t != m.out.t && sd.out.t != d.in.t && !e->tensors[sd.out.t].external)
      e->fork_side[i] = i + 1;
  }
  int rc = plan_arena(e);
  if (rc != EMD_OK) { g_create_error = e->err; emd_destroy(e); return rc; }
  const size_t io = (size_t)max_batch * cropsize * cropsize * sizeof(float);
  CUC(cudaMalloc(&e->d_stage_in, io));
  CUC(cudaMalloc(&e->d_stage_out, io));
  CUC(cudaMalloc(&e->d_minmax, 2 * sizeof(double)));
  CUC(cudaMalloc(&e->d_partial, minmax_partial_bytes()));
  CUC(cudaMalloc(&e->d_origins, 512 * sizeof(int)));
#undef CUC
  *out = e;
  return EMD_OK;
}

int emd_destroy(emd_engine* e) {
  if (!e) return EMD_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void* p : e->w16_allocs) cudaFree(p);
  for (void* p : {(void*)e->arena, (void*)e->d_blob, (void*)e->d_stage_in, (void*)e->d_stage_out, e->d_img_raw, (void*)e->d_q_a, (void*)e->d_q_b, (void*)e->d_q_partial,
                  (void*)e->d_img, (void*)e->d_crops, (void*)e->d_tiles, (void*)e->d_sout, (void*)e->d_minmax,
                  e->d_partial, (void*)e->d_origins})
    if (p) cudaFree(p);
  for (auto& sl : e->slots) {
    for (void* p : {(void*)sl.d_raw, (void*)sl.d_norm, (void*)sl.d_crops, (void*)sl.d_tiles, (void*)sl.d_sout, (void*)sl.d_minmax, (void*)sl.d_partial})
      if (p) cudaFree(p);
    for (cudaEvent_t ev : {sl.ev_pre, sl.ev_net, sl.ev_post}) if (ev) cudaEventDestroy(ev);
  }
  if (e->s_pre) { cudaStreamDestroy(e->s_pre); cudaStreamDestroy(e->s_post); cudaEventDestroy(e->ev_img_start); }
  drop_graphs(e);
  if (e->g_in) { cudaFree(e->g_in); cudaFree(e->g_out); }
  for (auto ev : e->events) cudaEventDestroy(ev);
  if (e->copy_in) {
    cudaStreamDestroy(e->copy_in); cudaStreamDestroy(e->copy_out);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(e->ev_in[i]); cudaEventDestroy(e->ev_done[i]); cudaEventDestroy(e->ev_out[i]); }
    cudaEventDestroy(e->ev_start);
    for (int i = 0; i < emd_engine::kSlices; ++i) { cudaEventDestroy(e->ev_sl_in[i]); cudaEventDestroy(e->ev_sl_fin[i]); }
    cudaEventDestroy(e->ev_in_free); cudaEventDestroy(e->ev_out_free);
  }
  if (e->fork_stream) cudaStreamDestroy(e->fork_stream);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return EMD_OK;
}

int emd_load_weights(emd_engine* e, const void* blob, size_t nbytes) {
  if (!e || !blob) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  if (nbytes < sizeof(BlobHeader)) return fail(e, EMD_EINVAL, "blob too small");
  const char* b = reinterpret_cast<const char*>(blob);
  BlobHeader h; memcpy(&h, b, sizeof h);
  if (memcmp(h.magic, "EMDW0001", 8) != 0) return fail(e, EMD_EINVAL, "bad blob magic");
  if ((int)h.variant != e->variant) return fail(e, EMD_EINVAL, "blob is variant %u, engine is %d", h.variant, e->variant);
  if ((nbytes - sizeof h) / sizeof(BlobRecord) < h.n_entries) return fail(e, EMD_EINVAL, "blob truncated");
  // parse into a fresh table first: a malformed blob leaves the engine exactly as it was
  std::map<std::string, BlobEntry> entries;
  for (uint32_t i = 0; i < h.n_entries; ++i) {
    BlobRecord r; memcpy(&r, b + sizeof h + (size_t)i * sizeof r, sizeof r);
    r.name[47] = 0;
    const uint64_t elems = (uint64_t)r.rows * r.cols;                    // < 2^64: both factors are 32-bit
    if ((r.offset & 3) || r.offset > nbytes || elems > (nbytes - r.offset) / 4)
      return fail(e, EMD_EINVAL, "blob entry %s out of range or misaligned", r.name);
    entries[r.name] = BlobEntry{r.rows, r.cols, (size_t)r.offset};
  }
  // from here on the old weights are gone: the engine is unusable until the new ones are bound
  { int drc = drain_chain(e); if (drc) return drc; }
  CU(e, cudaStreamSynchronize(e->stream));
  e->weights_loaded = false;
  drop_graphs(e);
  for (Step& s : e->steps) {
    s.w = s.scale = s.shift = s.dw = nullptr;
    for (int t = 0; t < 3; ++t) { s.w16[t] = nullptr; s.wr[t] = nullptr; }
  }
  e->entries.swap(entries);
  if (e->d_blob) { cudaFree(e->d_blob); e->d_blob = nullptr; }
  for (void* p : e->w16_allocs) cudaFree(p);
  e->w16_allocs.clear();
  CU(e, cudaMalloc(&e->d_blob, nbytes));
  CU(e, cudaMemcpy(e->d_blob, blob, nbytes, cudaMemcpyHostToDevice));
  e->blob_bytes = nbytes;
  int rc = bind_weights(e, b);
  if (rc != EMD_OK) return rc;
  e->weights_loaded = true;
  return EMD_OK;
}

static int forward_impl(emd_engine* e, const float* crops, int n, float* out, int mode, void* stream, bool async) {
  if (!e || !crops || !out || n < 0) return EMD_EINVAL;
  int rc = check_mode(e, mode);
  if (rc) return rc;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const bool in_dev = is_device_ptr(crops), out_dev = is_device_ptr(out);
  const size_t per = (size_t)e->S * e->S;
  if (e->chained && (s != e->chain_stream || (in_dev && out_dev)) && (rc = drain_chain(e))) return rc;
  if (in_dev && out_dev) {
    for (int c0 = 0; c0 < n; c0 += e->max_batch) {
      const int nb = std::min(e->max_batch, n - c0);
      if ((rc = run_network(e, crops + c0 * per, out + c0 * per, nb, mode, s))) return rc;
    }
    return EMD_OK;
  }
  // Host buffers: the batch goes through in chunks of half the workspace so that the H2D copy of chunk i+1 and the
  // D2H copy of chunk i-1 (copy streams) overlap the network pass of chunk i (compute stream).  The two halves of the
  // staging buffers alternate; events order reuse.
  const int ch = e->max_batch >= 16 ? e->max_batch / 2 : e->max_batch;
  const int nslots = e->max_batch >= 16 ? 2 : 1;
  const int nchunks = (n + ch - 1) / ch;
  if (!e->copy_in) {
    CU(e, cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking));
    CU(e, cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CU(e, cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
      CU(e, cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
      CU(e, cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
    }
    CU(e, cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming));
    for (int i = 0; i < emd_engine::kSlices; ++i) {
      CU(e, cudaEventCreateWithFlags(&e->ev_sl_in[i], cudaEventDisableTiming));
      CU(e, cudaEventCreateWithFlags(&e->ev_sl_fin[i], cudaEventDisableTiming));
    }
    CU(e, cudaEventCreateWithFlags(&e->ev_in_free, cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&e->ev_out_free, cudaEventDisableTiming));
  }
  const bool chained = e->chained;                 // pipelining behind an outstanding async call: its events already order the staging buffers
  if (!chained) {
    CU(e, cudaEventRecord(e->ev_start, s));        // work already queued on the caller's stream comes first
    CU(e, cudaStreamWaitEvent(e->copy_in, e->ev_start, 0));
    CU(e, cudaStreamWaitEvent(e->copy_out, e->ev_start, 0));
  }
  const bool sliceable = tuning().sliced_io /* A/B switch: the two-chunk pipeline below */ && !e->profile && !e->keep && e->steps.size() >= 3 &&
                         e->steps.front().in.t == e->t_input && e->steps.back().out.t == e->t_output && e->steps.back().in.t != e->t_input;
  if (sliceable && n >= 16 && e->max_batch >= 16) {
    // passes of up to max_batch crops (balanced, so that no pass is a small remainder)
    const int npass = (n + e->max_batch - 1) / e->max_batch, pb = (n + npass - 1) / npass;
    for (int c0 = 0, ip = 0; c0 < n; c0 += pb, ++ip) {
      const int nb = std::min(pb, n - c0);
      if ((rc = run_network_sliced(e, in_dev ? nullptr : crops + c0 * per, out_dev ? nullptr : out + c0 * per,
                                   in_dev ? crops + c0 * per : e->d_stage_in, out_dev ? out + c0 * per : e->d_stage_out, nb, mode, s,
                                   ip == 0 && !chained)))
        return rc;
    }
    if (async) { e->chained = true; e->chain_stream = s; return EMD_OK; }        // emd_synchronize waits
    CU(e, cudaStreamSynchronize(s));
    if (!out_dev) CU(e, cudaStreamSynchronize(e->copy_out));
    e->chained = false;
    return EMD_OK;
  }
  if (chained && (rc = drain_chain(e))) return rc;   // a different path follows an async chain: drain it first
  for (int i = 0; i < nchunks; ++i) {
    const int c0 = i * ch, nb = std::min(ch, n - c0), slot = i % nslots;
    const float* d_in = crops + c0 * per;
    float* d_out = out + c0 * per;
    if (!in_dev) {
      float* stage = e->d_stage_in + (size_t)slot * ch * per;
      if (i >= nslots) CU(e, cudaStreamWaitEvent(e->copy_in, e->ev_done[slot], 0));   // pass i - nslots has read this slot
      CU(e, cudaMemcpyAsync(stage, d_in, nb * per * 4, cudaMemcpyHostToDevice, e->copy_in));
      CU(e, cudaEventRecord(e->ev_in[slot], e->copy_in));
      CU(e, cudaStreamWaitEvent(s, e->ev_in[slot], 0));
      d_in = stage;
    }
    if (!out_dev) {
      d_out = e->d_stage_out + (size_t)slot * ch * per;
      if (i >= nslots) CU(e, cudaStreamWaitEvent(s, e->ev_out[slot], 0));              // D2H of pass i - nslots has drained this slot
    }
    if ((rc = run_network(e, d_in, d_out, nb, mode, s))) return rc;
    CU(e, cudaEventRecord(e->ev_done[slot], s));
    if (!out_dev) {
      CU(e, cudaStreamWaitEvent(e->copy_out, e->ev_done[slot], 0));
      CU(e, cudaMemcpyAsync(out + c0 * per, d_out, nb * per * 4, cudaMemcpyDeviceToHost, e->copy_out));
      CU(e, cudaEventRecord(e->ev_out[slot], e->copy_out));
    }
  }
  CU(e, cudaStreamSynchronize(s));
  if (!out_dev) CU(e, cudaStreamSynchronize(e->copy_out));
  return EMD_OK;
}

int emd_forward(emd_engine* e, const float* crops, int n, float* out, int mode, void* stream) {
  return forward_impl(e, crops, n, out, mode, stream, false);
}

int emd_forward_async(emd_engine* e, const float* crops, int n, float* out, int mode, void* stream) {
  return forward_impl(e, crops, n, out, mode, stream, true);
}

int emd_synchronize(emd_engine* e, void* stream) {
  if (!e) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  int rc = drain_chain(e);
  if (rc) return rc;
  CU(e, cudaStreamSynchronize(s));
  if (e->copy_out) CU(e, cudaStreamSynchronize(e->copy_out));
  return EMD_OK;
}

int emd_plan_tiles(int H, int W, int crop, int overlap, int* ys, int* xs, int* ny, int* nx) {
  if (!ys || !xs || !ny || !nx || crop <= 0 || overlap < 0 || overlap >= crop || H < crop || W < crop) return EMD_EINVAL;
  const int sizes[2] = {H, W};
  int* outs[2] = {ys, xs};
  int* counts[2] = {ny, nx};
  for (int a = 0; a < 2; ++a) {
    const int num = sizes[a] / (crop - overlap) + 1;          // DEN:661-662
    const double len = (double)sizes[a] / (double)num;         // DEN:663-664 (true division)
    for (int i = 0; i < num; ++i) {
      int o = (int)std::nearbyint((double)i * len);            // np.round: half-to-even (App. D-2)
      outs[a][i] = std::min(o, sizes[a] - crop);               // clamp (App. D-3)
    }
    *counts[a] = num;
  }
  return EMD_OK;
}

int emd_normalise(emd_engine* e, const void* img, int in_f64, int H, int W, float* out, void* stream) {
  if (!e || !img || !out || H <= 0 || W <= 0) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const size_t n = (size_t)H * W, esz = in_f64 ? 8 : 4;
  const bool in_dev = is_device_ptr(img), out_dev = is_device_ptr(out);
  const void* d_src = img;
  if (!in_dev) {
    int rc = grow(e, reinterpret_cast<char**>(&e->d_img_raw), &e->img_raw_bytes, n * esz);
    if (rc) return rc;
    CU(e, cudaMemcpyAsync(e->d_img_raw, img, n * esz, cudaMemcpyHostToDevice, s));
    d_src = e->d_img_raw;
  }
  float* d_dst = out;
  if (!out_dev) {
    int rc = grow(e, &e->d_img, &e->img_bytes, n * 4);
    if (rc) return rc;
    d_dst = e->d_img;
  }
  CU(e, launch_minmax(d_src, in_f64, n, e->d_minmax, e->d_partial, s));
  CU(e, launch_normalise_apply(d_src, in_f64, n, e->d_minmax, d_dst, s));
  e->cnt.launches += 3;
  if (!out_dev) {
    CU(e, cudaMemcpyAsync(out, d_dst, n * 4, cudaMemcpyDeviceToHost, s));
    CU(e, cudaStreamSynchronize(s));
  }
  return EMD_OK;
}

static int upload_origins(emd_engine* e, const int* ys, const int* xs, int ny, int nx, cudaStream_t s) {
  if (ny < 1 || nx < 1 || ny > 256 || nx > 256) return fail(e, EMD_EINVAL, "tile grid %dx%d", ny, nx);
  CU(e, cudaMemcpyAsync(e->d_origins, ys, ny * sizeof(int), cudaMemcpyHostToDevice, s));
  CU(e, cudaMemcpyAsync(e->d_origins + 256, xs, nx * sizeof(int), cudaMemcpyHostToDevice, s));
  return EMD_OK;
}

int emd_gather_crops(emd_engine* e, const float* img, int H, int W, const int* ys, const int* xs, int ny, int nx,
                     int crop, float* crops, void* stream) {
  if (!e || !img || !ys || !xs || !crops || crop % 4) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  int rc = upload_origins(e, ys, xs, ny, nx, s);
  if (rc) return rc;
  const size_t nimg = (size_t)H * W * 4, ncr = (size_t)ny * nx * crop * crop * 4;
  const bool in_dev = is_device_ptr(img), out_dev = is_device_ptr(crops);
  const float* d_src = img;
  if (!in_dev) {
    if ((rc = grow(e, &e->d_img, &e->img_bytes, nimg))) return rc;
    CU(e, cudaMemcpyAsync(e->d_img, img, nimg, cudaMemcpyHostToDevice, s));
    d_src = e->d_img;
  }
  float* d_dst = crops;
  if (!out_dev) {
    if ((rc = grow(e, &e->d_crops, &e->crops_bytes, ncr))) return rc;
    d_dst = e->d_crops;
  }
  CU(e, launch_gather(d_src, H, W, e->d_origins, e->d_origins + 256, ny, nx, crop, d_dst, s));
  e->cnt.launches++;
  if (!out_dev) {
    CU(e, cudaMemcpyAsync(crops, d_dst, ncr, cudaMemcpyDeviceToHost, s));
    CU(e, cudaStreamSynchronize(s));
  }
  return EMD_OK;
}

int emd_stitch(emd_engine* e, const float* tiles, const int* ys, const int* xs, int ny, int nx, int crop, int H, int W,
               int clip, double* out, void* stream) {
  if (!e || !tiles || !ys || !xs || !out) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  int rc = upload_origins(e, ys, xs, ny, nx, s);
  if (rc) return rc;
  const size_t ntl = (size_t)ny * nx * crop * crop * 4, nout = (size_t)H * W * 8;
  const bool in_dev = is_device_ptr(tiles), out_dev = is_device_ptr(out);
  const float* d_src = tiles;
  if (!in_dev) {
    if ((rc = grow(e, &e->d_tiles, &e->tiles_bytes, ntl))) return rc;
    CU(e, cudaMemcpyAsync(e->d_tiles, tiles, ntl, cudaMemcpyHostToDevice, s));
    d_src = e->d_tiles;
  }
  double* d_dst = out;
  if (!out_dev) {
    if ((rc = grow(e, &e->d_sout, &e->sout_bytes, nout))) return rc;
    d_dst = e->d_sout;
  }
  CU(e, launch_stitch(d_src, e->d_origins, e->d_origins + 256, ny, nx, crop, H, W, clip, d_dst, 0, s));
  e->cnt.launches++;
  if (!out_dev) {
    CU(e, cudaMemcpyAsync(out, d_dst, nout, cudaMemcpyDeviceToHost, s));
    CU(e, cudaStreamSynchronize(s));
  }
  return EMD_OK;
}

int emd_quality(emd_engine* e, const float* a, const float* b, int n, int H, int W, double* out, void* stream) {
  if (!e || !a || !b || !out || n < 0) return EMD_EINVAL;
  if (H < 11 || W < 11) return fail(e, EMD_EINVAL, "images of %dx%d are smaller than the 11x11 SSIM window", H, W);
  if (n == 0) return EMD_OK;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const size_t bytes = (size_t)n * H * W * sizeof(float);
  int rc;
  const float *da = a, *db = b;
  if (!is_device_ptr(a)) {
    if ((rc = grow(e, &e->d_q_a, &e->q_a_bytes, bytes))) return rc;
    CU(e, cudaMemcpyAsync(e->d_q_a, a, bytes, cudaMemcpyHostToDevice, s));
    da = reinterpret_cast<const float*>(e->d_q_a);
  }
  if (!is_device_ptr(b)) {
    if ((rc = grow(e, &e->d_q_b, &e->q_b_bytes, bytes))) return rc;
    CU(e, cudaMemcpyAsync(e->d_q_b, b, bytes, cudaMemcpyHostToDevice, s));
    db = reinterpret_cast<const float*>(e->d_q_b);
  }
  const size_t pbytes = quality_partial_bytes(n, H, W) + (size_t)n * 3 * sizeof(double);
  if ((rc = grow(e, &e->d_q_partial, &e->q_partial_bytes, pbytes))) return rc;
  double* d_partial = reinterpret_cast<double*>(e->d_q_partial);
  double* d_out = d_partial + quality_partial_bytes(n, H, W) / sizeof(double);
  CU(e, launch_quality(da, db, n, H, W, d_partial, d_out, s));
  e->cnt.launches += 2;
  CU(e, cudaMemcpyAsync(out, d_out, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(e, cudaStreamSynchronize(s));
  return EMD_OK;
}

// Core of the whole-micrograph path.  Three streams: `pre` (upload, min/max, normalise, tile gather), the caller's / engine's
// compute stream `s` (network), `post` (stitch, download); image i uses slot i & 1.  One image = the same chain without overlap.
static int denoise_images(emd_engine* e, const void* const* imgs, int count, int H, int W, int overlap, int flags, int mode,
                          void* const* outs, cudaStream_t s) {
  { int drc = drain_chain(e); if (drc) return drc; }
  const int crop = e->S;
  if (H < crop || W < crop) return fail(e, EMD_EINVAL, "image %dx%d smaller than the %d crop", H, W, crop);
  if (overlap < 0 || overlap >= crop) return fail(e, EMD_EINVAL, "overlap %d outside [0,%d)", overlap, crop);
  const bool f64 = flags & EMD_FLAG_INPUT_F64, out_f32 = flags & EMD_FLAG_OUTPUT_F32;
  if (f64 && !(flags & EMD_FLAG_PREPROCESS)) return fail(e, EMD_EINVAL, "float64 input needs EMD_FLAG_PREPROCESS");
  for (int i = 0; i < count; ++i)
    if (!imgs[i] || !outs[i]) return fail(e, EMD_EINVAL, "image %d: NULL pointer", i);
  int ys[256], xs[256], ny = 0, nx = 0;
  if (H / (crop - overlap) + 1 > 256 || W / (crop - overlap) + 1 > 256) return fail(e, EMD_EINVAL, "image too large");
  if (emd_plan_tiles(H, W, crop, overlap, ys, xs, &ny, &nx)) return fail(e, EMD_EINVAL, "bad tiling arguments");
  int rc;
  if ((rc = upload_origins(e, ys, xs, ny, nx, s))) return rc;
  if (!e->s_pre) {
    CU(e, cudaStreamCreateWithFlags(&e->s_pre, cudaStreamNonBlocking));
    CU(e, cudaStreamCreateWithFlags(&e->s_post, cudaStreamNonBlocking));
    CU(e, cudaEventCreateWithFlags(&e->ev_img_start, cudaEventDisableTiming));
    for (auto& sl : e->slots) {
      CU(e, cudaEventCreateWithFlags(&sl.ev_pre, cudaEventDisableTiming));
      CU(e, cudaEventCreateWithFlags(&sl.ev_net, cudaEventDisableTiming));
      CU(e, cudaEventCreateWithFlags(&sl.ev_post, cudaEventDisableTiming));
    }
  }
  const size_t npx = (size_t)H * W, esz = f64 ? 8 : 4, osz = out_f32 ? 4 : 8;
  const int T = ny * nx;
  const size_t per = (size_t)crop * crop, ncr = (size_t)T * per * 4;
  const int nslots = count > 1 ? 2 : 1;
  for (int k = 0; k < nslots; ++k) {          // all allocation up front: cudaMalloc / cudaFree would serialise the streams
    auto& sl = e->slots[k];
    if ((rc = grow(e, &sl.d_raw, &sl.raw_bytes, npx * esz))) return rc;
    if ((rc = grow(e, &sl.d_norm, &sl.norm_bytes, npx * 4))) return rc;
    if ((rc = grow(e, &sl.d_crops, &sl.crops_bytes, ncr))) return rc;
    if ((rc = grow(e, &sl.d_tiles, &sl.tiles_bytes, ncr))) return rc;
    if ((rc = grow(e, &sl.d_sout, &sl.sout_bytes, npx * 8))) return rc;
    if ((rc = grow(e, reinterpret_cast<char**>(&sl.d_minmax), &sl.mm_bytes, 2 * sizeof(double)))) return rc;
    if ((rc = grow(e, &sl.d_partial, &sl.partial_bytes, minmax_partial_bytes()))) return rc;
  }
  CU(e, cudaEventRecord(e->ev_img_start, s));            // the tile origins, and whatever the caller queued before, come first
  CU(e, cudaStreamWaitEvent(e->s_pre, e->ev_img_start, 0));
  CU(e, cudaStreamWaitEvent(e->s_post, e->ev_img_start, 0));
  const int nchunks = (T + e->max_batch - 1) / e->max_batch, chunk = (T + nchunks - 1) / nchunks;   // balanced passes
  bool any_host_out = false;
  for (int i = 0; i < count; ++i) {
    auto& sl = e->slots[i & 1];
    // ---- pre: upload, normalise, gather ----
    if (i >= 2) CU(e, cudaStreamWaitEvent(e->s_pre, sl.ev_net, 0));      // image i-2's network pass has read this slot's crops
    const void* d_raw = imgs[i];
    if (!is_device_ptr(imgs[i])) {
      CU(e, cudaMemcpyAsync(sl.d_raw, imgs[i], npx * esz, cudaMemcpyHostToDevice, e->s_pre));
      d_raw = sl.d_raw;
    }
    const float* d_norm = reinterpret_cast<const float*>(d_raw);
    if (flags & EMD_FLAG_PREPROCESS) {
      CU(e, launch_minmax(d_raw, f64, npx, sl.d_minmax, sl.d_partial, e->s_pre));
      CU(e, launch_normalise_apply(d_raw, f64, npx, sl.d_minmax, sl.d_norm, e->s_pre));
      e->cnt.launches += 3;
      d_norm = sl.d_norm;
    }
    CU(e, launch_gather(d_norm, H, W, e->d_origins, e->d_origins + 256, ny, nx, crop, sl.d_crops, e->s_pre));
    e->cnt.launches++;
    CU(e, cudaEventRecord(sl.ev_pre, e->s_pre));
    // ---- network ----
    CU(e, cudaStreamWaitEvent(s, sl.ev_pre, 0));
    if (i >= 2) CU(e, cudaStreamWaitEvent(s, sl.ev_post, 0));            // image i-2's stitch has read this slot's tiles
    for (int c0 = 0; c0 < T; c0 += chunk) {
      const int nb = std::min(chunk, T - c0);
      if ((rc = run_network(e, sl.d_crops + c0 * per, sl.d_tiles + c0 * per, nb, mode, s))) return rc;
    }
    CU(e, cudaEventRecord(sl.ev_net, s));
    // ---- post: stitch, download ----
    CU(e, cudaStreamWaitEvent(e->s_post, sl.ev_net, 0));
    const bool out_dev = is_device_ptr(outs[i]);
    void* d_dst = out_dev ? outs[i] : sl.d_sout;
    CU(e, launch_stitch(sl.d_tiles, e->d_origins, e->d_origins + 256, ny, nx, crop, H, W, (flags & EMD_FLAG_POSTPROCESS) ? 1 : 0,
                        d_dst, out_f32 ? 1 : 0, e->s_post));
    e->cnt.launches++;
    if (!out_dev) {
      CU(e, cudaMemcpyAsync(outs[i], d_dst, npx * osz, cudaMemcpyDeviceToHost, e->s_post));
      any_host_out = true;
    }
    CU(e, cudaEventRecord(sl.ev_post, e->s_post));
  }
  for (int k = 0; k < nslots && k < count; ++k) CU(e, cudaStreamWaitEvent(s, e->slots[k].ev_post, 0));   // the caller's stream sees the results
  if (any_host_out) CU(e, cudaStreamSynchronize(e->s_post));
  return EMD_OK;
}

int emd_denoise_image(emd_engine* e, const void* img, int H, int W, int overlap, int flags, int mode, void* out,
                      void* stream) {
  if (!e || !img || !out) return EMD_EINVAL;
  int rc = check_mode(e, mode);
  if (rc) return rc;
  CU(e, cudaSetDevice(e->device));
  return denoise_images(e, &img, 1, H, W, overlap, flags, mode, &out, stream ? (cudaStream_t)stream : e->stream);
}

int emd_denoise_stream(emd_engine* e, const void* const* imgs, int count, int H, int W, int overlap, int flags, int mode,
                       void* const* outs, void* stream) {
  if (!e || !imgs || !outs || count < 0) return EMD_EINVAL;
  int rc = check_mode(e, mode);
  if (rc) return rc;
  if (count == 0) return EMD_OK;
  CU(e, cudaSetDevice(e->device));
  return denoise_images(e, imgs, count, H, W, overlap, flags, mode, outs, stream ? (cudaStream_t)stream : e->stream);
}

int emd_preprocess_crop(emd_engine* e, const float* img, int H, int W, float* out, void* stream) {
  if (!e || !img || !out || H < 1 || W < 1) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const int S = e->S;
  const size_t n_in = (size_t)H * W * 4, n_out = (size_t)S * S * 4;
  int rc;
  const float* d_src = img;
  if (!is_device_ptr(img)) {
    if ((rc = grow(e, reinterpret_cast<char**>(&e->d_img_raw), &e->img_raw_bytes, n_in))) return rc;
    CU(e, cudaMemcpyAsync(e->d_img_raw, img, n_in, cudaMemcpyHostToDevice, s));
    d_src = reinterpret_cast<const float*>(e->d_img_raw);
  }
  if ((rc = grow(e, &e->d_img, &e->img_bytes, 2 * n_out + 64))) return rc;      // [tmp][out][4 floats of min / max]
  float* d_tmp = e->d_img;
  const bool out_dev = is_device_ptr(out);
  float* d_out = out_dev ? out : e->d_img + (size_t)S * S;
  float* d_mm = e->d_img + 2 * (size_t)S * S;
  CU(e, launch_preprocess_crop(d_src, H, W, S, d_tmp, d_mm, d_out, s));
  e->cnt.launches += 5;
  if (!out_dev) {
    CU(e, cudaMemcpyAsync(out, d_out, n_out, cudaMemcpyDeviceToHost, s));
    CU(e, cudaStreamSynchronize(s));
  }
  return EMD_OK;
}

int emd_set_keep_activations(emd_engine* e, int keep) {
  if (!e) return EMD_EINVAL;
  CU(e, cudaSetDevice(e->device));
  { int drc = drain_chain(e); if (drc) return drc; }
  CU(e, cudaStreamSynchronize(e->stream));
  e->keep = keep != 0;
  return plan_arena(e);
}

// copy a (possibly channel-sliced) activation view of n images to host as dense f32 NHWC
static int download_view(emd_engine* e, const View& v, int n, int et, bool is_f32, float* host_out, cudaStream_t s) {
  const size_t px = (size_t)n * v.H * v.W, esz = is_f32 ? 4 : elem_size(et);
  char* dense = nullptr; float* f32 = nullptr;
  CU(e, cudaMalloc(&dense, px * v.C * esz));
  CU(e, cudaMemcpy2DAsync(dense, v.C * esz, reinterpret_cast<const char*>(v.ptr) + (size_t)v.coff * esz, v.pitch * esz,
                          v.C * esz, px, cudaMemcpyDeviceToDevice, s));
  const void* src = dense;
  if (!is_f32 && et != ET_F32) {
    CU(e, cudaMalloc(&f32, px * v.C * 4));
    CU(e, launch_uncast(dense, f32, px * v.C, et, s));
    src = f32;
  }
  CU(e, cudaMemcpyAsync(host_out, src, px * v.C * 4, cudaMemcpyDeviceToHost, s));
  CU(e, cudaStreamSynchronize(s));
  cudaFree(dense);
  if (f32) cudaFree(f32);
  return EMD_OK;
}

int emd_get_activation(emd_engine* e, const char* name, float* out, size_t cap_elems, int dims[4]) {
  if (!e || !name || !dims) return EMD_EINVAL;
  auto it = e->named.find(name);
  if (it == e->named.end()) return fail(e, EMD_EINVAL, "no activation named %s", name);
  if (!e->keep) return fail(e, EMD_ESTATE, "emd_set_keep_activations(e,1) before the forward pass");
  if (e->last_n <= 0) return fail(e, EMD_ESTATE, "no forward pass yet");
  const Ref r = it->second;
  const Tensor& t = e->tensors[r.t];
  if (t.external) return fail(e, EMD_EINVAL, "%s is an I/O tensor", name);
  dims[0] = e->last_n; dims[1] = t.H; dims[2] = t.W; dims[3] = r.C;
  const size_t need = (size_t)e->last_n * t.H * t.W * r.C;
  if (!out) return EMD_OK;  // size query
  if (cap_elems < need) return fail(e, EMD_EINVAL, "buffer holds %zu elements, %zu needed", cap_elems, need);
  CU(e, cudaSetDevice(e->device));
  View v; v.ptr = e->arena + t.offset; v.H = t.H; v.W = t.W; v.pitch = t.pitch; v.coff = r.coff; v.C = r.C;
  return download_view(e, v, e->last_n, e->last_et, false, out, e->stream);
}

int emd_run_layer(emd_engine* e, const char* name, const float* in, const float* in2, int n, float* out,
                  size_t out_cap_elems, int mode, int out_dims[4]) {
  if (!e || !name || !in || !out_dims) return EMD_EINVAL;
  int rc = check_mode(e, mode);
  if (rc) return rc;
  if (n < 1 || n > e->max_batch) return fail(e, EMD_EINVAL, "n=%d outside [1,%d]", n, e->max_batch);
  if ((rc = drain_chain(e))) return rc;
  std::vector<int> idx;
  for (int i = 0; i < (int)e->steps.size(); ++i)
    if (e->steps[i].layer == name) idx.push_back(i);
  if (idx.empty()) return fail(e, EMD_EINVAL, "no layer named %s", name);
  const Ref rin = e->steps[idx.front()].in, rout = e->steps[idx.back()].out;
  Ref rres;
  for (int i : idx) if (e->steps[i].res.t >= 0) rres = e->steps[i].res;
  if (rres.t >= 0 && !in2) return fail(e, EMD_EINVAL, "layer %s needs the residual operand in2", name);
  const Tensor &ti = e->tensors[rin.t], &to = e->tensors[rout.t];
  out_dims[0] = n; out_dims[1] = to.H; out_dims[2] = to.W; out_dims[3] = rout.C;
  const size_t n_out = (size_t)n * to.H * to.W * rout.C;
  if (!out) return EMD_OK;
  if (out_cap_elems < n_out) return fail(e, EMD_EINVAL, "out holds %zu elements, %zu needed", out_cap_elems, n_out);
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  const int et = mode == EMD_MODE_FP32 ? ET_F32 : (mode == EMD_MODE_BF16 ? ET_BF16 : ET_F16);
  const size_t n_in = (size_t)n * ti.H * ti.W * rin.C;
  float *d_f32 = nullptr; void *d_in = nullptr, *d_res = nullptr, *d_out = nullptr;
  const size_t n_stage = std::max(n_in, n_out);
  CU(e, cudaMalloc(&d_f32, n_stage * 4));
  CU(e, cudaMalloc(&d_in, n_in * 4));
  CU(e, cudaMalloc(&d_out, n_out * 4));
  CU(e, cudaMemsetAsync(d_out, 0, n_out * 4, s));   // written once before any TMA access (plan_arena)
  CU(e, cudaMemcpyAsync(d_f32, in, n_in * 4, cudaMemcpyHostToDevice, s));
  CU(e, launch_cast(d_f32, d_in, n_in, ti.external ? ET_F32 : et, s));
  if (rres.t >= 0) {
    CU(e, cudaMalloc(&d_res, n_out * 4));
    CU(e, cudaStreamSynchronize(s));
    CU(e, cudaMemcpyAsync(d_f32, in2, n_out * 4, cudaMemcpyHostToDevice, s));
    CU(e, launch_cast(d_f32, d_res, n_out, et, s));
  }
  ExecCtx c{e, et, n, s, reinterpret_cast<const float*>(d_in), reinterpret_cast<float*>(d_out), {}};
  c.ov.push_back(Override{rin.t, d_in});
  c.ov.push_back(Override{rout.t, d_out});
  if (rres.t >= 0) c.ov.push_back(Override{rres.t, d_res});
  for (int i : idx) {
    cudaError_t r = run_step(c, i);
    if (r != cudaSuccess) return step_error(e, i, r);
  }
  View v; v.ptr = d_out; v.H = to.H; v.W = to.W; v.pitch = rout.C; v.coff = 0; v.C = rout.C;
  rc = download_view(e, v, n, et, to.external, out, s);
  CU(e, cudaStreamSynchronize(s));
  cudaError_t last = cudaGetLastError();
  cudaFree(d_f32); cudaFree(d_in); cudaFree(d_out);
  if (d_res) cudaFree(d_res);
  if (last != cudaSuccess) return fail(e, EMD_ECUDA, "layer %s: %s", name, cudaGetErrorString(last));
  return rc;
}

long long emd_kernel_launches(const emd_engine* e) { return e ? e->cnt.launches : -1; }
long long emd_tensor_core_launches(const emd_engine* e) { return e ? e->cnt.umma : -1; }

long long emd_counter(const emd_engine* e, const char* name) {
  if (!e || !name) return -1;
  static const struct { const char* n; int k; } kinds[] = {
      {"conv_cuda_core", LK_SIMT}, {"conv_tcgen05_gen1", LK_UMMA_GEN1}, {"conv_fused_taps", LK_FUSED_TAPS}, {"conv_fused_pair", LK_FUSED_PAIR},
      {"conv_fused_dw", LK_FUSED_DW}, {"final_tcgen05", LK_FINAL_UMMA}, {"final_cuda_core", LK_FINAL_TMA}};
  const std::string s = name;
  if (s == "launches") return e->cnt.launches;
  if (s == "tensor_core_launches") return e->cnt.umma;
  if (s == "graph_replays") return e->graph_replays;
  if (s == "workspace_bytes") return (long long)e->arena_bytes;
  for (auto& k : kinds) if (s == k.n) return e->cnt.kind[k.k];
  return -1;
}

int emd_set_option(emd_engine* e, const char* name, long long value) {
  if (!name) return EMD_EINVAL;
  if (!tuning_set(name, value)) return fail(e, EMD_EINVAL, "unknown option %s", name);
  if (e) {
    if (!strcmp(name, "umma")) e->use_umma = value != 0;
    drop_graphs(e);    // captured passes hold the old kernel choices
  }
  return EMD_OK;
}

long long emd_get_option(const char* name) {
  long long v = 0;
  return (name && tuning_get(name, &v)) ? v : -1;
}
long long emd_graph_replays(const emd_engine* e) { return e ? e->graph_replays : -1; }

int emd_set_tensor_cores(emd_engine* e, int on) {
  if (!e) return EMD_EINVAL;
  e->use_umma = on != 0;
  drop_graphs(e);
  return EMD_OK;
}

int emd_set_profile(emd_engine* e, int on) {
  if (!e) return EMD_EINVAL;
  e->profile = on != 0;
  return EMD_OK;
}

int emd_num_steps(const emd_engine* e) { return e ? (int)e->steps.size() : -1; }

int emd_step_launches(const emd_engine* e, int idx) {
  if (!e || idx < 0 || idx >= (int)e->steps.size()) return -1;
  return e->steps[idx].nlaunch;
}

int emd_step_info(const emd_engine* e, int idx, char* name, size_t name_cap, float* ms, double* flops, double* bytes) {
  if (!e || idx < 0 || idx >= (int)e->steps.size()) return EMD_EINVAL;
  const Step& s = e->steps[idx];
  if (name && name_cap) { strncpy(name, s.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ms) *ms = s.ms;
  double fl = s.flops, by = s.bytes;
  if (s.fused && s.kind == SK_DW) { fl = 0; by = 0; }   // accounted in the GEMM step that computed it
  if (s.fused && s.kind == SK_CONV && idx > 0 && e->steps[idx - 1].kind == SK_DW && e->steps[idx - 1].layer == s.layer) {
    // one kernel did depthwise + pointwise: it reads the depthwise INPUT once; the intermediate never reaches HBM
    const Step& d = e->steps[idx - 1];
    fl += d.flops;
    by += d.in_bytes - s.in_bytes;
  }
  if (flops) *flops = fl;
  if (bytes) *bytes = by;
  return EMD_OK;
}

}  // extern "C"
