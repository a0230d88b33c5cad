"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

CPU restatement (PyTorch, FP32 with an FP64 switch) of the reference's
atrous-convolutional Xception encoder-decoder denoiser graph:

  * variant "A": ``architecture()`` in misc_py/denoiser-multi-gpu.py:200-540
    (canonical: dense dilated ASPP, in-graph clip).
  * variant "B": ``architecture()`` in machine_learning/denoiser.py:58-398
    (deployed: separable ASPP + extra BN/ReLU6, identity image branch, no clip).

TensorFlow 1.x is not importable here, so every op follows the published TF
semantics listed in SURVEY.md App. A; each helper cites the reference line
whose behaviour it restates.  All tensors handed in/out are NHWC numpy arrays
(TF layout); NCHW is used only internally because torch convolutions want it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# Hyper-parameters: misc_py/denoiser-multi-gpu.py:51-63, 108-112
FEATURES = (64, 128, 256, 728, 728)
ASPP_FILTERS = 728
ASPP_OUTPUT = 256
ASPP_RATES = (6, 12, 18)
NUM_EXTRA_BLOCKS = 11
BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default (misc_py/apply_autoencoders.py:106-115)


# ----------------------------------------------------------------------------
# TF-semantic primitive ops (NCHW torch tensors)
# ----------------------------------------------------------------------------

def same_pad(in_size: int, k: int, stride: int, rate: int = 1):
    """TF 'SAME' padding amounts (before, after) along one axis (App. A.1)."""
    out = -(-in_size // stride)
    k_eff = (k - 1) * rate + 1
    total = max((out - 1) * stride + k_eff - in_size, 0)
    return total // 2, total - total // 2


def _pad_same(x, k, stride, rate):
    ht, hb = same_pad(x.shape[2], k, stride, rate)
    wl, wr = same_pad(x.shape[3], k, stride, rate)
    return F.pad(x, (wl, wr, ht, hb))


def depthwise3x3(x, w_tf, stride=1, rate=1):
    """DepthwiseConv2dNative of slim.separable_convolution2d (DMG:253-273).

    w_tf: [3,3,C,1] TF layout.  Stride and rate both apply here (App. A.2).
    """
    c = x.shape[1]
    w = w_tf.permute(2, 3, 0, 1).contiguous()  # [C,1,3,3]
    return F.conv2d(_pad_same(x, 3, stride, rate), w, None, stride=stride, dilation=rate, groups=c)


def conv2d(x, kernel_tf, bias, stride=1, rate=1):
    """tf.layers.conv2d, padding SAME (DMG:231-235, 299-327, 365-370); kernel [kh,kw,Cin,Cout]."""
    k = kernel_tf.shape[0]
    w = kernel_tf.permute(3, 2, 0, 1).contiguous()  # [Cout,Cin,kh,kw]
    return F.conv2d(_pad_same(x, k, stride, rate), w, bias, stride=stride, dilation=rate)


def conv2d_transpose_s2(x, kernel_tf, bias):
    """tf.layers.conv2d_transpose 3x3 stride 2 SAME (DMG:281-286); kernel [3,3,Cout,Cin].

    Adjoint of the SAME stride-2 conv (pad 0 before / 1 after), no kernel flip:
    out[2j] = in[j] w[0] + in[j-1] w[2], out[2j+1] = in[j] w[1]  (App. A.4).
    """
    h, w_ = x.shape[2], x.shape[3]
    w = kernel_tf.permute(3, 2, 0, 1).contiguous()  # conv_transpose weight [Cin,Cout,kh,kw]
    y = F.conv_transpose2d(x, w, None, stride=2, padding=0)[:, :, : 2 * h, : 2 * w_]
    return y + bias.view(1, -1, 1, 1)


def _resize_matrix(n_in: int, n_out: int, dtype):
    """1-D interpolation matrix of TF1 legacy ResizeBilinear, align_corners=False (App. A.5)."""
    m = torch.zeros(n_out, n_in, dtype=dtype)
    scale = n_in / n_out
    for o in range(n_out):
        src = o * scale
        lo = int(np.floor(src))
        hi = min(lo + 1, n_in - 1)
        w = src - lo
        m[o, lo] += 1.0 - w
        m[o, hi] += w
    return m


def resize_bilinear_legacy(x, oh, ow):
    """tf.image.resize_images(x,[oh,ow]) as TF1 did it (DMG:344, 494): no half-pixel centres."""
    if x.shape[2] == oh and x.shape[3] == ow:
        return x
    my = _resize_matrix(x.shape[2], oh, x.dtype)
    mx = _resize_matrix(x.shape[3], ow, x.dtype)
    # out = top + (bottom-top)*wy with top = tl + (tr-tl)*wx: separable, x first then y
    t = torch.einsum("nchw,xw->nchx", x, mx)
    return torch.einsum("nchx,yh->ncyx", t, my)


def avg_pool_2x2(x):
    """tf.nn.pool AVG 2x2 stride 2 SAME on even sizes (DMG:331-335, App. A.6)."""
    assert x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
    return F.avg_pool2d(x, 2, 2)


def relu6(x):
    """tf.nn.relu6 (DMG:222)."""
    return torch.clamp(x, 0.0, 6.0)


# ----------------------------------------------------------------------------
# Layer inventory (creation order of DMG:392-531) -- used by the weight generators
# ----------------------------------------------------------------------------

def layer_specs(variant: str = "A"):
    """[(name, kind, cin, cout, k)] in the textual creation order of the graph.

    kind: 'sep' (strided_conv_block, DMG:250-276), 'conv' (tf.layers.conv2d + BN,
    DMG:225-238 / 291-373), 'deconv' (DMG:278-289), 'bn' (stand-alone BN, variant B).
    """
    f0, f1, f2, f3, f4 = FEATURES
    s = []
    enc = [(1, f0, f0, f1), (f1, f1, f1, f1), (f1, f2, f2, f2), (f2, f3, f3, f3)]
    for i, (cin, a, b, c) in enumerate(enc):
        s += [(f"cnn{i}", "sep", cin, a, 3), (f"cnn{i}_last", "sep", a, b, 3),
              (f"cnn{i}_strided", "sep", b, c, 3), (f"residual{i}", "conv", cin, c, 1)]
    s += [(f"cnn4_{j}", "sep", f4, f4, 3) for j in range(3)]
    for b in range(NUM_EXTRA_BLOCKS):
        s += [(f"mid{b}_{j}", "sep", f4, f4, 3) for j in range(3)]
    if variant == "A":
        s += [("aspp_1x1", "conv", f4, ASPP_FILTERS, 1)]
        s += [(f"aspp_r{r}", "conv", f4, ASPP_FILTERS, 3) for r in ASPP_RATES]
        s += [("aspp_image", "conv", f4, ASPP_FILTERS, 1)]
    else:
        s += [("aspp_1x1", "conv", f4, ASPP_FILTERS, 1)]
        for r in ASPP_RATES:
            s += [(f"aspp_r{r}", "sep", f4, ASPP_FILTERS, 3), (f"aspp_r{r}_post", "bn", ASPP_FILTERS, ASPP_FILTERS, 0)]
        s += [("aspp_image", "bn", f4, f4, 0)]
    s += [("aspp_pellet", "conv", 5 * ASPP_FILTERS, ASPP_OUTPUT, 1)]
    s += [("deconv2_0", "sep", ASPP_OUTPUT + f1, f2, 3), ("deconv2_1", "sep", f2, f2, 3),
          ("residual2_d", "conv", ASPP_OUTPUT + f1, f2, 1), ("deconv2to1", "deconv", f2, f2, 3)]
    s += [("deconv1_0", "sep", f2 + f1, f1, 3), ("deconv1_1", "sep", f1, f1, 3),
          ("residual1_d", "conv", f2 + f1, f1, 1), ("deconv1to0", "deconv", f1, f1, 3)]
    s += [("deconv0_0", "sep", f1, f0, 3), ("deconv0_1", "sep", f0, f0, 3),
          ("residual0_d", "conv", f1, f0, 1), ("final", "conv", f0, 1, 3)]
    return s


def param_shapes(variant: str = "A"):
    """name -> shape, TF variable layouts."""
    shapes = {}
    for name, kind, cin, cout, k in layer_specs(variant):
        if kind == "sep":
            shapes[f"{name}/dw"] = (3, 3, cin, 1)
            shapes[f"{name}/pw"] = (1, 1, cin, cout)
            bns = ("bn1", "bn2")
        elif kind == "conv":
            shapes[f"{name}/kernel"] = (k, k, cin, cout)
            shapes[f"{name}/bias"] = (cout,)
            bns = ("bn",)
        elif kind == "deconv":
            shapes[f"{name}/tkernel"] = (3, 3, cout, cin)
            shapes[f"{name}/bias"] = (cout,)
            bns = ("bn",)
        else:
            bns = ("bn",)
        for b in bns:
            for v in ("beta", "gamma", "mean", "var"):
                shapes[f"{name}/{b}/{v}"] = (cout,)
    return shapes


# ----------------------------------------------------------------------------
# The graph
# ----------------------------------------------------------------------------

class OracleNet:
    """Forward pass of the reference graph on CPU.

    params: dict name -> numpy array in TF layouts (see param_shapes).
    collect: keep every named activation (NHWC numpy) in ``self.acts``.
    calibrate: overwrite each BN's moving mean/var with the statistics of the
               tensor it sees (what training would converge to) -- used to build
               the W1 weight set, SURVEY App. E.3.
    """

    def __init__(self, params, cropsize=512, variant="A", dtype=torch.float32):
        assert cropsize % 16 == 0
        self.S = cropsize
        self.variant = variant
        self.dtype = dtype
        self.p = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in params.items()}
        self.collect = False
        self.calibrate = False
        self.acts = {}
        # BF16-storage emulation (error-budget experiments for the tensor-core path): round GEMM
        # operands (activations as stored, depthwise output, weights) to bf16, accumulate in FP32.
        self.emulate_bf16 = False
        self.emulate_dtype = None      # torch.bfloat16 / torch.float16: same emulation with that storage type
        self.trunk_fp32 = False
        self.q_only = None             # None = every source; else a set of tags (see _q)

    _PRE_SUM = ("_strided", "cnn4_2", "deconv2_1", "deconv1_1", "deconv0_1")

    def _q(self, x, tag="act"):
        """Round to the emulated 16-bit storage type.  tag: what is being stored ('dw' = depthwise result handed to
        the GEMM, 'w' = GEMM weights, 'act' = an activation tensor); ``q_only`` restricts the emulation to some tags."""
        dt = self.emulate_dtype or (torch.bfloat16 if self.emulate_bf16 else None)
        if dt is None or (self.q_only is not None and tag not in self.q_only):
            return x
        return x.to(dt).to(self.dtype)

    def _qt(self, x):
        """The 728-wide trunk: bf16 storage unless ``trunk_fp32`` (then only GEMM/dw inputs are rounded)."""
        return x if self.trunk_fp32 else self._q(x)

    def _qout(self, name, y):
        """Layer outputs are stored in bf16 unless the layer's epilogue adds a residual first."""
        pre_sum = name.endswith(self._PRE_SUM) or (name.startswith("mid") and name.endswith("_2"))
        return y if pre_sum or name == "final" else self._q(y)

    # -- blocks ---------------------------------------------------------------
    def _bn(self, x, prefix):
        """_batch_norm_fn, inference mode (DMG:210-218, App. A.3)."""
        if self.calibrate:
            self.p[f"{prefix}/mean"] = x.mean(dim=(0, 2, 3))
            self.p[f"{prefix}/var"] = x.var(dim=(0, 2, 3), unbiased=False)
        g, b = self.p[f"{prefix}/gamma"], self.p[f"{prefix}/beta"]
        m, v = self.p[f"{prefix}/mean"], self.p[f"{prefix}/var"]
        a = g / torch.sqrt(v + BN_EPS)
        return x * a.view(1, -1, 1, 1) + (b - m * a).view(1, -1, 1, 1)

    def _rescale(self, y, wname):
        """Calibration only: scale the layer's weights so its pre-BN output has unit variance."""
        if self.calibrate:
            s = 1.0 / float(y.std().clamp_min(1e-20))
            self.p[wname] = self.p[wname] * s
            y = y * s
        return y

    def _keep(self, name, x):
        if self.collect:
            self.acts[name] = x.permute(0, 2, 3, 1).contiguous().numpy()
        return x

    def sep(self, x, name, stride=1, rate=1):
        """strided_conv_block: dw3x3 -> pw1x1 (no bias) -> BN -> BN -> ReLU6 (DMG:250-276)."""
        y = depthwise3x3(x, self.p[f"{name}/dw"], stride, rate)
        y = self._q(y, "dw")
        if self.collect:
            self.acts[f"{name}:dw"] = y.permute(0, 2, 3, 1).contiguous().numpy()
        y = self._rescale(conv2d(y, self._q(self.p[f"{name}/pw"], "w"), None), f"{name}/pw")
        y = self._bn(y, f"{name}/bn1")
        y = self._bn(y, f"{name}/bn2")
        return self._keep(name, self._qout(name, relu6(y)))

    def conv(self, x, name, stride=1, rate=1):
        """conv_block_not_sep / residual_conv / ASPP convs: conv + bias -> BN -> ReLU6 (DMG:225-238, 363-373)."""
        y = self._rescale(conv2d(x, self._q(self.p[f"{name}/kernel"], "w"), None, stride, rate), f"{name}/kernel")
        y = self._bn(y + self.p[f"{name}/bias"].view(1, -1, 1, 1), f"{name}/bn")
        return self._keep(name, self._qout(name, relu6(y)))

    def deconv(self, x, name):
        """deconv_block (DMG:278-289)."""
        zero = torch.zeros_like(self.p[f"{name}/bias"])
        y = self._rescale(conv2d_transpose_s2(x, self._q(self.p[f"{name}/tkernel"], "w"), zero), f"{name}/tkernel")
        y = self._bn(y + self.p[f"{name}/bias"].view(1, -1, 1, 1), f"{name}/bn")
        return self._keep(name, self._qout(name, relu6(y)))

    def aspp(self, x):
        """aspp_block: DMG:291-361 (variant A) / DEN:152-218 (variant B)."""
        s16 = self.S // 16
        b0 = self.conv(x, "aspp_1x1")
        if self.variant == "A":
            br = [self.conv(x, f"aspp_r{r}", rate=r) for r in ASPP_RATES]
            pool = self._q(avg_pool_2x2(x))
            pool = self._rescale(conv2d(pool, self.p["aspp_image/kernel"], None), "aspp_image/kernel")
            pool = pool + self.p["aspp_image/bias"].view(1, -1, 1, 1)
            pool = resize_bilinear_legacy(pool, s16, s16)
            pool = self._keep("aspp_image", self._q(relu6(self._bn(pool, "aspp_image/bn"))))
        else:
            br = []
            for r in ASPP_RATES:
                y = self.sep(x, f"aspp_r{r}", rate=r)
                br.append(self._keep(f"aspp_r{r}_post", relu6(self._bn(y, f"aspp_r{r}_post/bn"))))
            pool = resize_bilinear_legacy(x, s16, s16)  # identity (DEN:199)
            pool = self._keep("aspp_image", relu6(self._bn(pool, "aspp_image/bn")))
        cat = torch.cat([b0] + br + [pool], dim=1)  # order: DMG:348-350
        return self.conv(cat, "aspp_pellet")

    # -- whole graph ------------------------------------------------------------
    def forward(self, crops):
        """crops: [N,S,S] (or [N,S,S,1]) float -> [N,S,S] float32/64 numpy."""
        x = torch.as_tensor(np.asarray(crops)).to(self.dtype).reshape(-1, 1, self.S, self.S)
        with torch.no_grad():
            skips = []
            t = x
            for i in range(4):  # encoding blocks 0-3, DMG:395-453
                a = self.sep(t, f"cnn{i}")
                a = self.sep(a, f"cnn{i}_last")
                a = self.sep(a, f"cnn{i}_strided", stride=2)
                r = self.conv(t, f"residual{i}", stride=2)
                t = self._keep(f"enc{i}", self._q(a + r))
                skips.append(t)
            a = t
            for j in range(3):  # encoding block 4, DMG:455-466
                a = self.sep(a, f"cnn4_{j}")
            t = self._keep("trunk4", self._qt(a + t))
            for b in range(NUM_EXTRA_BLOCKS):  # DMG:468-469 -> 375-390
                a = t
                for j in range(3):
                    a = self.sep(a, f"mid{b}_{j}")
                t = self._keep(f"trunk_mid{b}", self._qt(a + t))
            aspp = self.aspp(t)  # DMG:472
            s4 = self.S // 4
            up = self._q(resize_bilinear_legacy(aspp, s4, s4))  # DMG:494
            cat2 = self._keep("concat2", torch.cat([up, skips[1]], dim=1))  # DMG:497-499
            d = self.sep(cat2, "deconv2_0")
            d = self.sep(d, "deconv2_1")
            d = self._keep("dec2", self._q(d + self.conv(cat2, "residual2_d")))
            d = self.deconv(d, "deconv2to1")
            cat1 = self._keep("concat1", torch.cat([d, skips[0]], dim=1))  # DMG:509-511
            d = self.sep(cat1, "deconv1_0")
            d = self.sep(d, "deconv1_1")
            d = self._keep("dec1", self._q(d + self.conv(cat1, "residual1_d")))
            d1to0 = self.deconv(d, "deconv1to0")
            d = self.sep(d1to0, "deconv0_0")
            d = self.sep(d, "deconv0_1")
            d = self._keep("dec0", self._q(d + self.conv(d1to0, "residual0_d")))
            y = self.conv(d, "final")  # 3x3, DMG:531 (kernel_size defaults to 3, DMG:225)
            if self.variant == "A":
                y = torch.clamp(y, 0.0, 1.0)  # DMG:534-538
            y = self._keep("output", y)
        return y.reshape(-1, self.S, self.S).numpy()

    def run_layer(self, name, x, res=None):
        """One layer of the schedule on an NHWC numpy input (``res``: NHWC tensor added after the activation, the
        post-ReLU6 adds of DMG:408/423/438/453/466/390/504/516/528) -> NHWC numpy.  The per-layer parity tests feed
        the same seeded tensors to this and to the engine's ``emd_run_layer``."""
        t = torch.as_tensor(np.asarray(x)).to(self.dtype).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            if name.startswith("residual") and not name.endswith("_d"):
                y = self.conv(t, name, stride=2)                       # residual_conv, DMG:363-373
            elif name.endswith("_strided"):
                y = self.sep(t, name, stride=2)
            elif name in ("deconv2to1", "deconv1to0"):
                y = self.deconv(t, name)
            elif name.startswith("aspp_r") and self.variant == "A":
                y = self.conv(t, name, rate=int(name[6:]))
            elif name in ("aspp_1x1", "aspp_pellet") or name.endswith("_d"):
                y = self.conv(t, name)
            elif name == "final":
                y = self.conv(t, name)
                if self.variant == "A":
                    y = torch.clamp(y, 0.0, 1.0)
            elif name == "upsample4":
                y = resize_bilinear_legacy(t, self.S // 4, self.S // 4)
            else:
                y = self.sep(t, name)                                   # every other layer is a strided_conv_block, stride 1
            if res is not None:
                y = y + torch.as_tensor(np.asarray(res)).to(self.dtype).permute(0, 3, 1, 2)
        return y.permute(0, 2, 3, 1).contiguous().numpy()

    def export_params(self):
        return {k: v.to(torch.float32).numpy().copy() for k, v in self.p.items()}
