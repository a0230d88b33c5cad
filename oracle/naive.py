"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

A SECOND, independent restatement of the reference graph, in plain numpy float64, written from the reference source and the
published TensorFlow op definitions without looking at how oracle/net.py does it: every op is spelled out from its
definition (explicit SAME padding and tap loops instead of torch convolutions, the sub-pixel closed form of the transposed
convolution instead of conv_transpose + crop, a per-pixel formula for the legacy bilinear resize instead of interpolation
matrices, un-fused BatchNorm).  tests/test_oracle_cpu.py checks that the two restatements agree to rounding on calibrated
weights: no TensorFlow output exists in this container to pin either of them, so what this buys is protection against a
transcription error in one of them, nothing more.

Graph: ``architecture()`` in misc_py/denoiser-multi-gpu.py:200-540 (variant "A") and machine_learning/denoiser.py:58-398
(variant "B").  Arrays are NHWC like TensorFlow's.  Parameter names are those of oracle/net.py:param_shapes.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (misc_py/apply_autoencoders.py:106-115)


def _same(n, k, stride, rate):
    """TensorFlow SAME: out = ceil(n / stride); total padding so that the last window fits; the smaller half goes first."""
    out = -(-n // stride)
    total = max((out - 1) * stride + (k - 1) * rate + 1 - n, 0)
    return out, total // 2, total - total // 2


def _windows(x, k, stride, rate):
    """Yield (ky, kx, view) where view[n, oy, ox, c] = padded x at (oy*stride + ky*rate, ox*stride + kx*rate)."""
    n, h, w, c = x.shape
    oh, pt, pb = _same(h, k, stride, rate)
    ow, pl, pr = _same(w, k, stride, rate)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, c), x.dtype)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    for ky in range(k):
        for kx in range(k):
            y0, x0 = ky * rate, kx * rate
            yield ky, kx, xp[:, y0:y0 + (oh - 1) * stride + 1:stride, x0:x0 + (ow - 1) * stride + 1:stride, :]


def depthwise(x, w, stride=1, rate=1):
    """DepthwiseConv2dNative, depth multiplier 1: out[.., c] = sum_taps in[.., c] * w[ky, kx, c, 0] (DMG:253-273)."""
    out = None
    for ky, kx, v in _windows(x, 3, stride, rate):
        t = v * w[ky, kx, :, 0]
        out = t if out is None else out + t
    return out


def conv(x, kernel, bias, stride=1, rate=1):
    """tf.layers.conv2d / slim.conv2d, SAME, kernel [kh, kw, Cin, Cout], bias added (DMG:231-235, 299-327, 365-370)."""
    out = None
    for ky, kx, v in _windows(x, kernel.shape[0], stride, rate):
        t = np.tensordot(v, kernel[ky, kx], axes=([3], [0]))
        out = t if out is None else out + t
    return out if bias is None else out + bias


def conv_transpose(x, kernel, bias):
    """conv2d_transpose 3x3, stride 2, SAME; kernel [3, 3, Cout, Cin] (DMG:281-286).  It is the adjoint of the SAME stride-2
    conv (window of output j covers inputs 2j..2j+2), so per axis out[2j] = in[j] k[0] + in[j-1] k[2], out[2j+1] = in[j] k[1]."""
    n, h, w, _ = x.shape
    cout = kernel.shape[2]
    out = np.zeros((n, 2 * h, 2 * w, cout), x.dtype)
    taps = {0: ((0, 0), (2, 1)), 1: ((1, 0),)}      # output parity -> ((kernel index, how far back the input sits), ...)
    for a in (0, 1):
        for b in (0, 1):
            acc = np.zeros((n, h, w, cout), x.dtype)
            for ky, dy in taps[a]:
                for kx, dx in taps[b]:
                    src = np.zeros_like(x)
                    src[:, dy:, dx:, :] = x[:, :h - dy, :w - dx, :]
                    acc += np.tensordot(src, kernel[ky, kx], axes=([3], [1]))
            out[:, a::2, b::2, :] = acc
    return out + bias


def resize(x, oh, ow):
    """tf.image.resize_images as TensorFlow 1 did it (bilinear, align_corners=False, no half-pixel centres; DMG:344, 494):
    src = dst * in / out, the four neighbours at floor(src) and floor(src) + 1 (clamped), top row interpolated, then rows."""
    n, h, w, c = x.shape
    if (h, w) == (oh, ow):
        return x
    out = np.empty((n, oh, ow, c), x.dtype)
    for oy in range(oh):
        sy = oy * (h / oh)
        y0 = int(math.floor(sy))
        y1 = min(y0 + 1, h - 1)
        fy = sy - y0
        for ox in range(ow):
            sx = ox * (w / ow)
            x0 = int(math.floor(sx))
            x1 = min(x0 + 1, w - 1)
            fx = sx - x0
            top = x[:, y0, x0] + (x[:, y0, x1] - x[:, y0, x0]) * fx
            bot = x[:, y1, x0] + (x[:, y1, x1] - x[:, y1, x0]) * fx
            out[:, oy, ox] = top + (bot - top) * fy
    return out


def avg_pool(x):
    """tf.nn.pool AVG 2x2 stride 2 SAME on an even-sized map: the plain mean of each 2x2 cell (DMG:331-335)."""
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).mean(axis=(2, 4))


def relu6(x):
    return np.minimum(np.maximum(x, 0.0), 6.0)


class NaiveNet:
    def __init__(self, params, cropsize, variant="A"):
        self.p = {k: np.asarray(v, np.float64) for k, v in params.items()}
        self.S = cropsize
        self.variant = variant

    def bn(self, x, prefix):
        """Inference BatchNorm exactly as written: gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta (DMG:210-218)."""
        p = self.p
        return p[prefix + "/gamma"] * (x - p[prefix + "/mean"]) / np.sqrt(p[prefix + "/var"] + EPS) + p[prefix + "/beta"]

    def sep(self, x, name, stride=1, rate=1):
        """strided_conv_block (DMG:250-276): separable conv whose normalizer_fn is a BatchNorm (so it has no bias), then
        batch_then_activ: a second BatchNorm and ReLU6."""
        y = depthwise(x, self.p[name + "/dw"], stride, rate)
        y = conv(y, self.p[name + "/pw"], None)
        return relu6(self.bn(self.bn(y, name + "/bn1"), name + "/bn2"))

    def dense(self, x, name, stride=1, rate=1):
        """conv + bias -> BatchNorm -> ReLU6 (conv_block_not_sep DMG:225-238, residual_conv DMG:363-373, ASPP convs)."""
        return relu6(self.bn(conv(x, self.p[name + "/kernel"], self.p[name + "/bias"], stride, rate), name + "/bn"))

    def up(self, x, name):
        """deconv_block (DMG:278-289)."""
        return relu6(self.bn(conv_transpose(x, self.p[name + "/tkernel"], self.p[name + "/bias"]), name + "/bn"))

    def aspp(self, x):
        size = self.S // 16          # aspp_size (32 at the reference's 512 x 512 crops, DMG:108-112)
        branches = [self.dense(x, "aspp_1x1")]
        if self.variant == "A":      # DMG:291-361
            branches += [self.dense(x, "aspp_r%d" % r, rate=r) for r in (6, 12, 18)]
            pooled = conv(avg_pool(x), self.p["aspp_image/kernel"], self.p["aspp_image/bias"])
            branches.append(relu6(self.bn(resize(pooled, size, size), "aspp_image/bn")))
        else:                        # DEN:152-218: separable dilated branches with one more BN + ReLU6, identity image branch
            for r in (6, 12, 18):
                branches.append(relu6(self.bn(self.sep(x, "aspp_r%d" % r, rate=r), "aspp_r%d_post/bn" % r)))
            branches.append(relu6(self.bn(resize(x, size, size), "aspp_image/bn")))
        return self.dense(np.concatenate(branches, axis=3), "aspp_pellet")

    def forward(self, crops):
        x = np.asarray(crops, np.float64).reshape(-1, self.S, self.S, 1)
        t, skips = x, []
        for i in range(4):           # encoding blocks 0-3 (DMG:395-453): the strided output plus the strided 1x1 of the block input
            a = self.sep(self.sep(self.sep(t, "cnn%d" % i), "cnn%d_last" % i), "cnn%d_strided" % i, stride=2)
            t = a + self.dense(t, "residual%d" % i, stride=2)
            skips.append(t)
        a = t
        for j in range(3):           # encoding block 4 (DMG:455-466)
            a = self.sep(a, "cnn4_%d" % j)
        t = a + t
        for b in range(11):          # xception_middle_block x num_extra_blocks (DMG:375-390, 468-469)
            a = t
            for j in range(3):
                a = self.sep(a, "mid%d_%d" % (b, j))
            t = a + t
        y = self.aspp(t)
        y = resize(y, self.S // 4, self.S // 4)                     # DMG:494
        cat = np.concatenate([y, skips[1]], axis=3)                 # [upsampled ASPP, cnn1_strided] (DMG:497-499)
        y = self.sep(self.sep(cat, "deconv2_0"), "deconv2_1") + self.dense(cat, "residual2_d")
        y = self.up(y, "deconv2to1")
        cat = np.concatenate([y, skips[0]], axis=3)                 # [deconv2to1, cnn0_strided] (DMG:509-511)
        y = self.sep(self.sep(cat, "deconv1_0"), "deconv1_1") + self.dense(cat, "residual1_d")
        u = self.up(y, "deconv1to0")
        y = self.sep(self.sep(u, "deconv0_0"), "deconv0_1") + self.dense(u, "residual0_d")
        y = self.dense(y, "final")                                  # conv_block_not_sep(deconv0, 1): kernel_size defaults to 3
        if self.variant == "A":
            y = np.clip(y, 0.0, 1.0)                                # DMG:534-538; variant B returns the raw prediction
        return y.reshape(-1, self.S, self.S)
