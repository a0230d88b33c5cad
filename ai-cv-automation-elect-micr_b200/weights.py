"""Host-side weight exporter: reference variables -> packed blob for ``emd_load_weights``.

Replaces the variable side of the reference: ``tf.train.Saver().restore``
(machine_learning/denoiser.py:626-627) and TF's initialisers for a fresh graph.

Parameter dict convention (TF variable layouts), per layer L:
  separable block (DMG:250-276):  L/dw [3,3,Cin,1], L/pw [1,1,Cin,Cout], L/bn1/*, L/bn2/*
  dense conv (DMG:225-238 ...):   L/kernel [kh,kw,Cin,Cout], L/bias [Cout], L/bn/*
  transposed conv (DMG:278-289):  L/tkernel [3,3,Cout,Cin], L/bias [Cout], L/bn/*
with bn/* = beta, gamma, mean, var.

Blob layout (little endian): 32-byte header {"EMDW0001", u32 n_entries, u32 variant, 4xu32 0},
n_entries x 64-byte records {char name[48], u32 rows, u32 cols, u64 offset}, then FP32 row-major
arrays at 64-byte aligned offsets:
  L/dw    [9][Cin]          tap-major depthwise weights
  L/w     [k*k*Cin][Cout]   GEMM B operand, K index = (ky*k+kx)*Cin + ci (TF kernel flattened)
  L/scale [1][Cout], L/shift [1][Cout]   folded BatchNorm(s) (+ conv bias), App. A.2/A.3
"""
from __future__ import annotations

import struct

import numpy as np

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default (misc_py/apply_autoencoders.py:106-115)

FEATURES = (64, 128, 256, 728, 728)   # DMG:51-55
ASPP_OUTPUT = 256                      # DMG:57
ASPP_RATES = (6, 12, 18)               # DMG:60-62
NUM_EXTRA_BLOCKS = 11                  # DMG:63


def layer_table(variant: str = "A"):
    """[(name, kind, cin, cout, k)] in the graph's creation order (variant A: DMG:392-531; variant B: DEN:250-389).
    kind 'bn' = a stand-alone BatchNorm + ReLU6 (variant B only, DEN:166-200)."""
    if variant not in ("A", "B"):
        raise ValueError(f"unknown graph variant {variant!r}")
    f0, f1, f2, f3, f4 = FEATURES
    t = []
    for i, (cin, a, b, c) in enumerate([(1, f0, f0, f1), (f1, f1, f1, f1), (f1, f2, f2, f2), (f2, f3, f3, f3)]):
        t += [(f"cnn{i}", "sep", cin, a, 3), (f"cnn{i}_last", "sep", a, b, 3),
              (f"cnn{i}_strided", "sep", b, c, 3), (f"residual{i}", "conv", cin, c, 1)]
    t += [(f"cnn4_{j}", "sep", f4, f4, 3) for j in range(3)]
    for blk in range(NUM_EXTRA_BLOCKS):
        t += [(f"mid{blk}_{j}", "sep", f4, f4, 3) for j in range(3)]
    t += [("aspp_1x1", "conv", f4, f4, 1)]
    if variant == "A":   # dense dilated branches + pooled image-level conv (DMG:306-345)
        t += [(f"aspp_r{r}", "conv", f4, f4, 3) for r in ASPP_RATES]
        t += [("aspp_image", "conv", f4, f4, 1)]
    else:                # separable dilated branches each followed by another BN + ReLU6; identity image branch (DEN:166-200)
        for r in ASPP_RATES:
            t += [(f"aspp_r{r}", "sep", f4, f4, 3), (f"aspp_r{r}_post", "bn", f4, f4, 0)]
        t += [("aspp_image", "bn", f4, f4, 0)]
    t += [("aspp_pellet", "conv", 5 * f4, ASPP_OUTPUT, 1)]
    t += [("deconv2_0", "sep", ASPP_OUTPUT + f1, f2, 3), ("deconv2_1", "sep", f2, f2, 3),
          ("residual2_d", "conv", ASPP_OUTPUT + f1, f2, 1), ("deconv2to1", "deconv", f2, f2, 3),
          ("deconv1_0", "sep", f2 + f1, f1, 3), ("deconv1_1", "sep", f1, f1, 3),
          ("residual1_d", "conv", f2 + f1, f1, 1), ("deconv1to0", "deconv", f1, f1, 3),
          ("deconv0_0", "sep", f1, f0, 3), ("deconv0_1", "sep", f0, f0, 3),
          ("residual0_d", "conv", f1, f0, 1), ("final", "conv", f0, 1, 3)]
    return t


def init_reference_weights(seed: int = 0, variant: str = "A"):
    """A fresh graph's variables: Glorot-uniform kernels with TF's fan rule
    (xavier_initializer DMG:265; tf.layers default), zero biases (DMG:267), BatchNorm
    beta 0 / gamma 1 / moving mean 0 / moving variance 1."""
    rng = np.random.default_rng(seed)
    p = {}

    def glorot(shape):
        rf = shape[0] * shape[1]
        lim = np.sqrt(6.0 / (shape[2] * rf + shape[3] * rf))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    def bn(prefix, c):
        p[f"{prefix}/beta"] = np.zeros(c, np.float32)
        p[f"{prefix}/gamma"] = np.ones(c, np.float32)
        p[f"{prefix}/mean"] = np.zeros(c, np.float32)
        p[f"{prefix}/var"] = np.ones(c, np.float32)

    for name, kind, cin, cout, k in layer_table(variant):
        if kind == "sep":
            p[f"{name}/dw"] = glorot((3, 3, cin, 1))
            p[f"{name}/pw"] = glorot((1, 1, cin, cout))
            bn(f"{name}/bn1", cout)
            bn(f"{name}/bn2", cout)
        elif kind == "conv":
            p[f"{name}/kernel"] = glorot((k, k, cin, cout))
            p[f"{name}/bias"] = np.zeros(cout, np.float32)
            bn(f"{name}/bn", cout)
        elif kind == "bn":
            bn(f"{name}/bn", cout)
        else:
            p[f"{name}/tkernel"] = glorot((3, 3, cout, cin))
            p[f"{name}/bias"] = np.zeros(cout, np.float32)
            bn(f"{name}/bn", cout)
    return p


def _bn_affine(p, prefix):
    """Inference BatchNorm as y = a*x + b (float64)."""
    g, b = p[f"{prefix}/gamma"].astype(np.float64), p[f"{prefix}/beta"].astype(np.float64)
    m, v = p[f"{prefix}/mean"].astype(np.float64), p[f"{prefix}/var"].astype(np.float64)
    a = g / np.sqrt(v + BN_EPS)
    return a, b - a * m


def fold(params, variant: str = "A"):
    """name -> 2-D float32 array, as the blob stores them."""
    out = {}
    for name, kind, cin, cout, k in layer_table(variant):
        if kind == "sep":
            out[f"{name}/dw"] = params[f"{name}/dw"].reshape(9, cin)
            out[f"{name}/w"] = params[f"{name}/pw"].reshape(cin, cout)
            a1, b1 = _bn_affine(params, f"{name}/bn1")
            a2, b2 = _bn_affine(params, f"{name}/bn2")
            scale, shift = a1 * a2, a2 * b1 + b2          # bn2(bn1(x))
        elif kind == "bn":
            scale, shift = _bn_affine(params, f"{name}/bn")
        else:
            if kind == "conv":
                w = params[f"{name}/kernel"].reshape(k * k * cin, cout)
            else:  # [3,3,Cout,Cin] -> [(ky,kx,ci), co]
                w = np.transpose(params[f"{name}/tkernel"], (0, 1, 3, 2)).reshape(9 * cin, cout)
            out[f"{name}/w"] = w
            a, b = _bn_affine(params, f"{name}/bn")
            bias = params[f"{name}/bias"].astype(np.float64)
            if name == "aspp_image" and variant == "A":
                # conv + bias, THEN resize, THEN BN/ReLU6 (DMG:338-345): keep them apart
                out["aspp_image/one"] = np.ones((1, cout))
                out["aspp_image/bias"] = bias.reshape(1, cout)
                out["aspp_image/bnshift"] = b.reshape(1, cout)
                scale, shift = a, a * bias + b             # (kept for completeness; unused by the engine)
            else:
                scale, shift = a, a * bias + b
        out[f"{name}/scale"] = scale.reshape(1, cout)
        out[f"{name}/shift"] = shift.reshape(1, cout)
    return {k_: np.ascontiguousarray(v, dtype=np.float32).reshape(v.shape[0], -1) for k_, v in out.items()}


def pack(params, variant: str = "A") -> bytes:
    """Serialise folded weights into the blob ``emd_load_weights`` takes."""
    arrays = fold(params, variant)
    names = sorted(arrays)
    head = 32 + 64 * len(names)
    off = (head + 63) & ~63
    records, chunks = [], []
    for n in names:
        a = arrays[n]
        nb = a.size * 4
        records.append(struct.pack("<48sIIQ", n.encode(), a.shape[0], a.shape[1], off))
        chunks.append((off, a.tobytes()))
        off = (off + nb + 63) & ~63
    blob = bytearray(off)
    blob[:32] = struct.pack("<8sII4I", b"EMDW0001", len(names), 0 if variant == "A" else 1, 0, 0, 0, 0)
    blob[32:head] = b"".join(records)
    for o, c in chunks:
        blob[o:o + len(c)] = c
    return bytes(blob)
