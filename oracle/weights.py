"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

Weight sets for parity checks:

  W0  TF-default initialisers: Glorot-uniform kernels with TF's fan rule
      (tf.contrib.layers.xavier_initializer, DMG:265; tf.layers default),
      zero biases (DMG:267), fresh BatchNorm (beta 0, gamma 1, mean 0, var 1).
  W1  W0, then non-trivial gamma/beta/bias and every layer rescaled and every
      BatchNorm's moving statistics calibrated on a batch, so that each
      layer's output is O(1).  W0 alone is blind to most of the network
      (App. E.3): a 728-wide separable branch carries ~5e-5 of the trunk.
"""
from __future__ import annotations

import numpy as np

from .net import OracleNet, param_shapes


def make_w0(seed: int = 0, variant: str = "A"):
    """TF-default initialisation (App. E.2)."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape in param_shapes(variant).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("dw", "pw", "kernel", "tkernel"):
            rf = shape[0] * shape[1]
            fan_in, fan_out = shape[2] * rf, shape[3] * rf
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf in ("gamma", "var"):
            p[name] = np.ones(shape, np.float32)
        else:  # bias, beta, mean
            p[name] = np.zeros(shape, np.float32)
    return p


def make_w1(calib_crops, seed: int = 0, variant: str = "A", cropsize: int | None = None):
    """BN-calibrated weight set.  calib_crops: [N,S,S] in [0,1]."""
    calib_crops = np.asarray(calib_crops, np.float32)
    cropsize = cropsize or calib_crops.shape[-1]
    p = make_w0(seed, variant)
    rng = np.random.default_rng(seed + 1000003)
    for name in sorted(p):
        leaf = name.rsplit("/", 1)[1]
        if leaf == "gamma":
            p[name] = rng.uniform(0.5, 1.5, p[name].shape).astype(np.float32)
        elif leaf == "beta":
            p[name] = rng.uniform(-0.5, 0.5, p[name].shape).astype(np.float32)
        elif leaf == "bias":
            p[name] = rng.uniform(-0.1, 0.1, p[name].shape).astype(np.float32)
    # the last layer feeds clip(0,1): centre it so the output is not pinned at a rail
    p["final/bn/beta"][:] = 0.5
    p["final/bn/gamma"][:] = 0.25
    net = OracleNet(p, cropsize, variant)
    net.calibrate = True
    net.forward(calib_crops)
    return net.export_params()
