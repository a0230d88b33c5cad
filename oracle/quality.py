"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): float64 restatement of the reference's quality metrics,
misc_py/denoiser-multi-gpu.py:124-167 (tf_ssim, _tf_fspecial_gauss) and :772-773 (MSE, Huberised loss).  PARITY UNPINNED like
the rest of the oracle: TensorFlow cannot run here and the reference holds no fixture for these functions; the window and
the formula are closed-form and are checked against hand-computable cases in tests/test_quality_cpu.py."""
import numpy as np


def fspecial_gauss(size=11, sigma=1.5):
    """_tf_fspecial_gauss, DMG:124-139 (window built in float32 like the TF constants, then normalised)."""
    x, y = np.mgrid[-size // 2 + 1:size // 2 + 1, -size // 2 + 1:size // 2 + 1]
    g = np.exp(-((x.astype(np.float32) ** 2 + y.astype(np.float32) ** 2) / np.float32(2.0 * sigma ** 2))).astype(np.float32)
    return g / g.sum(dtype=np.float32)


def _valid_filter(img, win):
    from numpy.lib.stride_tricks import sliding_window_view
    return np.einsum("yxij,ij->yx", sliding_window_view(img, win.shape), win)


def ssim(a, b, size=11, sigma=1.5):
    """tf_ssim(cs_map=False, mean_metric=True), DMG:142-167, in float64."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    w = fspecial_gauss(size, sigma).astype(np.float64)
    c1, c2 = (0.01 * 1) ** 2, (0.03 * 1) ** 2
    mu1, mu2 = _valid_filter(a, w), _valid_filter(b, w)
    s11 = _valid_filter(a * a, w) - mu1 * mu1
    s22 = _valid_filter(b * b, w) - mu2 * mu2
    s12 = _valid_filter(a * b, w) - mu1 * mu2
    v = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return float(v.mean())


def mse(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return float((d * d).mean())


def huberised(m):
    """DMG:773."""
    return 1000.0 * m if m < 0.001 else float(np.sqrt(1000.0 * m))
