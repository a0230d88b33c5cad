"""CPU tests of the quality-metric oracle (oracle/quality.py) on closed-form cases, and of the low-dose generator."""
import numpy as np

from oracle import quality as Q


def test_window_matches_the_reference_formula():
    w = Q.fspecial_gauss(11, 1.5)
    assert w.shape == (11, 11) and abs(float(w.sum()) - 1.0) < 1e-6
    assert np.allclose(w, w.T) and np.allclose(w, w[::-1, ::-1]) and w[5, 5] == w.max()
    # separable: exp(-(x^2+y^2)/2s^2) = g(x) g(y)
    g = np.exp(-(np.arange(-5, 6) ** 2) / (2 * 1.5 ** 2))
    assert np.allclose(w, np.outer(g, g) / np.outer(g, g).sum(), atol=1e-7)


def test_ssim_closed_form_cases():
    rng = np.random.default_rng(0)
    a = rng.random((40, 37))
    assert abs(Q.ssim(a, a) - 1.0) < 1e-12                                   # identical images
    ca, cb = np.full((20, 20), 0.3), np.full((20, 20), 0.6)                  # constants: variances vanish
    expect = (2 * 0.3 * 0.6 + 1e-4) / (0.3 ** 2 + 0.6 ** 2 + 1e-4)
    assert abs(Q.ssim(ca, cb) - expect) < 1e-4                               # (the float32 window sums to 1 +- 1e-7)
    b = rng.random((40, 37))
    assert abs(Q.ssim(a, b) - Q.ssim(b, a)) < 1e-12 and Q.ssim(a, b) < 0.2   # symmetric; unrelated noise is dissimilar
    assert Q.ssim(a, 1.0 - a) < 0.0                                          # anti-correlated


def test_mse_and_huberised_loss():
    a, b = np.zeros((8, 8)), np.full((8, 8), 0.02)
    m = Q.mse(a, b)
    assert abs(m - 4e-4) < 1e-15 and abs(Q.huberised(m) - 0.4) < 1e-12       # below 0.001: 1000 * mse
    m2 = Q.mse(a, np.full((8, 8), 0.1))
    assert abs(Q.huberised(m2) - np.sqrt(10.0)) < 1e-12                      # above: sqrt(1000 * mse)


def test_low_dose_generator(emd):
    q = emd.quality
    rng = np.random.default_rng(3)
    clean = np.clip(np.add.outer(np.linspace(0.1, 0.9, 64), np.linspace(0, 0.1, 64)), 0, 1).astype(np.float32)
    scales = [q.get_scale(rng) for _ in range(2000)]
    assert min(scales) >= 25.0 and abs(np.mean(scales) - 100.0) < 6.0        # 25 + Exp(75)
    lq = q.gen_lq(clean, 50.0, rng)
    assert lq.dtype == np.float32 and lq.min() == 0.0 and lq.max() == 1.0    # scale0to1 of the counts
    assert np.corrcoef(lq.mean(1), clean.mean(1))[0, 1] > 0.9
    hi = q.gen_lq(clean, 5000.0, np.random.default_rng(1))
    lo = q.gen_lq(clean, 5.0, np.random.default_rng(1))
    ref = (clean - clean.min()) / (clean.max() - clean.min())
    assert Q.mse(hi, ref) < Q.mse(lo, ref)                                   # more dose, less noise
