"""GPU parity of the quality metrics (csrc/emd_quality.cu, through the C ABI) against the float64 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(emd):
    return emd.Engine(cropsize=64, max_batch=2)


def test_quality_matches_oracle(eng):
    """MSE / Huberised loss to 1e-6 relative (FP32 differences, FP64 sums); SSIM to 2e-5 absolute (FP32 window moments like
    TensorFlow's, FP64 mean).  Odd sizes, tiles that straddle the 32x32 blocks, n > 1, host and device inputs."""
    from oracle import quality as Q
    rng = np.random.default_rng(7)
    for (n, H, W) in ((1, 11, 11), (3, 45, 70), (2, 96, 96), (1, 512, 512)):
        a = rng.random((n, H, W)).astype(np.float32)
        b = np.clip(a + rng.normal(0, 0.05, a.shape), 0, 1).astype(np.float32)
        b[0] = a[0] * 0.5 + 0.25 if n > 1 else b[0]
        got = eng.quality(a, b)
        assert got.shape == (n, 3)
        for i in range(n):
            m = Q.mse(a[i], b[i])
            assert abs(got[i, 0] - m) <= 1e-6 * m + 1e-12
            assert abs(got[i, 1] - Q.huberised(m)) <= 1e-6 * Q.huberised(m) + 1e-12
            assert abs(got[i, 2] - Q.ssim(a[i], b[i])) <= 2e-5
        dev = eng.quality(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        np.testing.assert_array_equal(dev, got)                                  # deterministic, same bits from device buffers
    same = eng.quality(a[0], a[0])
    assert same.shape == (1, 3) and same[0, 0] == 0.0 and abs(same[0, 2] - 1.0) < 1e-6
    small = eng.quality(np.full((16, 16), 0.02, np.float32), np.zeros((16, 16), np.float32))
    assert abs(small[0, 1] - 1000.0 * small[0, 0]) < 1e-12 and small[0, 0] < 1e-3  # the linear branch of the Huberised loss
    with pytest.raises(RuntimeError, match="11x11"):
        eng.quality(np.zeros((8, 20), np.float32), np.zeros((8, 20), np.float32))


def test_evaluate_runs_end_to_end(emd):
    """Low-dose generator -> Denoiser.denoise -> metrics: shapes, ranges, and that the metrics see the right pairs."""
    d = emd.Denoiser(checkpoint_loc=None, mode="bf16", cropsize=64, max_batch=4)
    yy, xx = np.mgrid[0:96, 0:80]
    clean = (0.5 + 0.4 * np.sin(yy / 9.0) * np.cos(xx / 7.0)).astype(np.float32)
    rows = emd.quality.evaluate(d, [clean, clean.T.copy()], seed=1, overlap=8)
    assert len(rows) == 2
    for r in rows:
        assert r["dose"] >= 25.0 and 0 <= r["mse_in"] < 1 and -1 <= r["ssim_in"] <= 1 and -1 <= r["ssim_out"] <= 1
        assert np.isfinite(r["psnr_in"]) and np.isfinite(r["psnr_out"]) and r["loss_out"] >= 0
