"""Quality evaluation of the denoiser: the reference's low-dose generator and its metrics.

  get_scale / gen_lq     misc_py/denoiser-multi-gpu.py:785-799  mean dose 25 + Exp(75); Poisson(img * scale), then scale0to1
  mse, huberised         misc_py/denoiser-multi-gpu.py:772-773  tf.losses.mean_squared_error; < 0.001 ? 1000 mse : sqrt(1000 mse)
  ssim                   misc_py/denoiser-multi-gpu.py:124-167  tf_ssim: 11x11 Gaussian window (sigma 1.5), VALID, L = 1

The metrics run on the GPU behind ``emd_quality`` (csrc/emd_quality.cu); the Poisson draw stays on the host like the
reference's (numpy) -- with an explicit generator instead of the reference's reseeding from ``np.random.rand``.
"""
from __future__ import annotations

import numpy as np

from .denoiser import scale0to1


def get_scale(rng: np.random.Generator) -> float:
    """DMG:785-786."""
    return 25.0 + rng.exponential(75.0)


def gen_lq(img, scale, rng: np.random.Generator, img_type=np.float32):
    """DMG:789-799: low-dose version of a clean image in [0,1]: Poisson counts at mean dose ``scale``, rescaled to [0,1]."""
    lq = rng.poisson(np.asarray(img, np.float64) * scale)
    return scale0to1(lq.astype(np.float64)).astype(img_type)


def psnr(mse, peak=1.0):
    return 10.0 * np.log10(peak * peak / np.maximum(mse, 1e-30))


def evaluate(denoiser, clean_images, seed=0, overlap=80):
    """Denoise low-dose versions of ``clean_images`` (2-D arrays in [0,1], at least one crop in size) and report, per image,
    the dose, MSE / PSNR / SSIM of the noisy input and of the denoised output against the clean image."""
    rng = np.random.default_rng(seed)
    rows = []
    for img in clean_images:
        clean = np.ascontiguousarray(img, np.float32)
        scale = get_scale(rng)
        noisy = gen_lq(clean, scale, rng)
        restored = denoiser.denoise(noisy, overlap=overlap).astype(np.float32)
        q_in = denoiser.engine.quality(noisy, clean)[0]
        q_out = denoiser.engine.quality(restored, clean)[0]
        rows.append({"dose": scale, "mse_in": q_in[0], "psnr_in": psnr(q_in[0]), "ssim_in": q_in[2],
                     "mse_out": q_out[0], "loss_out": q_out[1], "psnr_out": psnr(q_out[0]), "ssim_out": q_out[2]})
    return rows
