"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

numpy restatement of the reference's crop-tile / normalise / stitch wrapper:
``Denoiser.denoise`` (machine_learning/denoiser.py:653-682, identical copy in
misc_py/denoiser_class_function-tmp.py:3-32), ``scale0to1``
(machine_learning/denoiser.py:684-695 == misc_py/denoiser-multi-gpu.py:817-828)
and the training-time ``preprocess`` order (misc_py/denoiser-multi-gpu.py:853-858).

The reference method cannot execute as written; the five repairs of
SURVEY.md App. D (D-1..D-5) are applied and marked where they happen.
"""
from __future__ import annotations

import numpy as np


def scale0to1(img):
    """DEN:684-695.  Constant image -> 0.5 (filled in place, like the reference)."""
    mn = np.min(img)
    mx = np.max(img)
    if mn == mx:
        img.fill(0.5)
    else:
        img = (img - mn) / (mx - mn)
    return img.astype(np.float32)


def normalise(img):
    """Whole-image normalisation used by ``denoise`` (repair D-5): NaN -> 0.5,
    Inf -> 0.5, *then* one scale0to1 -- the order the net was trained with
    (DMG:853-858).  Works on a copy; arithmetic in the input dtype (App. A.10)."""
    img = np.array(img, copy=True)
    if not np.issubdtype(img.dtype, np.floating):
        img = img.astype(np.float32)
    img[np.isnan(img)] = 0.5
    img[np.isinf(img)] = 0.5
    return scale0to1(img)


def preprocess_crop(img, size=512):
    """``Denoiser.preprocess`` exactly as the class file has it (DEN:632-643), used by ``denoise_crop(preprocess=True)``:
    cv2.resize to size x size (INTER_LINEAR: half-pixel centres, edge replicate), scale0to1, NaN -> 0.5, Inf -> 0.5,
    scale0to1 again, reshape to (1, size, size, 1).  The first min-max runs BEFORE the NaN/Inf replacement (App. D-5): one NaN
    turns the whole crop into 0.5.  The resize is restated in numpy so the check does not depend on OpenCV."""
    img = np.asarray(img)
    h, w = img.shape
    fy = (np.arange(size, dtype=np.float64) + 0.5) * (h / size) - 0.5     # cv2: src = (dst + 0.5) * scale - 0.5
    fx = (np.arange(size, dtype=np.float64) + 0.5) * (w / size) - 0.5
    y0 = np.floor(fy).astype(np.int64); x0 = np.floor(fx).astype(np.int64)
    wy = (fy - y0).astype(np.float32); wx = (fx - x0).astype(np.float32)
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    x0c, x1c = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    src = img.astype(np.float32)
    rows = src[:, x0c] * (1.0 - wx)[None, :] + src[:, x1c] * wx[None, :]  # horizontal pass, then vertical (cv2's order)
    out = rows[y0c] * (1.0 - wy)[:, None] + rows[y1c] * wy[:, None]
    if (h, w) == (size, size):
        out = src.copy()            # cv2.resize copies when the size does not change (no 0 * Inf)
    out = scale0to1(out.astype(img.dtype if np.issubdtype(img.dtype, np.floating) else np.float32))
    out[np.isnan(out)] = 0.5
    out[np.isinf(out)] = 0.5
    return scale0to1(out).reshape(1, size, size, 1)


def tile_origins(size: int, crop: int = 512, overlap: int = 80):
    """1-D tile origins: DEN:661-667 with repairs D-2 (int, round-half-even) and D-3 (clamp)."""
    if size < crop:
        raise ValueError(f"image side {size} smaller than crop {crop}")
    num = size // (crop - overlap) + 1          # DEN:661
    length = size / num                          # DEN:663 (true division)
    return [min(int(np.round(i * length)), size - crop) for i in range(num)]


def coverage_counts(size: int, crop: int = 512, overlap: int = 80):
    """contributions[] along one axis (DEN:659, 675)."""
    c = np.zeros(size, np.int32)
    for o in tile_origins(size, crop, overlap):
        c[o:o + crop] += 1
    return c


def gather_crops(img, crop=512, overlap=80):
    """The slicing of DEN:671-673 for every (i,j): returns ([ny*nx,crop,crop], ys, xs)."""
    ys = tile_origins(img.shape[0], crop, overlap)
    xs = tile_origins(img.shape[1], crop, overlap)
    crops = np.stack([img[y:y + crop, x:x + crop] for y in ys for x in xs]).astype(np.float32)
    return crops, ys, xs


def stitch(tiles, ys, xs, H, W, crop=512, clip=True):
    """DEN:658-659, 671-680 with repair D-4 (accumulate with +=).  float64 accumulators."""
    den = np.zeros((H, W))                       # DEN:658
    cnt = np.zeros((H, W))                       # DEN:659
    t = 0
    for y in ys:
        for x in xs:
            den[y:y + crop, x:x + crop] += tiles[t].reshape(crop, crop)
            cnt[y:y + crop, x:x + crop] += 1     # DEN:675
            t += 1
    den /= cnt                                   # DEN:677
    return den.clip(0.0, 1.0) if clip else den   # DEN:679-682


def denoise(img, crop_fn, preprocess=True, postprocess=True, overlap=80, crop=512):
    """Repaired ``Denoiser.denoise``; crop_fn maps [n,crop,crop] f32 -> [n,crop,crop] (the network)."""
    img = normalise(img) if preprocess else np.asarray(img, np.float32)
    crops, ys, xs = gather_crops(img, crop, overlap)
    out = crop_fn(crops)
    return stitch(out, ys, xs, img.shape[0], img.shape[1], crop, clip=postprocess)
