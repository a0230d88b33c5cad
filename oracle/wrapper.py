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
