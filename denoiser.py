"""Module-level drop-in for the reference's ``machine_learning/denoiser.py``:
``from denoiser import Denoiser, scale0to1`` keeps working.  The package directory name contains
hyphens (it is fixed by the project name), so it is loaded through importlib."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
Denoiser = emd.Denoiser
scale0to1 = emd.scale0to1
Engine = emd.Engine
weights = emd.weights


def disp(img):
    """DEN:697-703 showed the image in an OpenCV window; there is no display here, so this
    returns the rescaled image it would have shown."""
    return scale0to1(img)


if __name__ == "__main__":  # DEN:705-708
    import numpy as np
    denoiser = Denoiser(visible_cuda=None)
    print(disp(denoiser.denoise(np.random.rand(512, 512))).shape)
