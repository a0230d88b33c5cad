#!/usr/bin/env python
"""Run one fused layer a few times on random inputs (for `ncu -k regex:... ` captures of a single kernel).
  python tools/run_layer.py --layer deconv0_0 --n 4 [--crop 512] [--mode bf16] [--reps 3]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", required=True)
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--crop", type=int, default=512)
    ap.add_argument("--mode", default="fp16")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import ctypes as C
    import numpy as np
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    eng = emd.Engine(cropsize=a.crop, max_batch=a.n)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(0)))
    # input / residual shapes from the layer table
    spec = {l[0]: l for l in emd.weights.layer_table("A")}[a.layer]
    dims = (C.c_int * 4)()
    probe = np.zeros((1,), np.float32)
    eng.lib.emd_run_layer(eng.h, a.layer.encode(), probe.ctypes.data_as(C.c_void_p), probe.ctypes.data_as(C.c_void_p), a.n, None, 0,
                          emd._lib.MODES[a.mode], dims)
    oh, ow, oc = dims[1], dims[2], dims[3]
    cin = spec[2]
    stride = 2 if a.layer.endswith("_strided") or (a.layer.startswith("residual") and not a.layer.endswith("_d")) else 1
    ih, iw = (oh * stride, ow * stride) if spec[1] != "deconv" else (oh // 2, ow // 2)
    rng = np.random.default_rng(0)
    x = rng.random((a.n, ih, iw, cin), dtype=np.float32)
    res = rng.random((a.n, oh, ow, oc), dtype=np.float32)
    for _ in range(a.reps):
        try:
            out = eng.run_layer(a.layer, x, None, mode=a.mode)
        except RuntimeError:
            out = eng.run_layer(a.layer, x, res, mode=a.mode)
    print(a.layer, x.shape, "->", out.shape, float(out.mean()))


if __name__ == "__main__":
    main()
