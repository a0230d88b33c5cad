#!/usr/bin/env python
"""Run a few whole forward passes (for `ncu -k regex:<kernel> -s <skip> -c <n>` captures of kernels that are not a layer of
their own, e.g. the stand-alone depthwise kernel of the 728-channel blocks).  python tools/run_steps.py [--batch 32] [--reps 2]"""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--crop", type=int, default=512)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import numpy as np
    import torch
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    eng = emd.Engine(cropsize=a.crop, max_batch=a.batch)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(0)))
    x = torch.from_numpy(np.random.default_rng(0).random((a.batch, a.crop, a.crop)).astype(np.float32)).cuda()
    out = torch.empty_like(x)
    for _ in range(a.reps):
        eng.forward(x, out=out, mode="bf16")
    torch.cuda.synchronize()
    print("ok", float(out.mean()))


if __name__ == "__main__":
    main()
