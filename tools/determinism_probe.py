#!/usr/bin/env python
"""Determinism / batch-invariance probe at BASELINE.json's crop size: repeats host-buffer (chunked, copy streams) and
device-buffer passes in several orders and reports per-crop mismatch counts.  python tools/determinism_probe.py [mode] [iters]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = np.random.default_rng(99)
crops = rng.random((16, 512, 512)).astype(np.float32)
eng = emd.Engine(cropsize=512, max_batch=16)
eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(1)))
ref = eng.forward(crops, mode=mode)
rev = np.ascontiguousarray(crops[::-1])
bad = 0
for it in range(iters):
    a = eng.forward(crops, mode=mode)
    c = eng.forward(crops[5:6], mode=mode)
    t = torch.from_numpy(rev).cuda()
    d = eng.forward(t, mode=mode)
    torch.cuda.synchronize()
    d = d.cpu().numpy()[::-1]
    t2 = torch.from_numpy(crops).cuda()
    f = eng.forward(t2, mode=mode)
    torch.cuda.synchronize()
    f = f.cpu().numpy()
    rows = {"host16": [int((a[i] != ref[i]).sum()) for i in range(16)], "single5": [int((c[0] != ref[5]).sum())],
            "dev16rev": [int((d[i] != ref[i]).sum()) for i in range(16)], "dev16": [int((f[i] != ref[i]).sum()) for i in range(16)]}
    for k, v in rows.items():
        if any(v):
            bad += 1
            print(f"iter {it} {k}: mismatching elements per crop {v}")
print("iterations", iters, "mismatching passes", bad)
