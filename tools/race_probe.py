#!/usr/bin/env python
"""Find the first activation that differs between repeated identical device-buffer passes (intermittent-race hunt).
  python tools/race_probe.py [iters] [keep(0/1)]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
keep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(99)
crops = rng.random((16, 512, 512)).astype(np.float32)
eng = emd.Engine(cropsize=512, max_batch=16)
eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(1)))
names = []
for i in range(4):
    names += [f"cnn{i}", f"cnn{i}_last", f"residual{i}", f"enc{i}"]
names += ["cnn4_0", "cnn4_1", "trunk4"]
for b in range(11):
    names += [f"mid{b}_0", f"mid{b}_1", f"trunk_mid{b}"]
names += ["aspp_1x1", "aspp_r6", "aspp_r12", "aspp_r18", "aspp_image", "aspp_pellet", "upsample4", "deconv2_0", "residual2_d", "dec2",
          "deconv2to1", "deconv1_0", "residual1_d", "dec1", "deconv1to0", "deconv0_0", "residual0_d", "dec0"]
x = torch.from_numpy(crops).cuda()
torch.cuda.synchronize()
if keep:
    eng.set_keep_activations(True)
out0 = eng.forward(x, mode="bf16"); torch.cuda.synchronize()
out0 = out0.cpu().numpy()
A = {n: eng.activation(n).copy() for n in names} if keep else {}
for it in range(iters):
    o = eng.forward(x, mode="bf16"); torch.cuda.synchronize()
    o = o.cpu().numpy()
    per = [int((o[i] != out0[i]).sum()) for i in range(16)]
    print("iter", it, "output mismatches per crop", per if any(per) else "none")
    if keep and any(per):
        for n in names:
            b = eng.activation(n)
            m = (b != A[n])
            if m.any():
                imgs = sorted(set(np.nonzero(m.reshape(16, -1).any(axis=1))[0].tolist()))
                print("   first differing activation:", n, "elements", int(m.sum()), "crops", imgs, "max abs", float(np.abs(b.astype(np.float64) - A[n]).max()))
                break
