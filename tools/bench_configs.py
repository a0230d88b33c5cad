#!/usr/bin/env python
"""Timings of the other BASELINE.json configs on one GPU (bench.py carries configs[1]):
  cfg1: single 512x512 crop, batch 1 (latency);  cfg3: 2048x2048 micrograph, end to end (normalise + gather + network +
  stitch; 25 crops, overlap 80), host in / host out and device in / device out;  cfg5: 96x96 crops at batch 4096.
Writes one JSON object.   python tools/bench_configs.py [--out profiles/x.json] [--fp32]"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    blob = emd.weights.pack(emd.weights.init_reference_weights(0))
    res = {}

    def timed(fn, reps=a.reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    rng = np.random.default_rng(1234)
    # cfg1 / cfg3: 512-crop engine, workspace for 25 crops
    eng = emd.Engine(cropsize=512, max_batch=25)
    eng.load_weights(blob)
    crop = rng.random((1, 512, 512)).astype(np.float32)
    dcrop = torch.from_numpy(crop).cuda()
    dout = torch.empty_like(dcrop)
    for mode in ("fp16", "bf16", "fp32"):
        res[f"cfg1_ms_per_512_crop_batch1_{mode}"] = timed(lambda: eng.forward(dcrop, out=dout, mode=mode))
    img = rng.poisson(rng.random((2048, 2048)) * 50).astype(np.float32)
    himg = torch.from_numpy(img).pin_memory()
    hout = torch.empty((2048, 2048), dtype=torch.float64).pin_memory()
    dimg = himg.cuda()
    dres = torch.empty((2048, 2048), dtype=torch.float64, device="cuda")
    res["cfg3_ms_per_2048_micrograph_fp16_host_io"] = timed(lambda: eng.denoise_image(himg, out=hout, mode="fp16"))
    res["cfg3_ms_per_2048_micrograph_fp16_device_io"] = timed(lambda: eng.denoise_image(dimg, out=dres, mode="fp16"))
    res["cfg3_ms_per_2048_micrograph_fp32_device_io"] = timed(lambda: eng.denoise_image(dimg, out=dres, mode="fp32"), reps=2, warm=1)
    res["cfg3_crops_per_micrograph"] = 25
    del eng
    # cfg5: 96x96 crops, batch 4096 (small_scans shape)
    eng = emd.Engine(cropsize=96, max_batch=4096)
    eng.load_weights(blob)
    x = torch.from_numpy(rng.random((4096, 96, 96)).astype(np.float32)).cuda()
    y = torch.empty_like(x)
    ms = timed(lambda: eng.forward(x, out=y, mode="fp16"), reps=3, warm=1)
    res["cfg5_ms_per_4096_crops_96_fp16"] = ms
    res["cfg5_crops_per_s_96_fp16"] = 4096 / ms * 1e3
    l0 = eng.kernel_launches
    eng.forward(x, out=y, mode="fp16")
    res["cfg5_kernel_launches_per_pass"] = eng.kernel_launches - l0
    print(json.dumps(res, indent=1))
    if a.out:
        json.dump(res, open(os.path.join(ROOT, a.out), "w"), indent=1)


if __name__ == "__main__":
    main()
