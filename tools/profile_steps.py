#!/usr/bin/env python
"""Per-step device times of one forward pass (CUDA events around every step of the schedule), with
the algorithmic GB/s and TFLOP/s of each step against the measured peaks.  Usage:
  python tools/profile_steps.py [--batch 32] [--crop 512] [--mode bf16] [--out profiles/x.txt]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--crop", type=int, default=512)
    ap.add_argument("--mode", default="fp16")
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-tc", action="store_true")
    a = ap.parse_args()
    import numpy as np
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    eng = emd.Engine(cropsize=a.crop, max_batch=a.batch)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(0)))
    if a.no_tc:
        eng.set_tensor_cores(False)
    x = np.random.default_rng(0).random((a.batch, a.crop, a.crop)).astype(np.float32)
    import torch
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    for _ in range(3):
        eng.forward(xd, out=out, mode=a.mode)
    eng.set_profile(True)
    acc = None
    reps = 3
    for _ in range(reps):
        eng.forward(xd, out=out, mode=a.mode)
        info = eng.step_info()
        acc = info if acc is None else [(i[0], i[1] + j[1], i[2], i[3], i[4]) for i, j in zip(acc, info)]
    info = [(n, ms / reps, fl, by, nl) for n, ms, fl, by, nl in acc]
    tot = sum(i[1] for i in info)
    B = a.batch
    lines = [f"# per-step device time, batch {B} x {a.crop}^2, mode {a.mode}, CUDA events, mean of {reps}; "
             f"peaks: HBM {pk['hbm_gbs']} GB/s, tensor {pk['bf16_tflops_sustained']} TFLOP/s",
             "# bytes / FLOPs are algorithmic per step as it ran (a separable block whose depthwise runs inside the GEMM kernel is one step)",
             f"{'step':20s} {'ms':>8s} {'share':>6s} {'GB/s':>8s} {'%hbm':>6s} {'TFLOP/s':>8s} {'%tc':>6s} {'roof_ms':>8s} {'launches':>8s}"]
    roof_tot = 0.0
    for n, ms, fl, by, nl in info:
        if nl == 0:
            continue   # computed inside the next step's kernel
        gbs = by * B / (ms * 1e-3) / 1e9 if ms > 0 else 0
        tf = fl * B / (ms * 1e-3) / 1e12 if ms > 0 else 0
        roof = max(by * B / (pk["hbm_gbs"] * 1e9), fl * B / (pk["bf16_tflops_sustained"] * 1e12)) * 1e3
        roof_tot += roof
        lines.append(f"{n:20s} {ms:8.3f} {100 * ms / tot:5.1f}% {gbs:8.0f} {100 * gbs / pk['hbm_gbs']:5.1f}% {tf:8.1f} "
                     f"{100 * tf / pk['bf16_tflops_sustained']:5.1f}% {roof:8.3f} {nl:8d}")
    lines.append(f"{'TOTAL':20s} {tot:8.3f} ms -> {B / tot * 1e3:.0f} crops/s; roofline {roof_tot:.3f} ms "
                 f"({B / roof_tot * 1e3:.0f} crops/s); fraction {roof_tot / tot:.3f}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.join(ROOT, a.out)), exist_ok=True)
        open(os.path.join(ROOT, a.out), "w").write(text + "\n")


if __name__ == "__main__":
    main()
