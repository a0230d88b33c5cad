#!/usr/bin/env python
"""Reproducer of the intermittent first-keep-pass mismatch: [new engine, arena-planned pass, keep mode on (workspace re-planned
and re-allocated), first keep pass] repeated; counts passes whose output differs from the planned pass and, for the first bad
one, names the first activation that differs from the (correct) second keep pass.
  python tools/keep_repro.py [iters] [batch] [max_batch]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

def names():
    out = []
    for i in range(4):
        out += [f"cnn{i}", f"cnn{i}_last", f"residual{i}", f"enc{i}"]
    out += ["cnn4_0", "cnn4_1", "trunk4"]
    for b in range(11):
        out += [f"mid{b}_0", f"mid{b}_1", f"trunk_mid{b}"]
    out += ["aspp_1x1", "aspp_r6", "aspp_r12", "aspp_r18", "aspp_image", "aspp_pellet", "upsample4", "deconv2_0", "residual2_d", "dec2",
            "deconv2to1", "deconv1_0", "residual1_d", "dec1", "deconv1to0", "deconv0_0", "residual0_d", "dec0"]
    return out

def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else n
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    rng = np.random.default_rng(99)
    crops = rng.random((n, 512, 512)).astype(np.float32)
    if os.environ.get("PROBE_PINNED"):
        pin = torch.from_numpy(crops).pin_memory()
        crops = pin.numpy()
    dirty = not os.environ.get("PROBE_CLEAN")
    blob = emd.weights.pack(emd.weights.init_reference_weights(1))
    bad = crashes = 0
    analysed = False
    for it in range(iters):
        eng = emd.Engine(cropsize=512, max_batch=mb)
        eng.load_weights(blob)
        a = eng.forward(crops, mode="fp16")
        d = eng.forward(torch.from_numpy(crops[::-1].copy()).cuda(), mode="fp16"); torch.cuda.synchronize()
        w0 = eng.counter("workspace_bytes")
        eng.set_keep_activations(True)
        if it == 0: print("workspace: planned %.2f GB, keep mode %.2f GB" % (w0 / 1e9, eng.counter("workspace_bytes") / 1e9), flush=True)
        try:
            k = eng.forward(crops, mode="fp16")
            per = [int((a[i] != k[i]).sum()) for i in range(n)]
            A1 = {}
            if any(per) and not analysed and n <= 8:
                for nm in names():
                    try: A1[nm] = eng.activation(nm).copy()
                    except Exception: pass
            k2 = eng.forward(crops, mode="fp16")
        except RuntimeError as ex:
            crashes += 1
            print("iter", it, "CRASH", str(ex)[:200], flush=True)
            break
        per2 = [int((a[i] != k2[i]).sum()) for i in range(n)]
        if any(per) or any(per2):
            bad += 1
            print("iter", it, "first keep pass mismatches per crop", per, "second", per2 if any(per2) else "none", flush=True)
        if A1:
            analysed = True
            for nm in names():
                if nm not in A1: continue
                b = eng.activation(nm)
                m = b != A1[nm]
                if m.any():
                    imgs = sorted(set(np.nonzero(m.reshape(n, -1).any(axis=1))[0].tolist()))
                    yy, xx = np.nonzero(m[imgs[0]].any(axis=-1))
                    chs = sorted(set(np.nonzero(m[imgs[0]].any(axis=(0, 1)))[0].tolist()))
                    w = A1[nm][imgs[0]][m[imgs[0]]]
                    print("   differs:", nm, b.shape, "elements", int(m.sum()), "crops", imgs, "rows", int(yy.min()), int(yy.max()), "cols", int(xx.min()), int(xx.max()),
                          "channels", chs[:4], "..", chs[-1], "n", len(chs), "| wrong values: NaN", int(np.isnan(w).sum()), "zeros", int((w == 0).sum()), "sample", w[:4].tolist(),
                          "right", b[imgs[0]][m[imgs[0]]][:4].tolist(), flush=True)
        eng.set_keep_activations(False)
        eng.close()
        if dirty:
            junk = [torch.full((1 << 28,), float("nan"), device="cuda") for _ in range(8)]   # dirty the freed memory
            del junk
        torch.cuda.empty_cache()
    print("env", {k_: v for k_, v in os.environ.items() if k_.startswith("EMD_")}, "iters", it + 1, "bad", bad, "crashes", crashes, flush=True)

if __name__ == "__main__":
    main()
