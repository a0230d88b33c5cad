"""Regenerates the committed fixtures in tests/golden/ from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference holds no golden vectors for this path
 and TensorFlow cannot run here, so these pin the ORACLE, not the reference:

  tile_plans.json   1-D tile origins / coverage runs for crop 512, overlap 80 (SURVEY App. D table,
                    typed in by hand below and cross-checked against oracle.wrapper.tile_origins)
  net_s64.npz       seeded 64x64 crops, the W1 oracle output for them, and per-layer statistics
  w0_digest.json    sha256 of the W0 weight set (TF-default initialisers, seed 0)
  net_s96_kat.npz   known-answer vectors for a 96x96 (small_scans shape) W1 case: for EVERY named activation of the graph
                    its float64 sum, sum of squares and 32 values at seeded positions, plus the full output (pin 3)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import wrapper  # noqa: E402
from oracle.net import OracleNet  # noqa: E402
from oracle.weights import make_w0, make_w1  # noqa: E402

# SURVEY.md App. D, "clamped origins (normative)" and coverage runs
APP_D = {
    512: {"origins": [0, 0], "runs": [[0, 512, 2]]},
    600: {"origins": [0, 88], "runs": [[0, 88, 1], [88, 512, 2], [512, 600, 1]]},
    1024: {"origins": [0, 341, 512], "runs": [[0, 341, 1], [341, 853, 2], [853, 1024, 1]]},
    2048: {"origins": [0, 410, 819, 1229, 1536],
           "runs": [[0, 410, 1], [410, 512, 2], [512, 819, 1], [819, 922, 2], [922, 1229, 1], [1229, 1331, 2],
                    [1331, 1536, 1], [1536, 1741, 2], [1741, 2048, 1]]},
    4096: {"origins": [0, 410, 819, 1229, 1638, 2048, 2458, 2867, 3277, 3584], "runs": None},
}


def runs_of(counts):
    out, start = [], 0
    for i in range(1, len(counts) + 1):
        if i == len(counts) or counts[i] != counts[start]:
            out.append([start, i, int(counts[start])])
            start = i
    return out


def main():
    plans = {}
    for size, g in APP_D.items():
        o = wrapper.tile_origins(size)
        assert o == g["origins"], (size, o)
        r = runs_of(wrapper.coverage_counts(size))
        if g["runs"] is not None:
            assert r == g["runs"], (size, r)
        plans[str(size)] = {"origins": o, "runs": r}
    # extra sizes (not in App. D): pinned from the oracle only
    for size in (513, 944, 1000, 1296, 3000):
        plans[str(size)] = {"origins": wrapper.tile_origins(size), "runs": runs_of(wrapper.coverage_counts(size))}
    for crop, ov, size in ((96, 16, 400), (64, 8, 200)):
        plans[f"{size}/{crop}/{ov}"] = {"origins": wrapper.tile_origins(size, crop, ov),
                                        "runs": runs_of(wrapper.coverage_counts(size, crop, ov))}
    json.dump(plans, open(os.path.join(HERE, "tile_plans.json"), "w"), indent=1)

    w0 = make_w0(0)
    h = hashlib.sha256()
    for k in sorted(w0):
        h.update(k.encode()); h.update(np.ascontiguousarray(w0[k]).tobytes())
    json.dump({"sha256": h.hexdigest(), "n_params": int(sum(v.size for v in w0.values()))},
              open(os.path.join(HERE, "w0_digest.json"), "w"))

    rng = np.random.default_rng(1234)
    crops = rng.random((2, 64, 64)).astype(np.float32)
    net = OracleNet(make_w1(crops, seed=0), 64)
    net.collect = True
    out = net.forward(crops)
    stats = {k: [float(v.mean()), float(v.std()), float(np.abs(v).max())] for k, v in net.acts.items()}
    np.savez_compressed(os.path.join(HERE, "net_s64.npz"), crops=crops, out=out,
                        layer_names=np.array(sorted(stats)), layer_stats=np.array([stats[k] for k in sorted(stats)]))
    print("wrote fixtures; W1 output mean %.4f std %.4f" % (out.mean(), out.std()))
    write_kat_s96()


def kat_s96_inputs():
    rng = np.random.default_rng(9696)
    return rng.random((2, 96, 96)).astype(np.float32)


def kat_positions(name, size, k=32):
    seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")
    return np.sort(np.random.default_rng(seed).choice(size, size=min(k, size), replace=False))


def write_kat_s96():
    import torch
    crops = kat_s96_inputs()
    net = OracleNet(make_w1(crops, seed=96), 96, dtype=torch.float64)
    net.collect = True
    out = net.forward(crops)
    names = sorted(net.acts)
    sums = np.array([[float(net.acts[k].sum()), float((net.acts[k].astype(np.float64) ** 2).sum())] for k in names])
    samples = np.stack([np.pad(net.acts[k].reshape(-1)[kat_positions(k, net.acts[k].size)].astype(np.float32), (0, 0)) for k in names])
    shapes = np.array([net.acts[k].shape for k in names])
    np.savez_compressed(os.path.join(HERE, "net_s96_kat.npz"), layer_names=np.array(names), sums=sums, samples=samples,
                        shapes=shapes, out=out.astype(np.float32))
    print("wrote net_s96_kat.npz:", len(names), "activations")


if __name__ == "__main__":
    main()
