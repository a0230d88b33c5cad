"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol that
include/emd.h declares, the integer tile planner matches the goldens, the weight exporter folds
BatchNorm correctly -- and nothing silently falls back to the CPU."""
import ctypes as C
import importlib
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def declared_functions():
    text = open(os.path.join(ROOT, "include", "emd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emd_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(emd):
    lib = emd._lib.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libemd.so does not export {n}"
    assert sorted(emd._lib.SIGNATURES) == names  # the ctypes table mirrors the header one to one
    assert lib.emd_version() >= 100


def test_plan_tiles_matches_goldens(emd):
    lib = emd._lib.load()
    plans = json.load(open(os.path.join(GOLD, "tile_plans.json")))
    for key, g in plans.items():
        parts = [int(v) for v in key.split("/")]
        size, crop, ov = (parts + [512, 80])[:3] if len(parts) == 1 else parts
        ys, xs = (C.c_int * 64)(), (C.c_int * 64)()
        ny, nx = C.c_int(), C.c_int()
        assert lib.emd_plan_tiles(size, size + 0, crop, ov, ys, xs, C.byref(ny), C.byref(nx)) == 0
        assert list(ys[: ny.value]) == g["origins"], key  # bit-exact (integer)
        assert list(xs[: nx.value]) == g["origins"], key


def test_plan_tiles_rectangular_and_errors(emd):
    from oracle import wrapper as W
    lib = emd._lib.load()
    ys, xs = (C.c_int * 64)(), (C.c_int * 64)()
    ny, nx = C.c_int(), C.c_int()
    for H, Wd in [(600, 2048), (1537, 999), (512, 4096)]:
        assert lib.emd_plan_tiles(H, Wd, 512, 80, ys, xs, C.byref(ny), C.byref(nx)) == 0
        assert list(ys[: ny.value]) == W.tile_origins(H) and list(xs[: nx.value]) == W.tile_origins(Wd)
    assert lib.emd_plan_tiles(511, 600, 512, 80, ys, xs, C.byref(ny), C.byref(nx)) != 0  # smaller than a crop
    assert lib.emd_plan_tiles(600, 600, 512, 512, ys, xs, C.byref(ny), C.byref(nx)) != 0  # overlap >= crop


def test_plan_tiles_random_sizes_match_oracle(emd):
    from oracle import wrapper as W
    lib = emd._lib.load()
    rng = np.random.default_rng(7)
    ys, xs = (C.c_int * 256)(), (C.c_int * 256)()
    ny, nx = C.c_int(), C.c_int()
    for _ in range(300):
        crop = int(rng.choice([64, 96, 512]))
        ov = int(rng.integers(0, crop // 2))
        size = int(rng.integers(crop, crop * 9))
        assert lib.emd_plan_tiles(size, size, crop, ov, ys, xs, C.byref(ny), C.byref(nx)) == 0
        assert list(ys[: ny.value]) == W.tile_origins(size, crop, ov), (size, crop, ov)


def test_no_cpu_fallback_without_gpu(emd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no usable CUDA device|CUDA"):
        emd.Engine(cropsize=64, max_batch=1)


def test_weight_table_matches_oracle_inventory(emd):
    from oracle.net import layer_specs, param_shapes
    assert emd.weights.layer_table("A") == layer_specs("A")
    p = emd.weights.init_reference_weights(0)
    shapes = param_shapes("A")
    assert set(p) == set(shapes)
    for k, v in p.items():
        assert v.shape == shapes[k], k
    assert sum(v.size for v in p.values()) == 38772462
    # Glorot-uniform with TF's fan rule (App. E.2): depthwise [3,3,C,1] -> limit sqrt(6/(9C+9))
    assert np.abs(p["cnn1/dw"]).max() <= np.sqrt(6.0 / (9 * 128 + 9)) + 1e-7
    assert np.abs(p["mid3_1/pw"]).max() <= np.sqrt(6.0 / (728 + 728)) + 1e-7
    assert not p["final/bias"].any() and (p["final/bn/var"] == 1).all()


def test_bn_fold_equals_unfused_batchnorm(emd):
    rng = np.random.default_rng(0)
    p = emd.weights.init_reference_weights(1)
    for k in p:
        if k.endswith(("beta", "mean", "bias")):
            p[k] = rng.normal(size=p[k].shape).astype(np.float32)
        elif k.endswith(("gamma", "var")):
            p[k] = rng.uniform(0.5, 2.0, size=p[k].shape).astype(np.float32)
    f = emd.weights.fold(p)
    x = rng.normal(size=(5, 128)).astype(np.float64)

    def bn(v, pre):
        return p[pre + "/gamma"] * (v - p[pre + "/mean"]) / np.sqrt(p[pre + "/var"] + 1e-3) + p[pre + "/beta"]
    ref = bn(bn(x, "cnn1/bn1"), "cnn1/bn2")                       # separable block: two BNs in a row (DMG:263, 274)
    np.testing.assert_allclose(x * f["cnn1/scale"] + f["cnn1/shift"], ref, rtol=2e-6, atol=2e-6)
    ref = bn(x + p["residual1/bias"], "residual1/bn")             # dense conv: bias then BN
    np.testing.assert_allclose(x * f["residual1/scale"] + f["residual1/shift"], ref, rtol=2e-6, atol=2e-6)
    # transposed-conv kernel [3,3,Cout,Cin] becomes [(ky,kx,ci), co]
    tk = p["deconv1to0/tkernel"]
    assert f["deconv1to0/w"][(1 * 3 + 2) * 128 + 5, 7] == tk[1, 2, 7, 5]
    assert f["aspp_r6/w"].shape == (9 * 728, 728) and f["cnn0/dw"].shape == (9, 1)


def test_blob_roundtrip(emd):
    import struct
    p = emd.weights.init_reference_weights(0)
    blob = emd.weights.pack(p)
    magic, n, variant = struct.unpack_from("<8sII", blob, 0)
    assert magic == b"EMDW0001" and variant == 0
    f = emd.weights.fold(p)
    assert n == len(f)
    seen = {}
    for i in range(n):
        name, rows, cols, off = struct.unpack_from("<48sIIQ", blob, 32 + 64 * i)
        name = name.rstrip(b"\0").decode()
        assert off % 64 == 0
        seen[name] = np.frombuffer(blob, np.float32, rows * cols, off).reshape(rows, cols)
    for k, v in f.items():
        np.testing.assert_array_equal(seen[k], v)


def test_scale0to1_dropin(emd):
    a = np.array([[1.0, 3.0], [2.0, 5.0]])
    out = emd.scale0to1(a)
    assert out.dtype == np.float32 and out.tolist() == [[0.0, 0.5], [0.25, 1.0]]
    c = np.full((2, 2), 3.0)
    assert (emd.scale0to1(c) == 0.5).all() and (c == 0.5).all()  # in-place fill like DEN:690-691


def test_weight_exporter_covers_both_graph_variants():
    """layer tables match the oracle's creation-order inventories; a packed variant-B blob carries the stand-alone BN layers."""
    import importlib
    from oracle.net import layer_specs
    w = importlib.import_module("ai-cv-automation-elect-micr_b200.weights")
    for v in "AB":
        assert [tuple(t) for t in w.layer_table(v)] == [tuple(t) for t in layer_specs(v)]
    f = w.fold(w.init_reference_weights(0, "B"), "B")
    assert f["aspp_r6_post/scale"].shape == (1, 728) and f["aspp_image/shift"].shape == (1, 728) and "aspp_image/bias" not in f
    assert f["aspp_r12/dw"].shape == (9, 728) and f["aspp_r12/w"].shape == (728, 728)
    blob = w.pack(w.init_reference_weights(0, "B"), "B")
    assert blob[:8] == b"EMDW0001" and int.from_bytes(blob[12:16], "little") == 1


def test_preprocess_crop_values_match_the_reference_order(emd):
    """Denoiser.preprocess (DEN:632-643), the single-crop path: VALUES against the numpy restatement in oracle/wrapper.py
    (which spells out cv2.resize's INTER_LINEAR arithmetic) -- up- and down-scaling, float32 and float64, and the class
    file's order of operations: min-max runs before the NaN/Inf replacement, so one NaN flattens the crop to 0.5 and an Inf
    ends up as the only 1.0."""
    from oracle import wrapper as W
    den = importlib.import_module("ai-cv-automation-elect-micr_b200.denoiser")
    rng = np.random.default_rng(12)
    for shape, dtype in (((100, 80), np.float32), ((64, 64), np.float32), ((150, 333), np.float64), ((40, 200), np.float32)):
        img = (rng.random(shape) * 900 + 50).astype(dtype)
        got, ref = den.preprocess_crop(img.copy(), 64), W.preprocess_crop(img.copy(), 64)
        assert got.shape == ref.shape == (1, 64, 64, 1) and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 2e-6, shape
        assert got.min() == 0.0 and got.max() == 1.0
    img = rng.random((64, 64)).astype(np.float32)
    img[3, 4] = np.nan
    assert (den.preprocess_crop(img.copy(), 64) == 0.5).all() and (W.preprocess_crop(img.copy(), 64) == 0.5).all()
    img = rng.random((64, 64)).astype(np.float32)
    img[10, 20] = np.inf
    got = den.preprocess_crop(img.copy(), 64)
    np.testing.assert_array_equal(got, W.preprocess_crop(img.copy(), 64))
    assert got[0, 10, 20, 0] == 1.0 and (np.delete(got.ravel(), 10 * 64 + 20) == 0.0).all()
    const = np.full((32, 48), 7.0, np.float32)
    assert (den.preprocess_crop(const, 64) == 0.5).all()


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: ONE JSON line on stdout (everything else -- NCCL's banner, progress -- goes to stderr); the reference
    arm times the CPU port and needs no GPU.  One step, no warm-up: a few seconds."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoised 512x512 crops/s" and d["unit"] == "crops/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("batch 32 of 512x512")


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without work or output."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
