"""CPU tests of the oracle itself: TF-semantic ops against closed-form mini-cases and identities
(SURVEY.md App. A), and the committed golden fixtures.  The reference has no tests for this path."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import net as O
from oracle import wrapper as W
from oracle.weights import make_w0, make_w1

GOLD = os.path.join(os.path.dirname(__file__), "golden")
f64 = torch.float64


def test_same_padding_amounts():  # App. A.1
    assert O.same_pad(512, 3, 2) == (0, 1)
    assert O.same_pad(96, 3, 2) == (0, 1)
    assert O.same_pad(32, 3, 1, 6) == (6, 6)
    assert O.same_pad(32, 3, 1, 18) == (18, 18)
    assert O.same_pad(512, 1, 2) == (0, 0)
    assert O.same_pad(5, 3, 2) == (1, 1)


def test_stride2_depthwise_window_is_2i_2i2():  # App. A.1: window of output i covers inputs 2i..2i+2
    x = torch.arange(8, dtype=f64).view(1, 1, 1, 8).expand(1, 1, 8, 8).contiguous()
    w = torch.zeros(3, 3, 1, 1, dtype=f64)
    w[0, 2, 0, 0] = 1.0  # picks input (2i+0, 2j+2)
    y = O.depthwise3x3(x, w, stride=2)
    assert y.shape == (1, 1, 4, 4)
    assert y[0, 0, 0].tolist() == [2.0, 4.0, 6.0, 0.0]  # last window runs into the 1-pixel pad after


def test_conv1x1_stride2_samples_even_pixels():  # DMG:365-370
    x = torch.arange(16, dtype=f64).view(1, 1, 4, 4)
    y = O.conv2d(x, torch.ones(1, 1, 1, 1, dtype=f64), None, stride=2)
    assert y.flatten().tolist() == [0.0, 2.0, 8.0, 10.0]


def test_transposed_conv_closed_form_1d():  # App. A.4
    x = torch.tensor([1.0, 2.0, 3.0], dtype=f64).view(1, 1, 1, 3)
    k = torch.zeros(3, 3, 1, 1, dtype=f64)
    k[0, :, 0, 0] = torch.tensor([10.0, 100.0, 1000.0])  # only the first kernel row -> output row 0
    y = O.conv2d_transpose_s2(x, k, torch.zeros(1, dtype=f64))
    assert y.shape == (1, 1, 2, 6)
    # out[2j] = in[j] w0 + in[j-1] w2 ; out[2j+1] = in[j] w1
    assert y[0, 0, 0].tolist() == [10.0, 100.0, 1020.0, 200.0, 2030.0, 300.0]


def test_transposed_conv_is_adjoint_of_same_stride2_conv():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 6, 10, dtype=f64, generator=g)
    k = torch.randn(3, 3, 5, 3, dtype=f64, generator=g)  # [3,3,Cout=5,Cin=3]
    z = torch.randn(2, 5, 12, 20, dtype=f64, generator=g)
    lhs = (O.conv2d_transpose_s2(x, k, torch.zeros(5, dtype=f64)) * z).sum()
    rhs = (x * O.conv2d(z, k, None, stride=2)).sum()
    assert abs(float(lhs - rhs)) < 1e-9 * abs(float(lhs))


def test_legacy_bilinear_x4_closed_form():  # App. A.5: out[4i+r] = (1-r/4) in[i] + (r/4) in[min(i+1,n-1)]
    x = torch.tensor([0.0, 4.0, 8.0], dtype=f64).view(1, 1, 1, 3).expand(1, 1, 3, 3).contiguous()
    y = O.resize_bilinear_legacy(x, 12, 12)
    assert y[0, 0, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 8]
    # and it is NOT torch's half-pixel interpolate
    t = torch.nn.functional.interpolate(x, size=(12, 12), mode="bilinear", align_corners=False)
    assert not torch.allclose(t, y)
    assert O.resize_bilinear_legacy(x, 3, 3) is x  # same size = identity (DEN:199)


def test_avg_pool_and_relu6():
    x = torch.arange(16, dtype=f64).view(1, 1, 4, 4)
    assert O.avg_pool_2x2(x).flatten().tolist() == [2.5, 4.5, 10.5, 12.5]
    assert O.relu6(torch.tensor([-1.0, 3.0, 7.0])).tolist() == [0.0, 3.0, 6.0]


def test_fresh_batchnorm_is_divide_by_sqrt_1_001():  # App. A.3
    p = make_w0(0)
    n = O.OracleNet(p, 32)
    x = torch.ones(1, 64, 2, 2)
    assert torch.allclose(n._bn(x, "cnn0/bn1"), x / np.sqrt(1.001))


def test_layer_inventory_counts():  # 54 separable, 14 dense, 2 transposed convs, 38.8 M params
    specs = O.layer_specs("A")
    kinds = [s[1] for s in specs]
    assert kinds.count("sep") == 54 and kinds.count("conv") == 14 and kinds.count("deconv") == 2
    assert sum(int(np.prod(s)) for s in O.param_shapes("A").values()) == 38772462
    assert [s[1] for s in O.layer_specs("B")].count("sep") == 57


def test_w0_matches_committed_digest():
    import hashlib
    w0 = make_w0(0)
    h = hashlib.sha256()
    for k in sorted(w0):
        h.update(k.encode()); h.update(np.ascontiguousarray(w0[k]).tobytes())
    gold = json.load(open(os.path.join(GOLD, "w0_digest.json")))
    assert gold["n_params"] == 38772462
    assert h.hexdigest() == gold["sha256"]


def test_network_golden_s64():
    g = np.load(os.path.join(GOLD, "net_s64.npz"))
    crops = g["crops"]
    net = O.OracleNet(make_w1(crops, seed=0), 64)
    net.collect = True
    out = net.forward(crops)
    assert out.shape == (2, 64, 64) and out.min() >= 0.0 and out.max() <= 1.0
    assert rel_l2(out, g["out"]) < 1e-4  # weights are re-derived by calibration; FP32 reductions may reorder
    for name, st in zip(g["layer_names"], g["layer_stats"]):
        a = net.acts[str(name)]
        assert abs(a.mean() - st[0]) < 1e-3 + 1e-3 * abs(st[0]), name
        assert abs(a.std() - st[1]) < 1e-3 + 1e-3 * abs(st[1]), name
    # every activation is O(1) with W1 (that is what W1 is for, App. E.3)
    for k in ("mid5_2", "aspp_r18", "dec0"):
        assert 0.1 < net.acts[k].std() < 5.0


def test_known_answer_vectors_s96():
    """SURVEY 8(c) pin 3: committed known-answer vectors for a 96x96 W1 case -- sum, sum of squares and 32 sampled values
    of EVERY named activation of the graph, plus the whole output.  (Weights are re-derived by calibration, so the
    comparison carries a small tolerance; a changed op, padding, concat order or fold shows up as O(1).)"""
    sys.path.insert(0, GOLD)
    from make_golden import kat_positions, kat_s96_inputs
    g = np.load(os.path.join(GOLD, "net_s96_kat.npz"))
    crops = kat_s96_inputs()
    net = O.OracleNet(make_w1(crops, seed=96), 96, dtype=torch.float64)
    net.collect = True
    out = net.forward(crops)
    assert rel_l2(out, g["out"]) < 1e-4
    assert len(g["layer_names"]) == 146
    for name, sums, samp, shape in zip(g["layer_names"], g["sums"], g["samples"], g["shapes"]):
        a = net.acts[str(name)]
        assert tuple(a.shape) == tuple(shape), name
        assert abs(a.sum() - sums[0]) <= 1e-4 * (abs(sums[0]) + np.sqrt(sums[1] * a.size) * 1e-2), name
        assert abs((a.astype(np.float64) ** 2).sum() - sums[1]) <= 2e-4 * sums[1] + 1e-12, name
        got = a.reshape(-1)[kat_positions(str(name), a.size)]
        assert np.linalg.norm(got - samp) <= 2e-4 * np.linalg.norm(samp) + 1e-6, name


def test_f64_switch_agrees():
    rng = np.random.default_rng(0)
    crops = rng.random((2, 64, 64)).astype(np.float32)  # (1 crop of 32^2 leaves BN statistics degenerate)
    p = make_w1(crops, seed=3)
    a = O.OracleNet(p, 64).forward(crops)
    b = O.OracleNet(p, 64, dtype=torch.float64).forward(crops)
    assert rel_l2(a, b) < 2e-5


def test_variant_b_runs_and_differs():
    rng = np.random.default_rng(0)
    crops = rng.random((2, 64, 64)).astype(np.float32)
    out = O.OracleNet(make_w1(crops, seed=0, variant="B"), 64, variant="B").forward(crops)
    assert out.shape == (2, 64, 64) and np.isfinite(out).all()


# ---- wrapper ---------------------------------------------------------------------------------

def test_tile_plans_golden():
    plans = json.load(open(os.path.join(GOLD, "tile_plans.json")))
    for key, g in plans.items():
        parts = [int(v) for v in key.split("/")]
        size, crop, ov = (parts + [512, 80])[:3] if len(parts) == 1 else parts
        assert W.tile_origins(size, crop, ov) == g["origins"], key
        c = W.coverage_counts(size, crop, ov)
        for a, b, n in g["runs"]:
            assert (c[a:b] == n).all(), key
        assert c.min() >= 1
    assert len(plans["2048"]["origins"]) ** 2 == 25 and len(plans["4096"]["origins"]) ** 2 == 100


def test_round_half_even_origin():  # App. D-2: np.round, not lround
    # size 1296 -> num 4, len 324.0 (exact); craft a .5 case: size 433*... use direct check of np.round semantics
    assert int(np.round(0.5)) == 0 and int(np.round(1.5)) == 2 and int(np.round(2.5)) == 2
    assert W.tile_origins(1000) == [min(int(np.round(i * 1000 / 3)), 488) for i in range(3)]


def test_scale0to1_and_normalise():
    a = np.array([[1.0, 3.0], [2.0, 5.0]], np.float32)
    assert W.scale0to1(a.copy()).tolist() == [[0.0, 0.5], [0.25, 1.0]]
    c = np.full((3, 3), 7.0, np.float32)
    assert (W.scale0to1(c) == 0.5).all()
    b = np.array([[np.nan, 0.0], [np.inf, 2.0]], np.float32)
    out = W.normalise(b)
    assert out.tolist() == [[0.25, 0.0], [0.25, 1.0]]  # 0.5 substituted BEFORE min/max (DMG:853-858)
    assert np.isnan(b[0, 0])  # works on a copy
    assert W.normalise(np.random.default_rng(0).random((4, 4))).dtype == np.float32


def test_stitch_identity_network_recovers_image():
    rng = np.random.default_rng(0)
    img = rng.random((150, 200)).astype(np.float32)
    out = W.denoise(img, lambda c: c, preprocess=False, postprocess=False, overlap=8, crop=64)
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, img.astype(np.float64))  # weights 1, 1/2, 1/4 are exact


def test_stitch_weights_are_reciprocal_counts():
    ys = W.tile_origins(600)
    ones = np.ones((len(ys) ** 2, 512, 512), np.float32)
    out = W.stitch(ones * 0.75, ys, ys, 600, 600)
    np.testing.assert_array_equal(out, 0.75)


def test_image_smaller_than_crop_rejected():
    with pytest.raises(ValueError):
        W.tile_origins(511)


@pytest.mark.parametrize("variant,size", [("A", 128), ("B", 64)])
def test_second_independent_restatement_agrees(variant, size):
    """oracle/naive.py is a second restatement of the reference graph in plain numpy float64, written op by op from the
    definitions (explicit SAME padding and tap loops, sub-pixel closed form of the transposed conv, per-pixel legacy bilinear,
    un-fused BatchNorm).  On calibrated weights (every layer O(1), so no branch hides behind another) the two restatements must
    agree to rounding -- a transcription slip in either one (tap order, padding side, concat order, BN order, kernel layout,
    a shifted transposed conv) is a difference of order one.  128 x 128 crops put in-bounds dilated taps on the 8 x 8 ASPP map."""
    from oracle.naive import NaiveNet
    from oracle.net import OracleNet
    from oracle.weights import make_w1
    rng = np.random.default_rng(31)
    crops = rng.random((2, size, size)).astype(np.float32)
    w1 = make_w1(crops, seed=9, variant=variant)
    ref = OracleNet(w1, size, variant=variant, dtype=torch.float64).forward(crops)
    got = NaiveNet(w1, size, variant=variant).forward(crops)
    assert ref.shape == got.shape == (2, size, size)
    assert float(np.abs(ref).max()) > 1e-2 and float((ref > 0).mean()) > 0.05      # a live output, not an all-zero clip
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err <= 1e-10, err
