"""CPU tests of the TensorFlow-checkpoint importer (tfckpt.py): table/proto parsing against a format-level writer, the
TF-1 variable naming of both graph variants, and that the imported parameters fold to the same blob as the originals."""
import os

import numpy as np
import pytest

from tf_bundle_writer import write_checkpoint


def as_tf_variables(emd, params, variant):
    names = emd.tfckpt.tf_variable_names(variant)
    assert len({t for t, _ in names}) == len(names) == len(params)          # one TF variable per parameter, no collisions
    v = {t: params[p] for t, p in names}
    # what a trainer's Saver also stores (DMG:1149): optimizer slots and the step counter -- ignored by the importer
    v["nn/SeparableConv2d/depthwise_weights/Momentum"] = np.zeros((3, 3, 1, 1), np.float32)
    v["global_step"] = np.array(123456, np.int64)
    return v


@pytest.mark.parametrize("variant", ["A", "B"])
def test_checkpoint_round_trip(emd, tmp_path, variant):
    params = emd.weights.init_reference_weights(3, variant)
    rng = np.random.default_rng(0)
    for k in params:                                                          # non-trivial BN statistics and biases
        if k.endswith(("/beta", "/mean", "/bias")):
            params[k] = rng.normal(0, 0.1, params[k].shape).astype(np.float32)
        elif k.endswith(("/gamma", "/var")):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
    prefix = str(tmp_path / "ckpt" / "model.ckpt-42")
    write_checkpoint(prefix, as_tf_variables(emd, params, variant), num_shards=2 if variant == "B" else 1)

    assert emd.tfckpt.latest_checkpoint(str(tmp_path / "ckpt")) == prefix
    raw = emd.tfckpt.read_checkpoint(prefix)
    assert raw["global_step"] == 123456 and raw["global_step"].dtype == np.int64
    for loc in (str(tmp_path / "ckpt"), prefix):                              # directory (DEN:626) or explicit prefix
        got = emd.tfckpt.load_params(loc, variant)
        assert sorted(got) == sorted(params)
        for k in params:
            assert got[k].dtype == np.float32 and np.array_equal(got[k], params[k]), k
    assert emd.weights.pack(got, variant) == emd.weights.pack(params, variant)


def test_variable_names_follow_tf1_auto_naming(emd):
    a = dict((p, t) for t, p in emd.tfckpt.tf_variable_names("A"))
    # first separable block, its nested normaliser and the stand-alone BatchNorm after it (DMG:250-276, 220-223)
    assert a["cnn0/dw"] == "nn/SeparableConv2d/depthwise_weights"
    assert a["cnn0/bn1/gamma"] == "nn/SeparableConv2d/BatchNorm/gamma"
    assert a["cnn0/bn2/mean"] == "nn/BatchNorm/moving_mean"
    assert a["cnn0_last/pw"] == "nn/SeparableConv2d_1/pointwise_weights"
    # 4th layer created: the tf.layers residual conv, then its BatchNorm_3 (three separable blocks came first)
    assert a["residual0/kernel"] == "nn/conv2d/kernel" and a["residual0/bn/beta"] == "nn/BatchNorm_3/beta"
    assert a["residual1/bias"] == "nn/conv2d_1/bias"
    assert a["aspp_r12/kernel"] == "nn/mediumRate/kernel" and a["aspp_pellet/bias"] == "nn/pellet/bias"
    assert a["deconv2to1/tkernel"] == "nn/conv2d_transpose/kernel" and a["deconv1to0/tkernel"] == "nn/conv2d_transpose_1/kernel"
    # 4 encoder residuals + 3 decoder residuals, then the final conv: conv2d_7
    assert a["final/kernel"] == "nn/conv2d_7/kernel"
    b = dict((p, t) for t, p in emd.tfckpt.tf_variable_names("B"))
    assert b["residual0/kernel"] == "nn/Conv/weights" and b["aspp_1x1/bias"] == "nn/Conv_4/biases"
    assert b["aspp_r6/dw"] == "nn/SeparableConv2d_48/depthwise_weights"      # 12 encoder + 3 + 33 middle-flow blocks precede it
    assert b["deconv2to1/tkernel"] == "nn/Conv2d_transpose/weights"
    n_bn = sum(1 for t in b.values() if t.endswith("/gamma") and "/SeparableConv2d" not in t)
    assert b["final/bn/var"] == f"nn/BatchNorm_{n_bn - 1}/moving_variance"


def test_mismatches_are_reported(emd, tmp_path):
    params = emd.weights.init_reference_weights(0, "A")
    v = as_tf_variables(emd, params, "A")
    del v["nn/pellet/bias"]
    v["nn/conv2d_7/kernel"] = np.zeros((1, 1, 64, 1), np.float32)            # wrong kernel size for the final 3x3 conv
    prefix = str(tmp_path / "model.ckpt")
    write_checkpoint(prefix, v)
    with pytest.raises(ValueError, match="does not match the graph"):
        emd.tfckpt.load_params(prefix, "A")
    got, problems = emd.tfckpt.params_from_variables(emd.tfckpt.read_checkpoint(prefix), "A", strict=False)
    assert len(problems) == 2 and "aspp_pellet/bias" not in got and "final/kernel" not in got
    with pytest.raises(FileNotFoundError):
        emd.tfckpt.load_params(str(tmp_path / "nothing_here"), "A")
    open(tmp_path / "junk.index", "wb").write(b"\x00" * 100)
    with pytest.raises(ValueError, match="bad table magic"):
        emd.tfckpt.read_index(str(tmp_path / "junk.index"))


def test_diagnose_reports_mismatches_per_variant(tmp_path):
    """tfckpt.diagnose: a variant-A checkpoint lines up with graph A and reports what graph B would miss; strict=False returns the
    problem list instead of raising (what to run on the first real checkpoint)."""
    import importlib
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    params = emd.weights.init_reference_weights(0)
    variables = {t: params[p] for t, p in emd.tfckpt.tf_variable_names("A")}
    variables["nn/extra/unrelated"] = np.zeros(3, np.float32)
    del variables[next(t for t, p in emd.tfckpt.tf_variable_names("A") if p == "mid3_1/pw")]
    write_checkpoint(str(tmp_path / "model.ckpt-3"), variables)
    rep = emd.tfckpt.diagnose(str(tmp_path))
    assert len(rep["A"]["problems"]) == 1 and "mid3_1/pw" in rep["A"]["problems"][0]
    assert rep["A"]["unused"] == ["nn/extra/unrelated"]
    assert len(rep["B"]["problems"]) > 3
    got, problems = emd.tfckpt.load_params(str(tmp_path), "A", strict=False)
    assert len(problems) == 1 and "mid3_0/pw" in got and "mid3_1/pw" not in got
    with pytest.raises(ValueError, match="does not match the graph"):
        emd.tfckpt.load_params(str(tmp_path), "A")
