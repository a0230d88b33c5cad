"""GPU parity tests of the 16-bit tensor-core modes (tcgen05 implicit-GEMM kernel).

Three references per check:
  (a) the same layer run on the CUDA-core kernel with identical 16-bit operand values
      (emd_set_tensor_cores(0)): isolates the tcgen05 data path (descriptors, swizzle, TMEM epilogue);
  (b) the FP64 oracle, per layer, O(1) inputs: north-star tolerance 5e-3 for BF16 operands;
  (c) the oracle end to end, plain and with BF16 operand rounding emulated on the CPU
      (OracleNet.emulate_bf16): the part of the end-to-end distance that is rounding, not kernels.
"""
import numpy as np
import pytest

import torch

from conftest import TOL_16BIT, rel_l2
from test_gpu_parity import layer_io_table, oracle_acts

pytestmark = pytest.mark.gpu

S, N = 64, 2


@pytest.fixture(scope="module")
def setup(emd):
    from oracle.weights import make_w0, make_w1
    rng = np.random.default_rng(1234)
    crops = rng.random((N, S, S)).astype(np.float32)
    eng = emd.Engine(cropsize=S, max_batch=4)
    return dict(emd=emd, eng=eng, crops=crops, w0=make_w0(0), w1=make_w1(crops, seed=0))


# BF16 operand rounding (2^-9 per stored value; weights, depthwise result and layer input each rounded once) measured per
# layer against the FP64 oracle: these layers read the un-normalised 728-channel trunk late in the middle flow, or (aspp_image)
# a 2x2 pooled map whose BatchNorm statistics are degenerate at this test's 64x64 crop.  Same kernels, FP16 operands: all
# within the contract -- which is why FP16 is the contract mode and BF16 the opt-in fast mode (Precision).
KNOWN_BF16_OVER = {"mid6_0", "mid7_0", "mid8_0", "mid9_0", "mid10_0", "aspp_1x1", "aspp_r6", "aspp_r12", "aspp_r18", "aspp_image"}


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_every_layer_tensor_core(setup, mode):
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    _, acts = oracle_acts(setup["w1"], setup["crops"])
    n_tc = 0
    table = []
    worst_ab, worst_or = ("", 0.0), ("", 0.0)
    for layer, (i, r, o) in layer_io_table().items():
        res = None if r is None else acts[r]
        before = eng.tensor_core_launches
        eng.set_tensor_cores(True)
        got = eng.run_layer(layer, acts[i], res, mode=mode)
        used_tc = eng.tensor_core_launches > before
        n_tc += used_tc
        eng.set_tensor_cores(False)
        ab = eng.run_layer(layer, acts[i], res, mode=mode)
        eng.set_tensor_cores(True)
        e_ab, e_or = rel_l2(got, ab), rel_l2(got, acts[o])
        worst_ab = max(worst_ab, (layer, e_ab), key=lambda t: t[1])
        worst_or = max(worst_or, (layer, e_or), key=lambda t: t[1])
        table.append((layer, e_ab, e_or))
        assert e_ab <= 5e-4, f"{layer}: tcgen05 vs CUDA-core with the same operands: {e_ab:.3e}"
    print("\n".join(f"  {l:16s} A/B {a:.2e}  vs oracle {o:.2e}" for l, a, o in table))
    over = {l: o for l, _, o in table if o > TOL_16BIT}
    print(mode, "layers over the 5e-3 contract:", over)
    print(mode, "tensor-core layers:", n_tc, "worst A/B", worst_ab, "worst vs oracle", worst_or)
    if mode == "fp16":      # the contract mode: no exceptions
        assert not over, f"fp16 vs FP64 oracle over {TOL_16BIT}: {over}"
    else:                   # the fast mode: only the documented layers may cross the line, and only by rounding (A/B above stays <= 5e-4)
        assert set(over) <= KNOWN_BF16_OVER, f"bf16: unexpected layers over {TOL_16BIT}: {set(over) - KNOWN_BF16_OVER}"
        assert all(v <= (5 if l == "aspp_image" else 1.5) * TOL_16BIT for l, v in over.items()), over
    assert n_tc >= 60  # everything GEMM-class except the 1-channel stem and the 64->1 final conv


def test_network_fp16_end_to_end_contract(setup):
    """THE end-to-end contract (BASELINE.json north_star): rel-L2 <= 5e-3 against the oracle on the random-init weight
    set (W0), in the mode bench.py reports (FP16 operands, FP32 accumulate)."""
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w0"]))
    ref, _ = oracle_acts(setup["w0"], setup["crops"])
    out = eng.forward(setup["crops"], mode="fp16")
    e = rel_l2(out, ref)
    print(f"w0: fp16 vs FP64 oracle {e:.3e} (contract {TOL_16BIT})")
    assert np.isfinite(out).all() and out.min() >= 0 and out.max() <= 1
    assert e <= TOL_16BIT


@pytest.mark.parametrize("mode,wset", [("fp16", "w1"), ("bf16", "w0"), ("bf16", "w1")])
def test_network_16bit_end_to_end_is_rounding_only(setup, mode, wset):
    """Outside the contract line, reported: FP16 on the BN-calibrated W1 set, BF16 on both.  What IS asserted: the CUDA
    path sits where operand rounding alone puts the CPU oracle (OracleNet.emulate_dtype) -- the kernels add nothing."""
    from oracle.net import OracleNet
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup[wset]))
    ref, _ = oracle_acts(setup[wset], setup["crops"])
    emu = OracleNet(setup[wset], S)
    emu.emulate_dtype = torch.float16 if mode == "fp16" else torch.bfloat16
    emu_out = emu.forward(setup["crops"])
    out = eng.forward(setup["crops"], mode=mode)
    eng.set_tensor_cores(False)
    out_cc = eng.forward(setup["crops"], mode=mode)
    eng.set_tensor_cores(True)
    e_ref, e_emu, e_cc = rel_l2(out, ref), rel_l2(out, emu_out), rel_l2(out, out_cc)
    budget = rel_l2(emu_out, ref)
    print(f"{wset} {mode}: vs oracle {e_ref:.3e}; vs {mode}-emulating oracle {e_emu:.3e}; vs CUDA-core {mode} {e_cc:.3e}; "
          f"rounding budget (emulated vs exact oracle) {budget:.3e}")
    assert np.isfinite(out).all() and out.min() >= 0 and out.max() <= 1
    assert e_ref <= 1.5 * budget + 1e-3
    assert e_cc <= 0.5 * budget + 1e-3


@pytest.mark.xfail(strict=True, reason="BF16 operands cannot meet 5e-3 end to end: the CPU oracle with nothing but BF16 rounding "
                                       "emulated is 1.7e-2 (W0) from the exact oracle; FP16 is the contract mode (DESIGN.md)")
def test_network_bf16_end_to_end_contract(setup):
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w0"]))
    ref, _ = oracle_acts(setup["w0"], setup["crops"])
    assert rel_l2(eng.forward(setup["crops"], mode="bf16"), ref) <= TOL_16BIT


def test_first_generation_kernel_and_small_map_tiles(setup):
    """Maps that do not tile into 8 x 16 pixel blocks (the 8x8 and 4x4 maps of 64x64 crops) run on the block-tiled kernel with
    whole-row, multi-image M tiles; the first-generation tcgen05 kernel (cp.async im2col gather) is what remains behind
    set_option('fused', 0).  Both against the oracle per layer, and against each other."""
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    _, acts = oracle_acts(setup["w1"], setup["crops"])
    table = layer_io_table()
    try:
        for layer in ("cnn3", "cnn3_last", "cnn3_strided", "residual3", "mid4_0", "mid4_2", "aspp_r6", "aspp_pellet", "deconv2to1", "residual2_d"):
            i, r, o = table[layer]
            res = None if r is None else acts[r]
            t0, g0 = eng.counter("conv_fused_taps") + eng.counter("conv_fused_pair"), eng.counter("conv_tcgen05_gen1")
            a = eng.run_layer(layer, acts[i], res, mode="fp16")
            assert eng.counter("conv_fused_taps") + eng.counter("conv_fused_pair") > t0 and eng.counter("conv_tcgen05_gen1") == g0, layer
            eng.set_option("fused", 0)
            b = eng.run_layer(layer, acts[i], res, mode="fp16")
            assert eng.counter("conv_tcgen05_gen1") > g0, layer
            eng.set_option("fused", 1)
            assert rel_l2(a, acts[o]) <= TOL_16BIT and rel_l2(b, acts[o]) <= TOL_16BIT, (layer, rel_l2(a, acts[o]), rel_l2(b, acts[o]))
            assert rel_l2(a, b) <= 2e-4, (layer, rel_l2(a, b))
    finally:
        eng.set_option("fused", 1)


def test_tensor_core_shapes_s96_and_ragged_batch(setup):
    """96x96 crops (small_scans shape): 6x6 maps at stride 16, M tiles that straddle images, dilated
    taps that never land inside the map; batch 3 leaves ragged last tiles."""
    from oracle.weights import make_w1
    emd = setup["emd"]
    rng = np.random.default_rng(77)
    crops = rng.random((3, 96, 96)).astype(np.float32)
    w1 = make_w1(crops, seed=5)
    eng = emd.Engine(cropsize=96, max_batch=3)
    eng.load_weights(emd.weights.pack(w1))
    from oracle.net import OracleNet
    ref, _ = oracle_acts(w1, crops)
    for mode, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        emu = OracleNet(w1, 96)
        emu.emulate_dtype = dt
        budget = rel_l2(emu.forward(crops), ref)
        eng.set_tensor_cores(True)
        s0 = eng.counter("conv_cuda_core")
        out = eng.forward(crops, mode=mode)
        assert eng.counter("conv_cuda_core") == s0          # 24^2 / 12^2 / 6^2 maps still run on tensor-core kernels
        eng.set_tensor_cores(False)
        out_cc = eng.forward(crops, mode=mode)
        eng.set_tensor_cores(True)
        print(f"S=96 {mode}: vs oracle {rel_l2(out, ref):.3e} (rounding budget {budget:.3e}), tcgen05 vs CUDA-core {rel_l2(out, out_cc):.3e}")
        assert rel_l2(out, ref) <= 1.5 * budget + 1e-3
        assert rel_l2(out, out_cc) <= 0.5 * budget + 1e-3


def test_wide_layers_block_tiled_s256(emd):
    """256x256 crops: the 728-channel trunk is a 16x16 map, so the middle flow, the ASPP branches (three N tiles,
    dilated taps partly / wholly outside the map, the 3640-channel pellet) and every encoder/decoder layer run on the
    block-tiled tcgen05 kernel (emd_fused.cu), incl. its depthwise-producer mode.  Each layer against the FP64 oracle
    and against the CUDA-core kernel with the same 16-bit operand values."""
    from oracle.weights import make_w1
    S2 = 256
    rng = np.random.default_rng(4321)
    crops = rng.random((1, S2, S2)).astype(np.float32)
    w1 = make_w1(crops, seed=3)
    eng = emd.Engine(cropsize=S2, max_batch=1)
    eng.load_weights(emd.weights.pack(w1))
    ref, acts = oracle_acts(w1, crops)
    bad = []
    for layer, (i, r, o) in layer_io_table().items():
        res = None if r is None else acts[r]
        eng.set_tensor_cores(True)
        got = eng.run_layer(layer, acts[i], res, mode="bf16")
        eng.set_tensor_cores(False)
        ab = eng.run_layer(layer, acts[i], res, mode="bf16")
        eng.set_tensor_cores(True)
        e_ab, e_or = rel_l2(got, ab), rel_l2(got, acts[o])
        # aspp_image: pooled 8x8 map, BN statistics nearly degenerate (same exception as in test_every_layer_tensor_core)
        if e_ab > 5e-4 or e_or > (2.5e-2 if layer == "aspp_image" else 7e-3):
            bad.append((layer, e_ab, e_or))
    assert not bad, f"layer, tcgen05-vs-CUDA-core, vs-oracle: {bad}"
    out = eng.forward(crops, mode="bf16")
    eng.set_tensor_cores(False)
    out_cc = eng.forward(crops, mode="bf16")
    print("S=256 end to end: bf16 vs oracle", rel_l2(out, ref), "tcgen05 vs CUDA-core", rel_l2(out, out_cc))
    assert rel_l2(out, out_cc) <= 5e-2


def test_full_size_batch_properties(emd):
    """BASELINE.json's full crop size (512x512, batch 16), size-independent properties: outputs finite and in [0,1]
    (in-graph clip, DMG:534-538), a crop's result does not depend on its batch neighbours or its position in the
    batch, and host-buffer (chunked, copy-overlapped) and device-buffer calls give the same bits."""
    import torch
    rng = np.random.default_rng(99)
    crops = rng.random((16, 512, 512)).astype(np.float32)
    eng = emd.Engine(cropsize=512, max_batch=16)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(1)))
    a = eng.forward(crops, mode="fp16")
    assert np.isfinite(a).all() and a.min() >= 0 and a.max() <= 1
    c = eng.forward(crops[5:6], mode="fp16")
    np.testing.assert_array_equal(c[0], a[5])
    d = eng.forward(torch.from_numpy(crops[::-1].copy()).cuda(), mode="fp16")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy()[::-1], a)
    # workspace reuse must not change a bit: same pass with every activation in its own buffer (this caught an arena
    # lifetime bug: a depthwise input aliased with the output of the GEMM kernel that computes the depthwise on the fly)
    eng.set_keep_activations(True)
    k = eng.forward(crops, mode="fp16")
    eng.set_keep_activations(False)
    np.testing.assert_array_equal(k, a)
    # and the tcgen05 path (CTA-pair kernels on at this batch) against the CUDA-core path with the same 16-bit operand values
    assert eng.counter("conv_fused_pair") > 0
    eng.set_tensor_cores(False)
    cc = eng.forward(crops[:2], mode="fp16")
    eng.set_tensor_cores(True)
    assert rel_l2(a[:2], cc) <= TOL_16BIT


@pytest.mark.parametrize("crop", [96, 160])
def test_depthwise_kernel_variants_bit_identical_small_maps(emd, crop):
    """The stand-alone depthwise 3x3 of maps that do not tile into 8 x 16 pixel blocks has four kernels: register strips (default;
    6- or 8-column strips; 160x160 crops give 40 / 20 / 10 pixel maps: 8-column strips, ragged last strips, several row blocks), shared-memory tiles, one-column strips and
    one thread per pixel.  Same tap order everywhere: the layer outputs must be bit-identical, and within the contract of the
    FP64 oracle."""
    from oracle.weights import make_w1
    rng = np.random.default_rng(100 + crop)
    crops = rng.random((3 if crop < 128 else 2, crop, crop)).astype(np.float32)
    w1 = make_w1(crops, seed=11)
    eng = emd.Engine(cropsize=crop, max_batch=3)
    eng.load_weights(emd.weights.pack(w1))
    _, acts = oracle_acts(w1, crops)
    table = layer_io_table()
    variants = [("dw_reg", {}), ("dw_tile", {"dw_reg": 0}), ("dw_strip", {"dw_reg": 0, "dw_tile": 0}),
                ("per pixel", {"dw_reg": 0, "dw_tile": 0, "dw_strip": 0})]
    try:
        for layer in ("cnn2", "cnn2_last", "cnn3", "cnn3_last", "cnn4_1", "mid3_0", "mid7_2", "deconv2_0", "deconv2_1"):
            i, r, o = table[layer]
            res = None if r is None else acts[r]
            outs = []
            for name, opts in variants:
                for k, v in opts.items():
                    eng.set_option(k, v)
                outs.append(eng.run_layer(layer, acts[i], res, mode="fp16"))
                for k in opts:
                    eng.set_option(k, 1)
            for (name, _), got in zip(variants[1:], outs[1:]):
                assert np.array_equal(outs[0], got), (crop, layer, name, rel_l2(outs[0], got))
            assert rel_l2(outs[0], acts[o]) <= TOL_16BIT, (crop, layer, rel_l2(outs[0], acts[o]))
    finally:
        for k in ("dw_reg", "dw_tile", "dw_strip"):
            eng.set_option(k, 1)


def test_depthwise_register_strips_match_tma_kernel_on_tiled_maps(emd):
    """On maps that DO tile (32^2 and 64^2 maps of 728 / 256 channels at the benchmark crop size: the blocks with three N tiles) the TMA-fed depthwise
    kernel is the default and the register-strip kernel an option (`dw_reg_all`): bit-identical layer outputs, random inputs."""
    eng = emd.Engine(cropsize=512, max_batch=2)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(3)))
    rng = np.random.default_rng(5)
    try:
        for layer, shape in (("mid5_1", (2, 32, 32, 728)), ("cnn3_last", (2, 64, 64, 728)), ("cnn3", (2, 64, 64, 256))):
            x = (rng.random(shape, dtype=np.float32) * 2).astype(np.float32)
            a = eng.run_layer(layer, x, None, mode="fp16")
            eng.set_option("dw_reg_all", 1)
            b = eng.run_layer(layer, x, None, mode="fp16")
            eng.set_option("dw_reg_all", 0)
            assert np.isfinite(a).all() and np.array_equal(a, b), (layer, rel_l2(a, b))
    finally:
        eng.set_option("dw_reg_all", 0)


def test_no_layer_reads_what_the_pass_has_not_written(emd):
    """Option `poison`: the whole activation workspace is filled with NaNs before every pass.  A layer that read padding channels
    of the 768-pitch trunk tensors, a halo outside its tensor, or a tile whose producer had not finished would turn the output
    into NaNs; identical passes normally hide such reads behind the previous pass's identical values.  Host (sliced, copy-
    overlapped), device (direct, then graph replay) and half-batch passes, FP16 and FP32; results bit-identical to clean passes."""
    import torch
    rng = np.random.default_rng(3)
    crops = rng.random((16, 512, 512)).astype(np.float32)
    eng = emd.Engine(cropsize=512, max_batch=16)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(1)))
    a = eng.forward(crops, mode="fp16")
    f = eng.forward(crops[:2], mode="fp32")
    try:
        eng.set_option("poison", 1)
        x = torch.from_numpy(crops).cuda()
        outs = [eng.forward(crops, mode="fp16")]
        for _ in range(3):                                   # direct, direct, graph replay
            o = eng.forward(x, mode="fp16")
            torch.cuda.synchronize()
            outs.append(o.cpu().numpy())
        for o in outs:
            assert not np.isnan(o).any()
            np.testing.assert_array_equal(o, a)
        np.testing.assert_array_equal(eng.forward(crops[:8], mode="fp16"), a[:8])
        np.testing.assert_array_equal(eng.forward(crops[:2], mode="fp32"), f)
    finally:
        eng.set_option("poison", 0)


def test_first_pass_over_a_fresh_large_workspace(emd):
    """Regression: the first pass over a newly allocated workspace of >= 28 GB (keep mode at max_batch 16: every activation in its
    own buffer) intermittently faulted or returned wrong tiles until the workspace was cleared once after allocation
    (plan_arena).  Five fresh engines, first keep pass each, bit-identical to the arena-planned pass."""
    rng = np.random.default_rng(99)
    crops = rng.random((8, 512, 512)).astype(np.float32)
    blob = emd.weights.pack(emd.weights.init_reference_weights(1))
    for it in range(5):
        eng = emd.Engine(cropsize=512, max_batch=16)
        eng.load_weights(blob)
        a = eng.forward(crops, mode="fp16")
        eng.set_keep_activations(True)
        assert eng.counter("workspace_bytes") > 25e9
        k = eng.forward(crops, mode="fp16")
        np.testing.assert_array_equal(k, a)
        eng.close()


def test_forked_side_convs_bit_identical(emd):
    """Option `fork_sms`: the decoder's 1x1 residual convs run beside the separable block that reads the same tensor, on their own
    share of the SMs and a second stream.  Same kernels, same arithmetic: bit-identical outputs for host (sliced, half-batch
    tail), device (direct, then graph replay) and small-batch passes, also over a NaN-poisoned workspace (an ordering mistake
    between the two streams would read unwritten data)."""
    import torch
    rng = np.random.default_rng(8)
    crops = rng.random((16, 512, 512)).astype(np.float32)
    eng = emd.Engine(cropsize=512, max_batch=16)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(2)))
    a = eng.forward(crops, mode="fp16")
    x = torch.from_numpy(crops).cuda()
    try:
        eng.set_option("fork_sms", 48)
        eng.set_option("poison", 1)
        np.testing.assert_array_equal(eng.forward(crops, mode="fp16"), a)
        for _ in range(3):
            o = eng.forward(x, mode="fp16")
            torch.cuda.synchronize()
            np.testing.assert_array_equal(o.cpu().numpy(), a)
        np.testing.assert_array_equal(eng.forward(crops[:3], mode="fp16"), a[:3])
        eng.set_option("poison", 0)
        ob = eng.forward(x, mode="bf16")
        torch.cuda.synchronize()
        np.testing.assert_array_equal(eng.forward(crops, mode="bf16"), ob.cpu().numpy())
    finally:
        eng.set_option("fork_sms", 0)
        eng.set_option("poison", 0)
