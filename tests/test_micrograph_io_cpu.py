"""CPU tests of the micrograph TIFF codec and the file-streaming front-end (micrograph_io.py): round trips, an independent
decoder / encoder (Pillow) on both sides, strip / byte-order / sample-type / PackBits variants built by hand, error behaviour
of ``load_image`` (misc_py/denoiser-multi-gpu.py:800-814) and rank ownership of ``denoise_files``."""
import os
import struct

import numpy as np
import pytest


def hand_tiff(path, img, bo="<", rows_per_strip=None, packbits=False, fmt=3, spp=1):
    """Minimal TIFF writer used only here: several strips, either byte order, optional PackBits, IFD BEFORE the pixel data."""
    a = np.ascontiguousarray(img).astype(img.dtype.newbyteorder(bo))
    h, w = a.shape[0], a.shape[1]
    rps = rows_per_strip or h
    strips = [a[r:r + rps].tobytes() for r in range(0, h, rps)]
    if packbits:   # literal runs of up to 128 bytes
        strips = [b"".join(bytes([len(s[i:i + 128]) - 1]) + s[i:i + 128] for i in range(0, len(s), 128)) for s in strips]
    n = len(strips)
    entries = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, a.dtype.itemsize * 8), (259, 3, 1, 32773 if packbits else 1), (262, 3, 1, 1),
               (273, 4, n, None), (277, 3, 1, spp), (278, 3, 1, rps), (279, 4, n, None), (339, 3, 1, fmt)]
    ifd_off = 8
    ifd_len = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_len
    data_off = extra_off + (8 * n if n > 1 else 0)
    offs, pos = [], data_off
    for s in strips:
        offs.append(pos)
        pos += len(s)
    out = (b"II" if bo == "<" else b"MM") + struct.pack(bo + "HI", 42, ifd_off) + struct.pack(bo + "H", len(entries))
    for tag, typ, count, value in entries:
        if tag in (273, 279):
            vals = offs if tag == 273 else [len(s) for s in strips]
            field = struct.pack(bo + "I", vals[0]) if n == 1 else struct.pack(bo + "I", extra_off + (0 if tag == 273 else 4 * n))
        else:
            field = struct.pack(bo + "HH", value, 0) if bo == "<" else struct.pack(bo + "HH", value, 0)
        out += struct.pack(bo + "HHI", tag, typ, count) + field
    out += struct.pack(bo + "I", 0)
    if n > 1:
        out += struct.pack(bo + "I" * n, *offs) + struct.pack(bo + "I" * n, *[len(s) for s in strips])
    out += b"".join(strips)
    open(path, "wb").write(out)


def test_float_tiff_round_trip_and_pillow_agree(emd, tmp_path):
    io = emd.micrograph_io
    rng = np.random.default_rng(5)
    img = (rng.standard_normal((37, 53)) * 1e3).astype(np.float32)
    img[3, 4] = np.nan
    img[5, 6] = np.inf
    p = str(tmp_path / "a.tif")
    io.write_tiff(p, img)
    back = io.read_tiff(p)
    assert back.dtype == np.float32 and back.tobytes() == img.tobytes()          # bit-exact, NaN payloads included
    Image = pytest.importorskip("PIL.Image")
    with Image.open(p) as im:                                                    # an independent decoder reads what we write
        assert im.mode == "F" and im.size == (53, 37)
        assert np.asarray(im, np.float32).tobytes() == img.tobytes()
    q = str(tmp_path / "b.tif")
    Image.fromarray(img, mode="F").save(q)                                       # ... and we read what an independent encoder writes
    assert io.read_tiff(q).tobytes() == img.tobytes()


@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("rows_per_strip,packbits", [(None, False), (5, False), (7, True)])
def test_strips_byte_order_packbits(emd, tmp_path, bo, rows_per_strip, packbits):
    img = np.arange(33 * 20, dtype=np.float32).reshape(33, 20) / 7
    p = str(tmp_path / "s.tif")
    hand_tiff(p, img, bo=bo, rows_per_strip=rows_per_strip, packbits=packbits)
    assert np.array_equal(emd.micrograph_io.read_tiff(p), img)


def test_integer_and_multisample_inputs_become_float32(emd, tmp_path):
    io = emd.micrograph_io
    u16 = (np.arange(12 * 9).reshape(12, 9) * 500).astype(np.uint16)
    hand_tiff(str(tmp_path / "u16.tif"), u16, fmt=1)
    got = io.read_tiff(str(tmp_path / "u16.tif"))
    assert got.dtype == np.float32 and np.array_equal(got, u16.astype(np.float32))
    rgb = np.stack([np.full((6, 4), v, np.uint8) for v in (10, 20, 30)], -1)
    hand_tiff(str(tmp_path / "rgb.tif"), rgb.reshape(6, 12), fmt=1, spp=3)       # chunky RGB: first sample kept
    open(str(tmp_path / "rgb.tif"), "r+b").close()
    # the hand writer stores width*spp as the width; fix the width tag to 4 pixels
    raw = bytearray(open(str(tmp_path / "rgb.tif"), "rb").read())
    raw[8 + 2 + 8:8 + 2 + 10] = struct.pack("<H", 4)
    open(str(tmp_path / "rgb.tif"), "wb").write(bytes(raw))
    assert np.array_equal(io.read_tiff(str(tmp_path / "rgb.tif")), np.full((6, 4), 10, np.float32))


def test_load_image_failure_semantics(emd, tmp_path, capsys):
    io = emd.micrograph_io
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    img = io.load_image(str(tmp_path / "junk.tif"))
    assert img.shape == (512, 512) and img.dtype == np.float32 and np.all(img == 0.5)   # DMG:806-809
    assert "Image read failed" in capsys.readouterr().out
    with pytest.raises(ValueError):
        io.read_tiff(str(tmp_path / "junk.tif"))
    with pytest.raises(ValueError):
        io.write_tiff(str(tmp_path / "x.tif"), np.zeros((2, 2, 2), np.float32))


class FakeDenoiser:
    """Stands in for Denoiser on a CPU box: records calls; 'denoises' by clipping min-max normalised input."""
    def __init__(self):
        self.calls = []

    def denoise(self, img, overlap=80):
        self.calls.append((img.shape, overlap))
        lo, hi = float(img.min()), float(img.max())
        return ((img - lo) / max(hi - lo, 1e-30)).astype(np.float64)


def test_denoise_files_rank_ownership_and_outputs(emd, tmp_path):
    io = emd.micrograph_io
    rng = np.random.default_rng(0)
    paths = []
    for k in range(5):
        p = str(tmp_path / f"mic{k}.tif")
        io.write_tiff(p, rng.random((16 + k, 24)).astype(np.float32))
        paths.append(p)
    seen = []
    out = {}
    for rank in range(2):
        d = FakeDenoiser()
        done = io.denoise_files(d, paths, str(tmp_path / "out"), rank=rank, world=2, overlap=8,
                                on_result=lambda k, dst, im: seen.append(k))
        assert [k for k, _ in done] == list(range(rank, 5, 2))
        assert d.calls == [((16 + k, 24), 8) for k in range(rank, 5, 2)]
        out.update(dict(done))
    assert sorted(seen) == list(range(5))
    for k, dst in out.items():
        res = io.read_tiff(dst)
        assert os.path.basename(dst) == f"mic{k}_denoised.tif" and res.shape == (16 + k, 24)
        assert res.min() == 0.0 and res.max() == 1.0
    # a file that cannot be decoded surfaces as an error at the caller, not a hang of the reader thread
    (tmp_path / "bad.tif").write_bytes(b"II*\0garbage")
    with pytest.raises(Exception):
        io.denoise_files(FakeDenoiser(), [str(tmp_path / "bad.tif")], None)
