"""Micrograph file I/O and the streaming front-end of the denoiser.

The reference's micrographs are single-channel 32-bit IEEE-float TIFFs -- written by ``DM3stoTIFs-batch/DM3stoTIFs.m:23-34``
(uncompressed, ``SampleFormat = IEEEFP``, ``BitsPerSample = 32``, ``MinIsBlack``, chunky) and read with
``imread(addr, mode='F')`` in ``load_image`` (misc_py/denoiser-multi-gpu.py:800-814).  ``scipy.misc.imread`` no longer
exists, so this module carries a small baseline-TIFF codec of its own (numpy only): strips, either byte order,
uncompressed or PackBits, 8/16/32-bit unsigned, 16/32-bit signed and 32/64-bit float samples, first channel of chunky
multi-sample images -- everything converted to float32 like ``mode='F'``.

``denoise_files`` is BASELINE.json config 4 from disk: file k is owned by rank k % world (``sharding.shard_indices``); a
reader thread decodes the next micrographs while the GPU works on the current one and a writer thread encodes finished ones,
so the device never waits for the file system.
"""
from __future__ import annotations

import os
import queue
import struct
import threading
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import sharding as _sharding

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 8: ("h", 2), 9: ("i", 4), 11: ("f", 4),
          12: ("d", 8), 16: ("Q", 8)}
# (SampleFormat, BitsPerSample) -> numpy dtype character; SampleFormat 1 = unsigned, 2 = signed, 3 = IEEE float
_SAMPLES = {(1, 8): "u1", (1, 16): "u2", (1, 32): "u4", (2, 16): "i2", (2, 32): "i4", (3, 32): "f4", (3, 64): "f8"}


def _unpackbits(src: bytes, expected: int) -> bytes:
    """TIFF compression 32773 (PackBits)."""
    out = bytearray()
    i = 0
    while i < len(src) and len(out) < expected:
        n = src[i]
        i += 1
        if n < 128:
            out += src[i:i + n + 1]
            i += n + 1
        elif n > 128:
            out += src[i:i + 1] * (257 - n)
            i += 1
    return bytes(out)


def read_tiff(path: str) -> np.ndarray:
    """First image of a TIFF file as a 2-D float32 array (the reference's ``imread(addr, mode='F')``)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 8 or data[:2] not in (b"II", b"MM"):
        raise ValueError(f"{path}: not a TIFF file")
    bo = "<" if data[:2] == b"II" else ">"
    magic, ifd = struct.unpack_from(bo + "HI", data, 2)
    if magic != 42:
        raise ValueError(f"{path}: TIFF magic {magic} (BigTIFF is not supported)")
    (n_entries,) = struct.unpack_from(bo + "H", data, ifd)
    tags = {}
    for k in range(n_entries):
        tag, typ, count, raw = struct.unpack_from(bo + "HHI4s", data, ifd + 2 + 12 * k)
        if typ not in _TYPES:
            continue
        fmt, size = _TYPES[typ]
        nbytes = size * count
        buf = raw[:nbytes] if nbytes <= 4 else data[struct.unpack(bo + "I", raw)[0]:][:nbytes]
        if typ == 2:
            tags[tag] = buf
        elif typ == 5:
            v = struct.unpack(bo + "II" * count, buf)
            tags[tag] = tuple(v[2 * i] / max(v[2 * i + 1], 1) for i in range(count))
        else:
            tags[tag] = struct.unpack(bo + fmt * count, buf)
    try:
        width, height = tags[256][0], tags[257][0]
        offsets, counts = tags[273], tags[279]
    except KeyError as e:
        raise ValueError(f"{path}: missing TIFF tag {e} (tiled TIFFs are not supported)") from None
    bits = tags.get(258, (1,))[0]
    spp = tags.get(277, (1,))[0]
    fmt = tags.get(339, (1,))[0]
    compression = tags.get(259, (1,))[0]
    planar = tags.get(284, (1,))[0]
    rows_per_strip = min(tags.get(278, (height,))[0], height)
    if (fmt, bits) not in _SAMPLES:
        raise ValueError(f"{path}: unsupported sample type (SampleFormat {fmt}, {bits} bits)")
    if compression not in (1, 32773):
        raise ValueError(f"{path}: unsupported compression {compression} (only none and PackBits)")
    if planar != 1 and spp != 1:
        raise ValueError(f"{path}: planar multi-sample TIFFs are not supported")
    dt = np.dtype(bo + _SAMPLES[(fmt, bits)])
    row_bytes = width * spp * dt.itemsize
    img = np.empty((height, width * spp), dt)
    row = 0
    for off, cnt in zip(offsets, counts):
        nrows = min(rows_per_strip, height - row)
        if nrows <= 0:
            break
        chunk = data[off:off + cnt]
        if compression == 32773:
            chunk = _unpackbits(chunk, nrows * row_bytes)
        if len(chunk) < nrows * row_bytes:
            raise ValueError(f"{path}: strip at row {row} is truncated")
        img[row:row + nrows] = np.frombuffer(chunk, dt, nrows * width * spp).reshape(nrows, width * spp)
        row += nrows
    if row != height:
        raise ValueError(f"{path}: strips cover {row} of {height} rows")
    if spp > 1:
        img = img.reshape(height, width, spp)[..., 0]
    if tags.get(262, (1,))[0] == 0 and fmt != 3:      # WhiteIsZero
        img = np.iinfo(dt).max - img
    return np.ascontiguousarray(img, np.float32)


def write_tiff(path: str, img: np.ndarray) -> None:
    """2-D array -> single-strip little-endian 32-bit float TIFF with the tag set of DM3stoTIFs.m:23-34."""
    a = np.ascontiguousarray(img, "<f4")
    if a.ndim != 2:
        raise ValueError(f"expected a 2-D image, got shape {a.shape}")
    h, w = a.shape
    nbytes = a.nbytes
    if 8 + nbytes + 200 >= 1 << 32:
        raise ValueError("image too large for a classic TIFF")
    entries = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 32), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8), (277, 3, 1, 1),
               (278, 4, 1, h), (279, 4, 1, nbytes), (284, 3, 1, 1), (339, 3, 1, 3)]
    ifd = struct.pack("<H", len(entries))
    for tag, typ, count, value in entries:
        ifd += struct.pack("<HHI", tag, typ, count) + (struct.pack("<HH", value, 0) if typ == 3 else struct.pack("<I", value))
    ifd += struct.pack("<I", 0)
    pad = nbytes & 1
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8 + nbytes + pad))
        f.write(a.tobytes())
        f.write(b"\0" * pad)
        f.write(ifd)
    os.replace(tmp, path)


def load_image(addr, resizeSize=None, imgType=np.float32):
    """misc_py/denoiser-multi-gpu.py:800-814: read an image as float; a failed read yields a 512x512 image of 0.5 (and the
    reference's message); optional area resize (needs OpenCV, like the reference)."""
    try:
        img = read_tiff(addr)
    except Exception:
        img = 0.5 * np.ones((512, 512))
        print("Image read failed")
    if resizeSize:
        import cv2
        img = cv2.resize(img, resizeSize, interpolation=cv2.INTER_AREA)
    return img.astype(imgType)


def _prefetch(items: Sequence, fn: Callable, depth: int):
    """Yield fn(item) in order, computed ``depth`` items ahead on a worker thread; exceptions surface at the consumer."""
    q: "queue.Queue" = queue.Queue(maxsize=max(depth, 1))
    stop = threading.Event()

    def work():
        for it in items:
            if stop.is_set():
                return
            try:
                q.put((it, fn(it), None))
            except BaseException as e:  # noqa: BLE001 - handed to the consumer
                q.put((it, None, e))
                return
        q.put(None)

    t = threading.Thread(target=work, daemon=True)
    t.start()
    try:
        while True:
            got = q.get()
            if got is None:
                return
            it, val, err = got
            if err is not None:
                raise err
            yield it, val
    finally:
        stop.set()
        while t.is_alive():
            try:
                q.get_nowait()
            except queue.Empty:
                t.join(0.01)


def denoise_files(denoiser, paths: Sequence[str], out_dir: Optional[str] = None, rank: int = 0, world: int = 1, prefetch: int = 2,
                  suffix: str = "_denoised", on_result: Optional[Callable] = None, **denoise_kw):
    """Denoise the micrographs this rank owns (path k belongs to rank k % world) and write ``<out_dir>/<name><suffix>.tif``
    (32-bit float).  Reading, decoding and writing run on helper threads, ``denoiser.denoise`` on the caller's.  Returns
    [(index, output path or None)] for the owned files; ``on_result(index, path, image)`` sees each result."""
    mine = _sharding.shard_indices(len(paths), rank, world)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    wq: "queue.Queue" = queue.Queue(maxsize=max(prefetch, 1))
    werr = []

    def writer():
        while True:
            job = wq.get()
            if job is None:
                return
            try:
                write_tiff(*job)
            except BaseException as e:  # noqa: BLE001
                werr.append(e)

    wt = threading.Thread(target=writer, daemon=True)
    wt.start()
    done = []
    try:
        for k, img in _prefetch(mine, lambda k: read_tiff(paths[k]), prefetch):
            out = denoiser.denoise(img, **denoise_kw)
            dst = None
            if out_dir:
                stem = os.path.splitext(os.path.basename(paths[k]))[0]
                dst = os.path.join(out_dir, stem + suffix + ".tif")
                wq.put((dst, out))
            if on_result:
                on_result(k, dst, out)
            done.append((k, dst))
    finally:
        wq.put(None)
        wt.join()
    if werr:
        raise werr[0]
    return done
