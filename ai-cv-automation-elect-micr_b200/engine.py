"""Thin Python handle over the C ABI (include/emd.h).  Tensor plumbing only: numpy arrays or
torch tensors go in as raw pointers; all arithmetic happens in libemd.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODES


def _stream_for(x, stream):
    """CUDA stream to launch on: the caller's, else torch's current stream when the buffer is a torch CUDA tensor (so the
    call is ordered after the work that produced it), else None = the engine's own stream."""
    if stream:
        return C.c_void_p(stream)
    if hasattr(x, "is_cuda") and x.is_cuda:
        import torch
        return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    return None


def _ptr(x):
    """Raw address of a numpy array or a torch tensor (host or CUDA)."""
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(f"unsupported buffer type {type(x)}")


def _dtype_name(x):
    return str(x.dtype).replace("torch.", "")


def _dense(x, what, dtypes=("float32",), shape=None):
    """The C ABI reads raw, dense memory: numpy inputs are converted (a copy only when needed), torch tensors must
    already be contiguous and of an accepted dtype (a silent copy could land on the wrong device or stream)."""
    if isinstance(x, np.ndarray):
        if _dtype_name(x) not in dtypes:
            x = x.astype(dtypes[0])
        x = np.ascontiguousarray(x)
    elif hasattr(x, "data_ptr"):
        if _dtype_name(x) not in dtypes:
            raise TypeError(f"{what}: torch dtype {x.dtype} (accepted: {dtypes})")
        if not x.is_contiguous():
            raise TypeError(f"{what}: torch tensor must be contiguous")
    else:
        raise TypeError(f"{what}: unsupported buffer type {type(x)}")
    if shape is not None and tuple(x.shape) != tuple(shape):
        raise ValueError(f"{what}: shape {tuple(x.shape)}, expected {tuple(shape)}")
    return x


def _check_out(out, what, dtype, shape):
    """A caller-supplied output buffer is written in place: it must already be exactly what the C side assumes."""
    if _dtype_name(out) != dtype:
        raise TypeError(f"{what}: dtype {out.dtype}, expected {dtype}")
    if tuple(out.shape) != tuple(shape):
        raise ValueError(f"{what}: shape {tuple(out.shape)}, expected {tuple(shape)}")
    contiguous = out.flags["C_CONTIGUOUS"] if isinstance(out, np.ndarray) else out.is_contiguous()
    if not contiguous:
        raise TypeError(f"{what}: must be contiguous")
    return out


class Engine:
    """One GPU + weights + workspace (an ``emd_engine``).  Not thread-safe (emd.h)."""

    def __init__(self, device: int = 0, cropsize: int = 512, max_batch: int = 8, variant: str = "A"):
        self.lib = _lib.load()
        self.S = cropsize
        self.device = device
        self.max_batch = max_batch
        h = C.c_void_p()
        rc = self.lib.emd_create(C.byref(h), device, cropsize, 0 if variant == "A" else 1, max_batch)
        if rc != 0:
            raise RuntimeError(f"emd_create failed ({rc}): {self.lib.emd_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.emd_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.emd_last_error(self.h).decode()}")

    def load_weights(self, blob: bytes):
        self._blob = blob
        self._check(self.lib.emd_load_weights(self.h, C.c_char_p(blob), len(blob)), "emd_load_weights")

    # -- network ---------------------------------------------------------------------------
    def forward(self, crops, out=None, mode="fp16", stream=None, sync=True):
        """crops [n,S,S] float32 (numpy, or torch host/CUDA tensor) -> [n,S,S] float32.
        sync=False (host buffers, a stream of batches): return without waiting (emd_forward_async); successive calls pipeline
        their copies under each other's network passes; call ``synchronize()`` before touching the buffers, and give every
        outstanding call its own ``out``."""
        n = int(crops.shape[0])
        if len(crops.shape) != 3 or tuple(crops.shape[1:3]) != (self.S, self.S):
            raise ValueError(f"crops must be [n,{self.S},{self.S}], got {tuple(crops.shape)}")
        crops = _dense(crops, "crops")
        if out is not None:
            _check_out(out, "out", "float32", (n, self.S, self.S))
        if out is None:
            if isinstance(crops, np.ndarray):
                out = np.empty((n, self.S, self.S), np.float32)
            else:
                import torch
                out = torch.empty((n, self.S, self.S), dtype=torch.float32, device=crops.device)
        fn = self.lib.emd_forward if sync else self.lib.emd_forward_async
        self._check(fn(self.h, _ptr(crops), n, _ptr(out), MODES[mode], _stream_for(crops, stream)), "emd_forward")
        if not sync:
            self._inflight = getattr(self, "_inflight", []) + [(crops, out)]       # keep the buffers alive until synchronize()
        return out

    def synchronize(self, stream=None):
        """Wait for every outstanding ``forward(..., sync=False)`` call (emd_synchronize)."""
        self._check(self.lib.emd_synchronize(self.h, C.c_void_p(stream) if stream else None), "emd_synchronize")
        self._inflight = []

    def run_layer(self, name, x, res=None, mode="fp32"):
        """One fused layer on NHWC float32 host inputs (parity hook)."""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        dims = (C.c_int * 4)()
        r = None if res is None else np.ascontiguousarray(res, np.float32)
        self._check(self.lib.emd_run_layer(self.h, name.encode(), _ptr(x), None if r is None else _ptr(r), n,
                                           None, 0, MODES[mode], dims), "emd_run_layer(size)")
        out = np.empty(tuple(dims), np.float32)
        self._check(self.lib.emd_run_layer(self.h, name.encode(), _ptr(x), None if r is None else _ptr(r), n,
                                           _ptr(out), out.size, MODES[mode], dims), f"emd_run_layer({name})")
        return out

    def set_keep_activations(self, keep=True):
        self._check(self.lib.emd_set_keep_activations(self.h, int(keep)), "emd_set_keep_activations")

    def activation(self, name):
        dims = (C.c_int * 4)()
        self._check(self.lib.emd_get_activation(self.h, name.encode(), None, 0, dims), "emd_get_activation(size)")
        out = np.empty(tuple(dims), np.float32)
        self._check(self.lib.emd_get_activation(self.h, name.encode(), _ptr(out), out.size, dims),
                    f"emd_get_activation({name})")
        return out

    # -- wrapper pieces ----------------------------------------------------------------------
    def plan_tiles(self, H, W, crop=None, overlap=80):
        crop = crop or self.S
        if crop <= 0 or not 0 <= overlap < crop or H < crop or W < crop:
            raise ValueError(f"cannot tile a {H}x{W} image with crop {crop}, overlap {overlap}")
        cap = max(H, W) // (crop - overlap) + 2
        ys, xs = (C.c_int * cap)(), (C.c_int * cap)()
        ny, nx = C.c_int(), C.c_int()
        rc = self.lib.emd_plan_tiles(H, W, crop, overlap, ys, xs, C.byref(ny), C.byref(nx))
        if rc != 0:
            raise ValueError(f"emd_plan_tiles({H},{W},{crop},{overlap}) failed ({rc})")
        return list(ys[: ny.value]), list(xs[: nx.value])

    def normalise(self, img):
        img = np.ascontiguousarray(img)
        if img.dtype not in (np.float32, np.float64):
            img = img.astype(np.float32)
        out = np.empty(img.shape, np.float32)
        self._check(self.lib.emd_normalise(self.h, _ptr(img), int(img.dtype == np.float64), img.shape[0],
                                           img.shape[1], _ptr(out), None), "emd_normalise")
        return out

    def preprocess_crop(self, img, out=None):
        """Denoiser.preprocess (DEN:632-643) on the GPU: img [H,W] float32 (numpy / torch host or CUDA) -> [S,S] float32."""
        if len(img.shape) != 2:
            raise ValueError(f"preprocess_crop expects a 2-D image, got shape {tuple(img.shape)}")
        img = _dense(img, "img")
        if out is None:
            if isinstance(img, np.ndarray):
                out = np.empty((self.S, self.S), np.float32)
            else:
                import torch
                out = torch.empty((self.S, self.S), dtype=torch.float32, device=img.device)
        else:
            _check_out(out, "out", "float32", (self.S, self.S))
        self._check(self.lib.emd_preprocess_crop(self.h, _ptr(img), int(img.shape[0]), int(img.shape[1]), _ptr(out),
                                                 _stream_for(img, None)), "emd_preprocess_crop")
        return out

    def gather_crops(self, img, ys, xs, crop=None):
        crop = crop or self.S
        img = np.ascontiguousarray(img, np.float32)
        out = np.empty((len(ys) * len(xs), crop, crop), np.float32)
        ya, xa = (C.c_int * len(ys))(*ys), (C.c_int * len(xs))(*xs)
        self._check(self.lib.emd_gather_crops(self.h, _ptr(img), img.shape[0], img.shape[1], ya, xa, len(ys),
                                              len(xs), crop, _ptr(out), None), "emd_gather_crops")
        return out

    def stitch(self, tiles, ys, xs, H, W, crop=None, clip=True):
        crop = crop or self.S
        tiles = np.ascontiguousarray(tiles, np.float32)
        out = np.empty((H, W), np.float64)
        ya, xa = (C.c_int * len(ys))(*ys), (C.c_int * len(xs))(*xs)
        self._check(self.lib.emd_stitch(self.h, _ptr(tiles), ya, xa, len(ys), len(xs), crop, H, W, int(clip),
                                        _ptr(out), None), "emd_stitch")
        return out

    def quality(self, a, b):
        """MSE, the trainer's Huberised loss and mean SSIM (tf_ssim) between image pairs: a, b [n,H,W] (or [H,W]) float32,
        numpy or torch (host / CUDA) -> float64 array [n,3] (emd_quality)."""
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, np.float32)
        if isinstance(b, np.ndarray):
            b = np.ascontiguousarray(b, np.float32)
        if tuple(a.shape) != tuple(b.shape) or len(a.shape) not in (2, 3):
            raise ValueError(f"quality: shapes {tuple(a.shape)} and {tuple(b.shape)}")
        a, b = _dense(a, "quality: a"), _dense(b, "quality: b")
        n = 1 if len(a.shape) == 2 else int(a.shape[0])
        H, W = int(a.shape[-2]), int(a.shape[-1])
        out = np.empty((n, 3), np.float64)
        self._check(self.lib.emd_quality(self.h, _ptr(a), _ptr(b), n, H, W, _ptr(out), _stream_for(a, None)), "emd_quality")
        return out

    def _image_args(self, img, preprocess):
        if len(img.shape) != 2:
            raise ValueError(f"expected a 2-D micrograph, got shape {tuple(img.shape)}")
        if isinstance(img, np.ndarray) and img.dtype == np.float64 and not preprocess:
            img = img.astype(np.float32)
        return _dense(img, "img", ("float32", "float64") if preprocess else ("float32",))

    def _image_out(self, img, out, H, W, out_dtype):
        name = np.dtype(out_dtype).name
        if name not in ("float32", "float64"):
            raise TypeError(f"out_dtype {out_dtype}: float64 (the reference's accumulator type) or float32")
        if out is not None:
            return _check_out(out, "out", name, (H, W))
        if isinstance(img, np.ndarray):
            return np.empty((H, W), out_dtype)
        import torch
        return torch.empty((H, W), dtype=getattr(torch, name), device=img.device)

    def _image_flags(self, preprocess, postprocess, f64, out):
        return (_lib.EMD_FLAG_PREPROCESS if preprocess else 0) | (_lib.EMD_FLAG_POSTPROCESS if postprocess else 0) \
            | (_lib.EMD_FLAG_INPUT_F64 if f64 else 0) | (_lib.EMD_FLAG_OUTPUT_F32 if _dtype_name(out) == "float32" else 0)

    def denoise_image(self, img, overlap=80, preprocess=True, postprocess=True, mode="fp16", out=None, out_dtype=np.float64):
        """Whole micrograph in one call: normalise -> tile -> batched forward -> stitch (emd_denoise_image).
        out_dtype: float64 like the reference's accumulators (DEN:658), or float32 (the same values rounded once; half the
        download)."""
        img = self._image_args(img, preprocess)
        H, W = int(img.shape[0]), int(img.shape[1])
        if H < self.S or W < self.S:
            raise ValueError(f"image {H}x{W} is smaller than the {self.S}x{self.S} crop")
        if not 0 <= overlap < self.S:
            raise ValueError(f"overlap {overlap} outside [0,{self.S})")
        out = self._image_out(img, out, H, W, out_dtype)
        flags = self._image_flags(preprocess, postprocess, _dtype_name(img) == "float64", out)
        self._check(self.lib.emd_denoise_image(self.h, _ptr(img), H, W, overlap, flags, MODES[mode], _ptr(out),
                                               _stream_for(img, None)), "emd_denoise_image")
        return out

    def denoise_images(self, imgs, overlap=80, preprocess=True, postprocess=True, mode="fp16", outs=None, out_dtype=np.float64,
                       stream=None):
        """A stream of same-sized micrographs (emd_denoise_stream): bit-identical to denoise_image on each, with the
        next image's upload / normalise / tile gather and the previous image's stitch / download under the current
        image's network pass.  Pinned host buffers (torch ``pin_memory``) are what makes the copies overlap."""
        imgs = [self._image_args(im, preprocess) for im in imgs]
        if not imgs:
            return []
        H, W = int(imgs[0].shape[0]), int(imgs[0].shape[1])
        kind = _dtype_name(imgs[0])
        for im in imgs:
            if tuple(im.shape) != (H, W) or _dtype_name(im) != kind:
                raise ValueError("denoise_images: all micrographs of one call must have the same shape and dtype")
        if H < self.S or W < self.S:
            raise ValueError(f"image {H}x{W} is smaller than the {self.S}x{self.S} crop")
        if not 0 <= overlap < self.S:
            raise ValueError(f"overlap {overlap} outside [0,{self.S})")
        if outs is None:
            outs = [None] * len(imgs)
        if len(outs) != len(imgs):
            raise ValueError("denoise_images: one output buffer per image")
        outs = [self._image_out(im, o, H, W, out_dtype) for im, o in zip(imgs, outs)]
        flags = self._image_flags(preprocess, postprocess, kind == "float64", outs[0])
        n = len(imgs)
        ins_a = (C.c_void_p * n)(*[_ptr(im) for im in imgs])
        outs_a = (C.c_void_p * n)(*[_ptr(o) for o in outs])
        self._check(self.lib.emd_denoise_stream(self.h, ins_a, n, H, W, overlap, flags, MODES[mode], outs_a,
                                                _stream_for(imgs[0], stream)), "emd_denoise_stream")
        return outs

    # -- measurement ---------------------------------------------------------------------------
    @property
    def kernel_launches(self):
        return int(self.lib.emd_kernel_launches(self.h))

    @property
    def tensor_core_launches(self):
        return int(self.lib.emd_tensor_core_launches(self.h))

    @property
    def graph_replays(self):
        return int(self.lib.emd_graph_replays(self.h))

    def counter(self, name):
        """Launch counters by name (emd_counter): e.g. 'conv_fused_pair', 'conv_fused_dw', 'conv_cuda_core'."""
        v = int(self.lib.emd_counter(self.h, name.encode()))
        if v < 0:
            raise KeyError(name)
        return v

    def set_option(self, name, value):
        """Tuning / A-B switches (emd_set_option; process-wide, see include/emd.h)."""
        self._check(self.lib.emd_set_option(self.h, name.encode(), int(value)), f"emd_set_option({name})")

    def get_option(self, name):
        return int(self.lib.emd_get_option(name.encode()))

    def set_tensor_cores(self, on=True):
        self._check(self.lib.emd_set_tensor_cores(self.h, int(on)), "emd_set_tensor_cores")

    def set_profile(self, on=True):
        self._check(self.lib.emd_set_profile(self.h, int(on)), "emd_set_profile")

    def step_info(self):
        """[(name, ms, flops_per_crop, bytes_per_crop, launches)] of the last profiled forward."""
        out = []
        name = C.create_string_buffer(64)
        ms, fl, by = C.c_float(), C.c_double(), C.c_double()
        for i in range(self.lib.emd_num_steps(self.h)):
            self.lib.emd_step_info(self.h, i, name, 64, C.byref(ms), C.byref(fl), C.byref(by))
            out.append((name.value.decode(), ms.value, fl.value, by.value, int(self.lib.emd_step_launches(self.h, i))))
        return out
