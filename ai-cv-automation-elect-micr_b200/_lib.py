"""ctypes binding of the C ABI declared in include/emd.h (libemd.so, built by csrc/Makefile).

There is no CPU fallback: if the shared library is missing the import of the
engine fails loudly, and if no CUDA device is usable ``emd_create`` returns
EMD_ECUDA and the Python side raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EMD_LIB") or os.path.join(_HERE, "libemd.so")   # EMD_LIB: A/B builds of the same ABI

EMD_MODE_FP32, EMD_MODE_BF16, EMD_MODE_FP16 = 0, 1, 2
EMD_VARIANT_A, EMD_VARIANT_B = 0, 1
EMD_FLAG_PREPROCESS, EMD_FLAG_POSTPROCESS, EMD_FLAG_INPUT_F64, EMD_FLAG_OUTPUT_F32 = 1, 2, 4, 8

MODES = {"fp32": EMD_MODE_FP32, "bf16": EMD_MODE_BF16, "fp16": EMD_MODE_FP16}

# name -> (restype, argtypes); mirrors include/emd.h one to one
_P, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
_IP = C.POINTER(C.c_int)
SIGNATURES = {
    "emd_version": (_I, []),
    "emd_last_error": (C.c_char_p, [_P]),
    "emd_create": (_I, [C.POINTER(_P), _I, _I, _I, _I]),
    "emd_destroy": (_I, [_P]),
    "emd_load_weights": (_I, [_P, _P, _SZ]),
    "emd_forward": (_I, [_P, _P, _I, _P, _I, _P]),
    "emd_forward_async": (_I, [_P, _P, _I, _P, _I, _P]),
    "emd_synchronize": (_I, [_P, _P]),
    "emd_plan_tiles": (_I, [_I, _I, _I, _I, _IP, _IP, _IP, _IP]),
    "emd_normalise": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "emd_preprocess_crop": (_I, [_P, _P, _I, _I, _P, _P]),
    "emd_gather_crops": (_I, [_P, _P, _I, _I, _IP, _IP, _I, _I, _I, _P, _P]),
    "emd_stitch": (_I, [_P, _P, _IP, _IP, _I, _I, _I, _I, _I, _I, _P, _P]),
    "emd_denoise_image": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "emd_denoise_stream": (_I, [_P, C.POINTER(_P), _I, _I, _I, _I, _I, _I, C.POINTER(_P), _P]),
    "emd_quality": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "emd_set_keep_activations": (_I, [_P, _I]),
    "emd_get_activation": (_I, [_P, C.c_char_p, _P, _SZ, _IP]),
    "emd_run_layer": (_I, [_P, C.c_char_p, _P, _P, _I, _P, _SZ, _I, _IP]),
    "emd_kernel_launches": (C.c_longlong, [_P]),
    "emd_tensor_core_launches": (C.c_longlong, [_P]),
    "emd_graph_replays": (C.c_longlong, [_P]),
    "emd_counter": (C.c_longlong, [_P, C.c_char_p]),
    "emd_set_option": (_I, [_P, C.c_char_p, C.c_longlong]),
    "emd_get_option": (C.c_longlong, [C.c_char_p]),
    "emd_set_tensor_cores": (_I, [_P, _I]),
    "emd_set_profile": (_I, [_P, _I]),
    "emd_num_steps": (_I, [_P]),
    "emd_step_launches": (_I, [_P, _I]),
    "emd_step_info": (_I, [_P, _I, C.c_char_p, _SZ, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libemd.so and declare every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or __graft_entry__.build()).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
