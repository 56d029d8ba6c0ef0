"""ctypes binding of libgsraster.so (include/gsraster.h).  No CPU fallback: if the library is
missing or a call fails, this raises -- the CUDA path is the only path."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSR_LIB_PATH") or os.path.join(HERE, "libgsraster.so")

GSR_FLAG_OVERWRITE = 0x1
GSR_FLAG_CHW = 0x2
GSR_FLAG_U8 = 0x4
GSR_FLAG_BGR = 0x8
GSR_FLAG_ROW_STORES = 0x10
GSR_FLAG_DETERMINISTIC = 0x20
DEFAULT_KSIGMA = 5.0
EXACT_KSIGMA = float("inf")

_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_i = ctypes.c_int
_f = ctypes.c_float
_u32 = ctypes.c_uint32


class GsrSample(ctypes.Structure):
    """struct gsr_sample (include/gsraster.h)."""

    _fields_ = [
        ("sigmas", _vp), ("coords", _vp), ("colors", _vp), ("img", _vp), ("grads", _vp),
        ("grads_sigmas", _vp), ("grads_coords", _vp), ("grads_colors", _vp),
        ("s", _i), ("h", _i), ("w", _i), ("dmax", _f),
    ]


GSR_MAX_CLIP = 8


class GsrWindow(ctypes.Structure):
    """struct gsr_window (include/gsraster.h)."""

    _fields_ = [("row_stride", ctypes.c_longlong), ("pix_stride", ctypes.c_longlong),
                ("chan_stride", ctypes.c_longlong), ("nclip", _i), ("clip", (_i * 4) * GSR_MAX_CLIP)]


# name -> (restype, argtypes): every symbol include/gsraster.h declares
SIGNATURES = {
    "gsr_version": (_i, []),
    "gsr_status_string": (ctypes.c_char_p, [_i]),
    "gsr_last_cuda_error": (_i, []),
    "gsr_workspace_bytes": (_sz, [_i, _i, _i]),
    "gsr_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_backward": (_i, [_vp] * 7 + [_i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_forward_band": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_backward_band": (_i, [_vp] * 7 + [_i, _i, _i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_forward_window": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(GsrWindow), _i, _i, _i, _i, _f, _f, _u32, _vp,
                                _sz, _vp]),
    "gsr_frontend_forward_window": (_i, [_vp, _vp, _vp, ctypes.POINTER(GsrWindow), _i, _i, _i, _f, _f, _f, _u32, _vp,
                                         _sz, _vp]),
    "gsr_prepare": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _sz, _vp]),
    "gsr_forward_prepared": (_i, [_vp, _i, _i, _i, _f, _u32, _vp, _sz, _vp]),
    "gsr_backward_prepared": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _u32, _vp, _sz, _vp]),
    "gsr_workspace_bytes_batch": (_sz, [ctypes.POINTER(GsrSample), _i]),
    "gsr_forward_batch": (_i, [ctypes.POINTER(GsrSample), _i, _f, _u32, _vp, _sz, _vp]),
    "gsr_backward_batch": (_i, [ctypes.POINTER(GsrSample), _i, _f, _u32, _vp, _sz, _vp]),
    "gsr_workspace_bytes_batch_uniform": (_sz, [_i, _i, _i, _i]),
    "gsr_forward_batch_uniform": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_backward_batch_uniform": (_i, [_vp] * 7 + [_i, _i, _i, _i, _i, _f, _f, _u32, _vp, _sz, _vp]),
    "gsr_frontend_forward_batch_uniform": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "gsr_frontend_backward_batch_uniform": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "gsr_workspace_bytes_batch_padded": (_sz, [_i, _i, _i, _i]),
    "gsr_forward_batch_padded": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_f), _f, _f,
                                      _u32, _vp, _sz, _vp]),
    "gsr_backward_batch_padded": (_i, [_vp] * 7 + [_i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_f), _f, _f, _u32,
                                                  _vp, _sz, _vp]),
    "gsr_frontend_forward_batch_padded": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_f),
                                               ctypes.POINTER(_f), _f, _f, _vp, _sz, _vp]),
    "gsr_frontend_backward_batch_padded": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(_i),
                                                ctypes.POINTER(_f), ctypes.POINTER(_f), _f, _f, _vp, _sz, _vp]),
    "gsr_frontend_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "gsr_frontend_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "gsr_head_tail_forward": (_i, [_vp] * 8 + [_i, _i, _i, _vp]),
    "gsr_l1_crop_workspace_bytes": (_sz, []),
    "gsr_l1_crop_loss": (_i, [_vp, ctypes.POINTER(ctypes.c_longlong), _vp, ctypes.POINTER(ctypes.c_longlong), _vp, _vp,
                              _i, _i, _i, ctypes.POINTER(_i), _f, _i, _vp, _sz, _vp]),
}

# CPU test hooks (include/gsraster_test.h)
TEST_SIGNATURES = {
    "gsr_host_setup": (None, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "gsr_host_setup_band": (None, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp]),
    "gsr_host_window_range": (None, [_i, _f, _f, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "gsr_host_region_mask": (ctypes.c_uint, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _i]),
    "gsr_host_entries": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _i]),
    "gsr_host_geometry": (None, [ctypes.POINTER(_i)] * 5),
    "gsr_test_umma_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
}

_lib = None


class GsrError(RuntimeError):
    pass


def load():
    """Load libgsraster.so (building it is build.py's job).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GsrError(
                f"{LIB_PATH} not found: build it with `python -m gsasr_b200.build` "
                "(there is no CPU fallback for the rasteriser)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **TEST_SIGNATURES}.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        L = load()
        msg = L.gsr_status_string(status).decode()
        if status == 6:
            msg += f" (cudaError {L.gsr_last_cuda_error()})"
        raise GsrError(f"libgsraster: {msg}")
