"""ctypes front of the CPU oracle (oracle/gs_oracle.c) and of the compiled reference kernels
(oracle/_ref/).  TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from gsasr_b200/.

Oracle status: PINNED against (a) the reference's own brute-force formula, check.py's
``torch_version`` (utils/gs_cuda_dmax/check.py:4-31, run on CPU by tests/golden/make_golden.py,
fixtures committed), and (b) the reference's CUDA kernels themselves, compiled unmodified from
/root/reference into oracle/_ref/ and run on the B200 (tests/test_gpu_parity.py,
tests/test_gpu_fullsize.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgsoracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """gcc-build the C restatement (and, when /root/reference exists, the reference kernels)."""
    src = os.path.join(HERE, "gs_oracle.c")
    stale = lambda: not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale():
        import fcntl

        with open(LIB_PATH + ".lock", "w") as lock:  # several processes may get here at once
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def build_ref(reference: str = "/root/reference") -> bool:
    """Compile the reference's own kernels into oracle/_ref/ (only where the reference tree exists)."""
    if not os.path.isdir(os.path.join(reference, "utils", "gs_cuda_dmax")):
        return have_ref()
    subprocess.run(["make", "-C", HERE, "ref", f"REFERENCE={reference}"], check=True, capture_output=True)
    return have_ref()


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in ("libgsref_dmax.so", "libgsref_nodmax.so"))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.gso_num_threads.restype = ctypes.c_int
        L.gso_set_num_threads.argtypes = [ctypes.c_int]
        L.gso_ranges.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, _i32p]
        L.gso_ranges.restype = ctypes.c_int
        L.gso_forward.argtypes = [_f32p, _f32p, _f32p, _f64p, _i32p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_float, ctypes.c_int]
        L.gso_forward.restype = None
        L.gso_backward.argtypes = [_f32p, _f32p, _f32p, _f32p, _f64p, _f64p, _f64p, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.gso_backward.restype = None
        L.gso_forward_crop.argtypes = [_f32p, _f32p, _f32p, _f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_int] + [ctypes.c_int] * 4
        L.gso_forward_crop.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t):
    return a.ctypes.data_as(t)


def num_threads() -> int:
    return int(lib().gso_num_threads())


def set_num_threads(n: int) -> None:
    lib().gso_set_num_threads(int(n))


def ranges(coords, h: int, w: int, dmax: float) -> np.ndarray:
    """(s,4) int32 inclusive [x0,x1,y0,y1] of the reference's dmax inclusion set (gs.cu:40-50)."""
    coords = _f32(coords)
    s = coords.shape[0]
    out = np.zeros((max(s, 1), 4), dtype=np.int32)
    bad = lib().gso_ranges(_p(coords, _f32p), s, h, w, float(dmax), _p(out, _i32p))
    assert bad == 0, "inclusion set not contiguous"
    return out[:s]


def forward(sigmas, coords, colors, h: int, w: int, dmax: float = float("inf"), mode: int = 0,
            init=None, with_count: bool = False):
    """Reference forward (gs.cu:24-62) on the CPU.  Returns float64 (h,w,3) [and int32 (h,w) counts]."""
    sigmas, coords, colors = _f32(sigmas), _f32(coords), _f32(colors)
    s = sigmas.shape[0]
    img = np.zeros((h, w, 3), dtype=np.float64) if init is None else np.array(init, dtype=np.float64)
    cnt = np.zeros((h, w), dtype=np.int32) if with_count else None
    lib().gso_forward(_p(sigmas, _f32p), _p(coords, _f32p), _p(colors, _f32p), _p(img, _f64p),
                      _p(cnt, _i32p) if with_count else None, s, h, w, float(dmax), int(mode))
    return (img, cnt) if with_count else img


def forward_crop(sigmas, coords, colors, h: int, w: int, dmax: float, y0: int, x0: int, ch: int, cw: int,
                 mode: int = 0, rr=None) -> np.ndarray:
    """Rows [y0, y0+ch) x columns [x0, x0+cw) of forward(...): float64 (ch,cw,3).  Exact -- every Gaussian
    whose dmax window reaches the rectangle is summed -- but only the rectangle's pixels are evaluated, so
    the full-size configs (4096^2) are within the CPU's reach."""
    sigmas, coords, colors = _f32(sigmas), _f32(coords), _f32(colors)
    assert 0 <= y0 and y0 + ch <= h and 0 <= x0 and x0 + cw <= w
    if rr is None:  # (s,4) inclusion ranges; pass them in when several crops share the field
        rr = ranges(coords, h, w, dmax)
    sel = np.nonzero((rr[:, 0] < x0 + cw) & (rr[:, 1] >= x0) & (rr[:, 2] < y0 + ch) & (rr[:, 3] >= y0))[0]
    sg, xy, col = (np.ascontiguousarray(a[sel]) for a in (sigmas, coords, colors))
    crop = np.zeros((ch, cw, 3), dtype=np.float64)
    lib().gso_forward_crop(_p(sg, _f32p), _p(xy, _f32p), _p(col, _f32p), _p(crop, _f64p), sg.shape[0], h, w,
                           float(dmax), int(mode), int(y0), int(x0), int(ch), int(cw))
    return crop


def backward(sigmas, coords, colors, grads, dmax: float = float("inf")):
    """Reference backward (gs.cu:97-162) on the CPU.  Returns float64 (g_sigmas, g_coords, g_colors)."""
    sigmas, coords, colors, grads = _f32(sigmas), _f32(coords), _f32(colors), _f32(grads)
    s = sigmas.shape[0]
    h, w, c = grads.shape
    assert c == 3
    gs = np.zeros((max(s, 1), 3), dtype=np.float64)
    gc = np.zeros((max(s, 1), 2), dtype=np.float64)
    gk = np.zeros((max(s, 1), 3), dtype=np.float64)
    lib().gso_backward(_p(sigmas, _f32p), _p(coords, _f32p), _p(colors, _f32p), _p(grads, _f32p),
                       _p(gs, _f64p), _p(gc, _f64p), _p(gk, _f64p), s, h, w, float(dmax))
    return gs[:s], gc[:s], gk[:s]


def brute_force_numpy(sigmas, coords, colors, h: int, w: int, dmax: float = 100.0) -> np.ndarray:
    """Pure-numpy restatement of check.py's torch_version (utils/gs_cuda_dmax/check.py:13-29):
    per-pixel closed form, `<= dmax` mask on the exact (python double) pixel coordinates.
    Small sizes only; used to cross-check the C oracle, not the product."""
    sg, xy, col = (np.asarray(a, dtype=np.float64) for a in (sigmas, coords, colors))
    img = np.zeros((h, w, col.shape[1]))
    for hi in range(h):
        for wi in range(w):
            curh = 2 * hi / (h - 1) - 1.0
            curw = 2 * wi / (w - 1) - 1.0
            dx, dy = curw - xy[:, 0], curh - xy[:, 1]
            v = dx ** 2 / sg[:, 0] ** 2
            v -= 2 * sg[:, 2] * dx * dy / sg[:, 0] / sg[:, 1]
            v += dy ** 2 / sg[:, 1] ** 2
            v *= -1.0 / (2.0 * (1 - sg[:, 2] ** 2))
            v = np.exp(v)
            m = (np.abs(dx) <= dmax) & (np.abs(dy) <= dmax)
            img[hi, wi] = (v[:, None] * col)[m].sum(0)
    return img


# ---- the reference's own CUDA kernels (GPU box only) ---------------------------------------------
class RefKernels:
    """oracle/_ref/libgsref_{dmax,nodmax}.so: the UNMODIFIED reference kernels rebuilt for sm_100a.
    Takes torch CUDA tensors; launches on the legacy default stream like the reference does."""

    def __init__(self, dmax_variant: bool = True):
        name = "libgsref_dmax.so" if dmax_variant else "libgsref_nodmax.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        self.lib.gsref_forward.argtypes = [vp, vp, vp, vp] + [ctypes.c_int] * 4 + [ctypes.c_float]
        self.lib.gsref_backward.argtypes = [vp] * 7 + [ctypes.c_int] * 4 + [ctypes.c_float]
        self.lib.gsref_sync.restype = ctypes.c_int

    def forward(self, sigmas, coords, colors, img, dmax: float = float("inf")):
        import torch

        torch.cuda.current_stream().synchronize()
        h, w, c = img.shape
        rc = self.lib.gsref_forward(sigmas.data_ptr(), coords.data_ptr(), colors.data_ptr(),
                                    img.data_ptr(), sigmas.shape[0], h, w, c, float(dmax))
        assert rc == 0, f"reference forward launch failed: cuda error {rc}"
        assert self.lib.gsref_sync() == 0
        return img

    def backward(self, sigmas, coords, colors, grads, gs, gc, gk, dmax: float = float("inf")):
        import torch

        torch.cuda.current_stream().synchronize()
        h, w, c = grads.shape
        rc = self.lib.gsref_backward(sigmas.data_ptr(), coords.data_ptr(), colors.data_ptr(),
                                     grads.data_ptr(), gs.data_ptr(), gc.data_ptr(), gk.data_ptr(),
                                     sigmas.shape[0], h, w, c, float(dmax))
        assert rc == 0, f"reference backward launch failed: cuda error {rc}"
        assert self.lib.gsref_sync() == 0
        return gs, gc, gk
