"""The C-ABI library: loads, exports every declared symbol, validates arguments (no GPU work)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from gsasr_b200 import _lib


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared("gsraster.h") + _declared("gsraster_test.h")
    assert "gsr_forward" in names and "gsr_backward" in names and len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by libgsraster.so"
    # and the ctypes table covers the whole public header
    assert set(_declared("gsraster.h")) == set(_lib.SIGNATURES)
    assert set(_declared("gsraster_test.h")) == set(_lib.TEST_SIGNATURES)


def test_no_cpu_fallback_symbols():
    """The product library must not link the oracle."""
    out = os.popen(f"nm -D {_lib.LIB_PATH}").read()
    assert "gso_" not in out


def test_version_and_strings():
    L = _lib.load()
    assert L.gsr_version() == 100
    assert L.gsr_status_string(0) == b"ok"
    for code in range(1, 7):
        assert len(L.gsr_status_string(code)) > 3


def test_workspace_bytes():
    L = _lib.load()
    a = L.gsr_workspace_bytes(1000, 64, 64)
    b = L.gsr_workspace_bytes(2000, 64, 64)
    c = L.gsr_workspace_bytes(1000, 1024, 1024)
    assert 0 < a < b and a < c
    assert L.gsr_workspace_bytes(0, 64, 64) > 0
    for bad in ((-1, 64, 64), (10, 1, 64), (10, 64, 1), (10, 40000, 64), (10, 64, 40000)):
        assert L.gsr_workspace_bytes(*bad) == 0
    # bounded by sizes alone: ~250 B per Gaussian (records, boxes, 28 bucket slots, the backward's moment row)
    # + ~140 B per 16x8 region
    assert L.gsr_workspace_bytes(2097152, 2048, 4096) < 520 * 2**20


def test_argument_validation_without_gpu():
    L = _lib.load()
    fake = ctypes.c_void_p(256)  # never dereferenced: validation fails first
    ws = ctypes.c_void_p(4096)
    f = L.gsr_forward
    assert f(fake, fake, fake, fake, 10, 64, 64, 4, 0.1, 0.0, 0, ws, 1 << 30, None) == 3      # c != 3
    assert f(fake, fake, fake, fake, 10, 1, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 2       # h < 2
    assert f(fake, fake, fake, fake, -1, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 2      # s < 0
    assert f(fake, fake, fake, None, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 1      # img NULL
    assert f(None, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 1      # sigmas NULL
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, 0, None, 1 << 30, None) == 4    # no workspace
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 16, None) == 4           # too small
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, 0, ctypes.c_void_p(4100), 1 << 30, None) == 4  # misaligned
    b = L.gsr_backward
    assert b(fake, fake, fake, fake, fake, fake, fake, 10, 64, 64, 2, 0.1, 0.0, 0, ws, 1 << 30, None) == 3
    assert b(fake, fake, fake, None, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 1
    assert L.gsr_frontend_forward(fake, fake, fake, 10, 64, 64, 0.0, 0.1, 0.0, ws, 1 << 30, None) == 5  # step<=0
    assert L.gsr_forward_batch(None, 2, 0.0, 0, ws, 1 << 30, None) == 1
    assert L.gsr_forward_batch(None, -1, 0.0, 0, ws, 1 << 30, None) == 5
    assert L.gsr_forward_batch(None, 0, 0.0, 0, None, 0, None) == 0


def test_batch_workspace_is_sum_of_samples():
    L = _lib.load()
    arr = (_lib.GsrSample * 3)()
    dims = [(100, 48, 48), (5000, 190, 130), (0, 64, 64)]
    for a, (s, h, w) in zip(arr, dims):
        a.s, a.h, a.w, a.dmax = s, h, w, 0.5
    assert L.gsr_workspace_bytes_batch(arr, 3) == sum(L.gsr_workspace_bytes(*d) for d in dims)
    arr[1].h = 1
    assert L.gsr_workspace_bytes_batch(arr, 3) == 0


def test_python_wrapper_rejects_cpu_tensors_like_the_reference():
    """CHECK_CUDA / CHECK_CONTIGUOUS wording of gswrapper.cpp:5-7."""
    import torch

    from gsasr_b200 import gscuda

    s = torch.rand(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        gscuda.gs_render(s, torch.rand(4, 2), torch.rand(4, 3), torch.zeros(8, 8, 3), 4, 8, 8, 3, 0.5)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        gscuda.gs_render_backward(s, torch.rand(4, 2), torch.rand(4, 3), torch.zeros(8, 8, 3), s, s, s, 4, 8, 8, 3, 0.5)


def test_top_level_gscuda_module_has_reference_surface():
    import gscuda  # what utils/gs_cuda_dmax/gswrapper.py:19 imports

    assert callable(gscuda.gs_render) and callable(gscuda.gs_render_backward)


def test_padded_batch_with_no_gaussians_is_validated_not_divided():
    """s_per == 0 used to reach `s / s_per` (SIGFPE): the sample count now travels explicitly.  Without a
    GPU the call must come back with an error CODE (the first CUDA call fails), not kill the process."""
    L = _lib.load()
    fake, ws = ctypes.c_void_p(256), ctypes.c_void_p(4096)
    hw = (ctypes.c_int * 4)(64, 64, 40, 48)
    need = L.gsr_workspace_bytes_batch_padded(2, 0, 64, 64)
    assert need > 0
    rc = L.gsr_forward_batch_padded(None, None, None, fake, 2, 0, 64, 64, hw, None, 0.1, 0.0, 0, ws, need, None)
    assert rc in (0, 6)   # GSR_OK on a GPU box, GSR_ERR_CUDA here
    rc = L.gsr_backward_batch_padded(None, None, None, fake, None, None, None, 2, 0, 64, 64, hw, None, 0.1, 0.0, 0,
                                     ws, need, None)
    assert rc in (0, 6)


def test_argument_validation_of_the_band_batch_window_entry_points():
    """No GPU work: every call fails validation before anything is launched."""
    L = _lib.load()
    fake, ws = ctypes.c_void_p(256), ctypes.c_void_p(4096)
    fb = L.gsr_forward_band
    assert fb(fake, fake, fake, fake, 10, 64, 64, 3, 60, 8, 0.1, 0.0, 0, ws, 1 << 30, None) == 2   # band leaves the image
    assert fb(fake, fake, fake, fake, 10, 64, 64, 3, 0, 1, 0.1, 0.0, 0, ws, 1 << 30, None) == 2    # rows < 2
    assert fb(fake, fake, fake, fake, 10, 40000, 64, 3, 0, 8, 0.1, 0.0, 0, ws, 1 << 30, None) == 2  # h too large
    assert fb(fake, fake, fake, None, 10, 64, 64, 3, 8, 8, 0.1, 0.0, 0, ws, 1 << 30, None) == 1
    assert L.gsr_backward_band(fake, fake, fake, fake, fake, fake, fake, 10, 64, 64, 3, -8, 8, 0.1, 0.0, 0, ws,
                               1 << 30, None) == 2
    # uniform batch
    assert L.gsr_workspace_bytes_batch_uniform(0, 10, 64, 64) > 0
    assert L.gsr_workspace_bytes_batch_uniform(-1, 10, 64, 64) == 0
    assert L.gsr_workspace_bytes_batch_uniform(4, 10, 1, 64) == 0
    one, four = L.gsr_workspace_bytes(10, 64, 64), L.gsr_workspace_bytes_batch_uniform(4, 10, 64, 64)
    assert four == L.gsr_workspace_bytes(40, 256, 64) and four >= one      # four samples: one 256-row stack
    assert L.gsr_workspace_bytes_batch_uniform(4, 10, 60, 64) == L.gsr_workspace_bytes(10, 60, 64)  # h % 8: per sample
    fu = L.gsr_forward_batch_uniform
    assert fu(fake, fake, fake, fake, -1, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 5
    assert fu(fake, fake, fake, None, 2, 10, 64, 64, 3, 0.1, 0.0, 0, ws, 1 << 30, None) == 1
    assert fu(fake, fake, fake, fake, 2, 10, 64, 64, 4, 0.1, 0.0, 0, ws, 1 << 30, None) == 3
    assert fu(fake, fake, fake, fake, 0, 10, 64, 64, 3, 0.1, 0.0, 0, None, 0, None) == 0          # empty batch
    # window destination
    win = _lib.GsrWindow()
    win.row_stride, win.pix_stride, win.chan_stride, win.nclip = 64, 1, 4096, 0
    fw = L.gsr_forward_window
    assert fw(fake, fake, fake, fake, None, 10, 64, 64, 3, 0.1, 0.0, 1, ws, 1 << 30, None) == 1     # no window
    win.nclip = _lib.GSR_MAX_CLIP + 1
    assert fw(fake, fake, fake, fake, ctypes.byref(win), 10, 64, 64, 3, 0.1, 0.0, 1, ws, 1 << 30, None) == 5
    win.nclip = 0
    assert fw(fake, fake, fake, fake, ctypes.byref(win), 10, 64, 64, 3, 0.1, 0.0, _lib.GSR_FLAG_CHW, ws, 1 << 30, None) == 5
    # uint8 output needs OVERWRITE and the (h,w,3) layout; BGR needs U8
    f = L.gsr_forward
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, _lib.GSR_FLAG_U8, ws, 1 << 30, None) == 5
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, _lib.GSR_FLAG_U8 | 1 | _lib.GSR_FLAG_CHW, ws, 1 << 30, None) == 5
    assert f(fake, fake, fake, fake, 10, 64, 64, 3, 0.1, 0.0, _lib.GSR_FLAG_BGR | 1, ws, 1 << 30, None) == 5


def test_band_rows_and_tile_regions_need_no_gpu():
    from gsasr_b200 import sharding
    from gsasr_b200.split_and_joint_image import plan_tiles, tile_regions

    assert sharding.band_rows(2048, 1, 2) == (1024, 1024)
    plan = plan_tiles(1024, 1024, 4.0, 480, 8)
    regs = tile_regions(plan, 4, True)
    assert len(regs) == 9 and all(1 <= len(r) <= 8 for r in regs)
    area = sum((y1 - y0) * (x1 - x0) for r in regs for y0, y1, x0, x1 in r)
    assert area == plan.sr_h * plan.sr_w       # integer scale: the regions tile the whole canvas


@pytest.mark.parametrize("lang,compiler", [("c", "gcc"), ("c++", "g++")])
def test_public_headers_compile_as_c_and_cpp(lang, compiler, tmp_path):
    """include/*.h are plain C declarations (no torch / CUDA types): they must compile on their own."""
    import shutil
    import subprocess

    if shutil.which(compiler) is None:
        pytest.skip(f"{compiler} not available")
    src = tmp_path / ("t.c" if lang == "c" else "t.cpp")
    src.write_text('#include "gsraster.h"\n#include "gsraster_test.h"\n'
                   'int main(void) { gsr_window w; gsr_sample s; (void)w; (void)s; '
                   'return gsr_version() < 0; }\n')
    r = subprocess.run([compiler, "-x", lang, "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_argument_validation_of_the_padded_batch_entry_points():
    L = _lib.load()
    fake, ws = ctypes.c_void_p(256), ctypes.c_void_p(4096)
    hw = (ctypes.c_int * 4)(40, 48, 64, 64)
    assert L.gsr_workspace_bytes_batch_padded(2, 10, 64, 64) == L.gsr_workspace_bytes(20, 128, 64)
    assert L.gsr_workspace_bytes_batch_padded(2, 10, 60, 64) == 0                      # hmax % 8 != 0
    fp = L.gsr_forward_batch_padded
    ok_args = (2, 10, 64, 64, hw, None, 0.1, 0.0, 1, ws, 1 << 30, None)
    assert fp(fake, fake, fake, None, *ok_args) == 1                                     # imgs NULL
    assert fp(fake, fake, fake, fake, 2, 10, 60, 64, hw, None, 0.1, 0.0, 1, ws, 1 << 30, None) == 2   # hmax % 8
    assert fp(fake, fake, fake, fake, 2, 10, 64, 64, None, None, 0.1, 0.0, 1, ws, 1 << 30, None) == 1  # no sizes
    big = (ctypes.c_int * 4)(40, 48, 65, 64)
    assert fp(fake, fake, fake, fake, 2, 10, 64, 64, big, None, 0.1, 0.0, 1, ws, 1 << 30, None) == 2   # h_b > hmax
    tiny = (ctypes.c_int * 4)(40, 48, 64, 1)
    assert fp(fake, fake, fake, fake, 2, 10, 64, 64, tiny, None, 0.1, 0.0, 1, ws, 1 << 30, None) == 2  # w_b < 2
    assert fp(fake, fake, fake, fake, 2, 10, 64, 64, hw, None, 0.1, 0.0, _lib.GSR_FLAG_CHW, ws, 1 << 30, None) == 5
    assert fp(fake, fake, fake, fake, 0, 10, 64, 64, hw, None, 0.1, 0.0, 1, None, 0, None) == 0         # empty batch
    st = (ctypes.c_float * 2)(0.3, 0.0)
    assert L.gsr_frontend_forward_batch_padded(fake, fake, fake, 2, 10, 64, 64, hw, st, None, 0.1, 0.0, ws, 1 << 30,
                                               None) == 5                                 # step <= 0


def test_argument_validation_of_the_fused_loss():
    L = _lib.load()
    fake = ctypes.c_void_p(256)
    ws = ctypes.c_void_p(4096)
    st = (ctypes.c_longlong * 4)(1, 1, 1, 1)
    hw = (ctypes.c_int * 4)(8, 8, 4, 9)
    need = L.gsr_l1_crop_workspace_bytes()
    f = L.gsr_l1_crop_loss
    assert need >= 256 + 8
    assert f(None, st, fake, st, fake, fake, 2, 8, 8, hw, 1.0, 0, ws, need, None) == 1        # sr NULL
    assert f(fake, st, fake, st, fake, None, 2, 8, 8, hw, 1.0, 0, ws, need, None) == 1        # loss NULL
    assert f(fake, st, fake, st, fake, fake, 2, 0, 8, hw, 1.0, 0, ws, need, None) == 2        # hmax < 1
    assert f(fake, st, fake, st, fake, fake, 2, 8, 8, hw, 1.0, 0, ws, need, None) == 2        # w_1 = 9 > wmax
    assert f(fake, st, fake, st, fake, fake, 1, 8, 8, hw, 1.0, 0, ws, 16, None) == 4          # workspace too small
    assert f(fake, st, fake, st, fake, fake, 1, 8, 8, hw, 1.0, 0, None, need, None) == 4


def test_argument_validation_of_the_head_tail():
    L = _lib.load()
    fake = ctypes.c_void_p(256)
    f = L.gsr_head_tail_forward
    assert f(None, fake, fake, fake, fake, fake, fake, fake, 128, 8, 16, None) == 1     # x NULL
    assert f(fake, fake, fake, fake, fake, fake, fake, None, 128, 8, 16, None) == 1     # raw NULL
    assert f(fake, fake, fake, fake, fake, fake, fake, fake, -1, 8, 16, None) == 2      # m < 0
    assert f(fake, fake, fake, fake, fake, fake, fake, fake, 128, 0, 16, None) == 2     # empty grid
    assert f(fake, fake, fake, fake, fake, fake, fake, fake, 0, 8, 16, None) == 0       # nothing to do


def test_head_tail_weight_packing_needs_no_gpu():
    """PackedHeadTail: C = 180 / 4C = 720 zero-padded to 192 / 768, rows of the last Linear in output order."""
    import torch
    import torch.nn as nn
    from gsasr_b200 import head_tail

    torch.manual_seed(0)
    blks = [nn.Sequential(nn.Linear(180, 180), nn.ReLU(), nn.Linear(180, 720), nn.ReLU(), nn.Linear(720, k)) for k in (2, 1, 1, 3, 2)]
    pk = head_tail.PackedHeadTail(blks, "cpu")
    assert pk.w1.shape == (5 * 192, 192) and pk.w2.shape == (5 * 768, 192) and pk.w3.shape == (9, 768)
    assert pk.w1.dtype == torch.bfloat16 and pk.b2.dtype == torch.float32
    assert float(pk.w1[180:192].abs().max()) == 0 and float(pk.w1[:, 180:].abs().max()) == 0     # padding
    assert float(pk.w2[720:768].abs().max()) == 0 and float(pk.w3[:, 720:].abs().max()) == 0
    assert torch.equal(pk.w3[4:7, :720], blks[3][4].weight.detach())                             # rgb rows 4..6
    assert torch.equal(pk.b3, torch.cat([b[4].bias.detach() for b in blks]))
    with pytest.raises(RuntimeError):
        head_tail.PackedHeadTail([nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 1024), nn.ReLU(), nn.Linear(1024, k)) for k in (2, 1, 1, 3, 2)], "cpu")
