"""CPU emulation of the forward kernel's CULLING semantics, built from the library's own host
hooks (include/gsraster_test.h): which (Gaussian, pixel) pairs the sm_100a kernel evaluates.
Values are computed in float64 from fp32-rounded dx, dy, so any difference against the oracle is
due to culling (k-sigma truncation, region masks), never to arithmetic.  Test infrastructure."""
from __future__ import annotations

import ctypes

import numpy as np

from gsasr_b200 import _lib


def pix_coords(n):
    return (2.0 * np.arange(n, dtype=np.float64) / (n - 1) - 1.0).astype(np.float32)


def host_setup(sigmas, coords, colors, h, w, dmax, ksigma):
    L = _lib.load()
    s = sigmas.shape[0]
    out = np.zeros((max(s, 1), 9), dtype=np.int32)
    L.gsr_host_setup(sigmas.ctypes.data, coords.ctypes.data, colors.ctypes.data, s, h, w,
                     float(dmax), float(ksigma), out.ctypes.data)
    return out[:s]


def geometry():
    L = _lib.load()
    v = [ctypes.c_int() for _ in range(5)]
    L.gsr_host_geometry(*[ctypes.byref(x) for x in v])
    return tuple(x.value for x in v)


def emulate_forward(sigmas, coords, colors, h, w, dmax, ksigma):
    """Returns (img float64 (h,w,3), pairs evaluated, pairs inside the reference window)."""
    L = _lib.load()
    sigmas = np.ascontiguousarray(sigmas, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    colors = np.ascontiguousarray(colors, np.float32)
    TW, TH, BIN, REG, _ = geometry()
    NRX, NRY = TW // REG, TH // REG
    st = host_setup(sigmas, coords, colors, h, w, dmax, ksigma)
    px, py = pix_coords(w), pix_coords(h)
    img = np.zeros((h, w, 3))
    npairs = 0
    for g in np.nonzero(st[:, 0])[0]:
        _, x0, x1, y0, y1, binds, _, _, _ = st[g]
        sx, sy, rho = (float(v) for v in sigmas[g])
        w1 = -0.5 / (1.0 - rho * rho)
        for ty in range(y0 // TH, y1 // TH + 1):
            for tx in range(x0 // TW, x1 // TW + 1):
                m = L.gsr_host_region_mask(sigmas.ctypes.data, coords.ctypes.data, colors.ctypes.data,
                                           int(g), h, w, float(dmax), float(ksigma), tx * TW, ty * TH)
                for r in range(NRX * NRY):
                    if not (m >> r) & 1:
                        continue
                    ry, rx = divmod(r, NRX)
                    xa, ya = tx * TW + rx * REG, ty * TH + ry * REG
                    xb, yb = min(xa + REG, w), min(ya + REG, h)
                    if binds:  # exact predicate: only pixels of the cull box
                        xa, xb, ya, yb = max(xa, x0), min(xb, x1 + 1), max(ya, y0), min(yb, y1 + 1)
                    if xa >= xb or ya >= yb:
                        continue
                    dx = (px[xa:xb] - coords[g, 0]).astype(np.float64)[None, :]
                    dy = (py[ya:yb] - coords[g, 1]).astype(np.float64)[:, None]
                    q = dx * dx / (sx * sx) - 2 * rho * dx * dy / (sx * sy) + dy * dy / (sy * sy)
                    v = np.exp(w1 * q)
                    img[ya:yb, xa:xb, :] += v[:, :, None] * colors[g].astype(np.float64)[None, None, :]
                    npairs += v.size
    return img, npairs


def emulate_forward_cells(sigmas, coords, colors, h, w, dmax, ksigma):
    """Same for the region-bucket fast path: exactly the (Gaussian, 4x4-pixel cell) pairs the bucket entries
    name (gsr_host_entries: 16x8-pixel regions, 8-bit cell masks).  Returns (img float64, pairs evaluated)."""
    L = _lib.load()
    sigmas = np.ascontiguousarray(sigmas, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    colors = np.ascontiguousarray(colors, np.float32)
    st = host_setup(sigmas, coords, colors, h, w, dmax, ksigma)
    px, py = pix_coords(w), pix_coords(h)
    img = np.zeros((h, w, 3))
    npairs = 0
    cap = 4096
    out = np.zeros((cap, 3), dtype=np.int32)
    for g in np.nonzero(st[:, 0])[0]:
        _, x0, x1, y0, y1, binds, _, _, _ = st[g]
        n = L.gsr_host_entries(sigmas.ctypes.data, coords.ctypes.data, colors.ctypes.data, int(g), h, w,
                               float(dmax), float(ksigma), out.ctypes.data, cap)
        assert n <= cap
        sx, sy, rho = (float(v) for v in sigmas[g])
        w1 = -0.5 / (1.0 - rho * rho)
        for c, b, m in out[:n]:
            for cell in range(8):
                if not (m >> cell) & 1:
                    continue
                xa, ya = c * 16 + (cell & 3) * 4, b * 8 + (cell >> 2) * 4
                xb, yb = min(xa + 4, w), min(ya + 4, h)
                if binds:
                    xa, xb, ya, yb = max(xa, x0), min(xb, x1 + 1), max(ya, y0), min(yb, y1 + 1)
                if xa >= xb or ya >= yb:
                    continue
                dx = (px[xa:xb] - coords[g, 0]).astype(np.float64)[None, :]
                dy = (py[ya:yb] - coords[g, 1]).astype(np.float64)[:, None]
                q = dx * dx / (sx * sx) - 2 * rho * dx * dy / (sx * sy) + dy * dy / (sy * sy)
                v = np.exp(w1 * q)
                img[ya:yb, xa:xb, :] += v[:, :, None] * colors[g].astype(np.float64)[None, None, :]
                npairs += v.size
    return img, npairs


def emulate_backward_cells(sigmas, coords, colors, grads, h, w, dmax, ksigma):
    """The region backward's culling semantics: every Gaussian's gradients summed over exactly the (Gaussian,
    4x4-pixel cell) pairs its bucket entries name (pixels of the cull box only where the dmax window binds) -- what
    gsr_backward_region_kernel evaluates -- with the analytic derivatives of gs.cu:133-159 in float64.
    Returns (g_sigmas, g_coords, g_colors), float64."""
    L = _lib.load()
    sigmas = np.ascontiguousarray(sigmas, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    colors = np.ascontiguousarray(colors, np.float32)
    grads = np.asarray(grads, np.float64)
    st = host_setup(sigmas, coords, colors, h, w, dmax, ksigma)
    px, py = pix_coords(w), pix_coords(h)
    s = sigmas.shape[0]
    gs, gc, gk = np.zeros((s, 3)), np.zeros((s, 2)), np.zeros((s, 3))
    cap = 4096
    out = np.zeros((cap, 3), dtype=np.int32)
    for g in np.nonzero(st[:, 0])[0]:
        _, x0, x1, y0, y1, binds, _, _, _ = st[g]
        n = L.gsr_host_entries(sigmas.ctypes.data, coords.ctypes.data, colors.ctypes.data, int(g), h, w,
                               float(dmax), float(ksigma), out.ctypes.data, cap)
        assert n <= cap
        sx, sy, rho = (float(v) for v in sigmas[g])
        om = 1.0 - rho * rho
        w1 = -0.5 / om
        col = colors[g].astype(np.float64)
        for c, b, m in out[:n]:
            for cell in range(8):
                if not (m >> cell) & 1:
                    continue
                xa, ya = c * 16 + (cell & 3) * 4, b * 8 + (cell >> 2) * 4
                xb, yb = min(xa + 4, w), min(ya + 4, h)
                if binds:
                    xa, xb, ya, yb = max(xa, x0), min(xb, x1 + 1), max(ya, y0), min(yb, y1 + 1)
                if xa >= xb or ya >= yb:
                    continue
                dx = (px[xa:xb] - coords[g, 0]).astype(np.float64)[None, :]
                dy = (py[ya:yb] - coords[g, 1]).astype(np.float64)[:, None]
                q = dx * dx / (sx * sx) - 2 * rho * dx * dy / (sx * sy) + dy * dy / (sy * sy)
                v = np.exp(w1 * q)
                gp = grads[ya:yb, xa:xb, :]
                gk[g] += (v[:, :, None] * gp).sum((0, 1))
                u = v * (gp * col[None, None, :]).sum(2)        # dL/dvalue per pixel
                gc[g, 0] += (u * w1 * -(2 * dx / (sx * sx) - 2 * rho * dy / (sx * sy))).sum()
                gc[g, 1] += (u * w1 * -(2 * dy / (sy * sy) - 2 * rho * dx / (sx * sy))).sum()
                gs[g, 0] += (u * w1 * (-2 * dx * dx / sx ** 3 + 2 * rho * dx * dy / (sx * sx * sy))).sum()
                gs[g, 1] += (u * w1 * (-2 * dy * dy / sy ** 3 + 2 * rho * dx * dy / (sx * sy * sy))).sum()
                gs[g, 2] += (u * (-rho / (om * om) * q + w1 * -2 * dx * dy / (sx * sy))).sum()
    return gs, gc, gk
