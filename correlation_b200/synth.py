"""Synthetic speckle inputs with known ground truth.

The reference ships no images (its AUTO_PILOT sequence lives on the author's disk,
mainapp.cpp:387-408), so every test / bench input is generated here:

    f(x, y) = sum_k a_k cos(kx_k x + ky_k y + phi_k),   wavelengths U(5, 20) px
    und     = clip(rint(128 + contrast * f / std f))            (u8, row-major)
    def(X)  = f(W^-1(X))   with W the deformation model of model_class.cpp:150-202
              (affine) or its 12-parameter quadratic extension, so that
              def(W(x)) == und(x) before u8 rounding and the truth is known.

Works on numpy arrays (tests) or torch tensors on a CUDA device (bench, where a
16384^2 field would take minutes on the host).
"""
from __future__ import annotations

import math

import numpy as np

N_WAVES = 48


def speckle_waves(seed: int, n_waves: int = N_WAVES, spectrum=None):
    """spectrum=None: wavelengths U(5, 20) px. spectrum=(lo, hi): log-uniform in [lo, hi]
    px -- multi-scale speckle, so that coarse pyramid levels still carry signal (config 5)."""
    rng = np.random.default_rng(seed)
    if spectrum is None:
        lam = rng.uniform(5.0, 20.0, n_waves)
    else:
        lam = np.exp(rng.uniform(math.log(spectrum[0]), math.log(spectrum[1]), n_waves))
    theta = rng.uniform(0.0, 2.0 * math.pi, n_waves)
    amp = rng.uniform(0.5, 1.0, n_waves)
    phase = rng.uniform(0.0, 2.0 * math.pi, n_waves)
    k = 2.0 * math.pi / lam
    kx = k * np.cos(theta)
    ky = k * np.sin(theta)
    # std of a sum of independent-phase cosines
    std = math.sqrt(float(np.sum(amp**2) / 2.0))
    return kx, ky, amp, phase, std


def inverse_warp(X, Y, params, cx, cy, xp=np, iters: int = 8):
    """Solve W(x, y) = (X, Y) for (x, y).

    params: 6 values (u, v, ux, uy, vx, vy) or 12 (… + uxx, uxy, uyy, vxx, vxy, vyy) with
    x' = x + u + ux dx + uy dy + 1/2 uxx dx^2 + uxy dx dy + 1/2 uyy dy^2 (same for y'),
    dx = x - cx, dy = y - cy.
    """
    p = [float(v) for v in params] + [0.0] * (12 - len(params))
    u, v, ux, uy, vx, vy, uxx, uxy, uyy, vxx, vxy, vyy = p
    a11, a12, a21, a22 = 1.0 + ux, uy, vx, 1.0 + vy
    det = a11 * a22 - a12 * a21
    rx = X - cx - u
    ry = Y - cy - v
    dx = (a22 * rx - a12 * ry) / det
    dy = (-a21 * rx + a11 * ry) / det
    if any(abs(q) > 0 for q in (uxx, uxy, uyy, vxx, vxy, vyy)):
        for _ in range(iters):  # fixed point on the (small) quadratic part
            qx = 0.5 * uxx * dx * dx + uxy * dx * dy + 0.5 * uyy * dy * dy
            qy = 0.5 * vxx * dx * dx + vxy * dx * dy + 0.5 * vyy * dy * dy
            sx = rx - qx
            sy = ry - qy
            dx = (a22 * sx - a12 * sy) / det
            dy = (-a21 * sx + a11 * sy) / det
    return dx + cx, dy + cy


def _field_block(xs, ys, waves, xp):
    kx, ky, amp, phase, std = waves
    f = xp.zeros_like(xs)
    for i in range(len(kx)):
        f = f + float(amp[i]) * xp.cos(float(kx[i]) * xs + float(ky[i]) * ys + float(phase[i]))
    return f / std


def make_image(rows, cols, seed, params=None, center=None, contrast=45.0,
               device=None, block_rows=512, spectrum=None, n_waves=N_WAVES):
    """u8 image (numpy uint8 [rows, cols], or a torch uint8 CUDA tensor if device is given).

    params=None -> the undeformed field; else the field sampled at W^-1 (the deformed image).
    """
    waves = speckle_waves(seed, n_waves, spectrum)
    if device is None:
        xp = np
        out = np.empty((rows, cols), np.uint8)
        mk = lambda a, b: np.meshgrid(np.arange(cols, dtype=np.float64),
                                      np.arange(a, b, dtype=np.float64))
    else:
        import torch
        xp = torch
        out = torch.empty((rows, cols), dtype=torch.uint8, device=device)

        def mk(a, b):
            yy, xx = torch.meshgrid(
                torch.arange(a, b, dtype=torch.float64, device=device),
                torch.arange(cols, dtype=torch.float64, device=device), indexing="ij")
            return xx, yy
    cx, cy = center if center is not None else (cols / 2.0, rows / 2.0)
    for r0 in range(0, rows, block_rows):
        r1 = min(rows, r0 + block_rows)
        X, Y = mk(r0, r1)
        if params is not None:
            X, Y = inverse_warp(X, Y, params, cx, cy, xp=xp)
        f = _field_block(X, Y, waves, xp)
        img = 128.0 + contrast * f
        if device is None:
            out[r0:r1] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        else:
            out[r0:r1] = xp.clamp(xp.round(img), 0, 255).to(xp.uint8)
    return out


def make_pair(rows, cols, seed, params, center=None, **kw):
    """(und, def) pair with def = und warped by `params` about `center`."""
    und = make_image(rows, cols, seed, None, center, **kw)
    dfm = make_image(rows, cols, seed, params, center, **kw)
    return und, dfm


def star_polygon(cx, cy, mean_radius, n_vertices=64, seed=3, wobble=0.25):
    """Star-shaped non-convex simple polygon (C3 blob contour), float32 [n, 2]."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0.0, 2.0 * math.pi, n_vertices, endpoint=False)
    rad = mean_radius * (1.0 + wobble * rng.uniform(-1.0, 1.0, n_vertices))
    pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
    return pts.astype(np.float32)
