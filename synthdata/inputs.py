"""Synthetic inputs of the hot path (binary images in the format of utils.py:80-81, dense training targets with the value
conventions of utils.py:83-228): used by tests, bench.py and the profiling tools. Derived from ``detrand`` so that the same bits
are produced on every box."""
from __future__ import annotations

import numpy as np

from . import detrand


def binary_images(seed: int, B: int, H: int = 512, W: int = 512, p: float = 0.05) -> np.ndarray:
    """[B,1,H,W] float32 in {0,1} -- the input format of utils.py:80-81 / utils_for_test.py:22-27."""
    return (detrand.uniform(detrand.key("img", seed), (B, 1, H, W)) < np.float32(p)).astype(np.float32)


def dense_targets(seed: int, B: int, H: int = 128, W: int = 128, n_omega: int = 60, n_types: int = 6,
                  n_atoms: int = 25, n_bonds: int = 27):
    """Dense training targets with the value conventions of utils.py:83-228
    ({0, 0.8, 1} centre maps, {0, 0.5, 1} class maps, rho / omega in float64).
    Returns the 8 arrays in the order of train.py:86-87 (without the image)."""
    k = detrand.key("targets", seed, B, H, W)
    ta = np.zeros((B, 1, H, W), np.float32)
    tt = np.zeros((B, 14, H, W), np.float32)
    tc = np.zeros((B, 3, H, W), np.float32)
    th = np.zeros((B, 2, H, W), np.float32)
    tb = np.zeros((B, 1, H, W), np.float32)
    tbt = np.zeros((B, n_types, n_omega, H, W), np.float32)
    tr = np.zeros((B, n_omega, H, W), np.float64)
    tw = np.zeros((B, n_omega, H, W), np.float64)
    ra = detrand.integers(k + 1, (B, n_atoms, 5), 0, 1 << 30)
    rb = detrand.integers(k + 2, (B, n_bonds, 5), 0, 1 << 30)
    rr = detrand.uniform(k + 3, (B, n_bonds), 2.0, 12.0)
    for b in range(B):
        for x, y, t, c, h in ra[b].tolist():
            x, y = x % H, y % W
            x0, y0 = max(x - 1, 0), max(y - 1, 0)
            ta[b, 0, x0:x + 2, y0:y + 2] = 0.8
            ta[b, 0, x, y] = 1
            tt[b, t % 14, x0:x + 2, y0:y + 2] = 0.5
            tt[b, t % 14, x, y] = 1
            tc[b, c % 3, x0:x + 2, y0:y + 2] = 0.5
            tc[b, c % 3, x, y] = 1
            if h % 3 < 2:
                th[b, h % 3, x0:x + 2, y0:y + 2] = 0.5
                th[b, h % 3, x, y] = 1
        for j, (x, y, t, w, d) in enumerate(rb[b].tolist()):
            x, y, t, w = x % H, y % W, t % n_types, w % (n_omega // 2)
            x0, y0 = max(x - 1, 0), max(y - 1, 0)
            tb[b, 0, x0:x + 2, y0:y + 2] = 0.8
            tb[b, 0, x, y] = 1
            rho = float(rr[b, j])
            ws = [w + (n_omega // 2) * (d % 2)] if t >= 4 else [w, w + n_omega // 2]
            for wi in ws:
                w0 = max(wi - 1, 0)
                tr[b, w0:wi + 2, x0:x + 2, y0:y + 2] = rho
                tw[b, w0:wi + 2, x0:x + 2, y0:y + 2] = 0.8
                tw[b, wi, x, y] = 1
                tbt[b, t, w0:wi + 2, x0:x + 2, y0:y + 2] = 0.5
                tbt[b, t, wi, x, y] = 1
                wrap = -1 if wi == 0 else (0 if wi == n_omega - 1 else None)
                if wrap is not None:
                    tr[b, wrap, x0:x + 2, y0:y + 2] = rho
                    tw[b, wrap, x0:x + 2, y0:y + 2] = 0.8
                    tbt[b, t, wrap, x0:x + 2, y0:y + 2] = 0.5
    return ta, tt, tc, th, tb, tbt, tr, tw
