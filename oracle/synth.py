"""Synthetic inputs for the hot path (test infrastructure).

Everything is derived from ``detrand`` so the same bits are produced in the build container
(where goldens are minted from the reference) and on the GPU box (where they are checked).
"""
from __future__ import annotations

import numpy as np

from . import detrand

V2_HEADS = (1, 14, 3, 2, 1, 360, 60, 60)


from synthdata.inputs import binary_images, dense_targets  # noqa: E402,F401  (plain synthetic inputs live outside the oracle)


def planted_logits(seed: int, H: int = 128, W: int = 128, n_omega: int = 60, n_types: int = 6,
                   n_atoms=None, n_bonds=None, edge_cases: bool = True):
    """The 8 logit maps of ONE image with planted atom / bond peaks (float32, NCHW without N).

    Background -6 +- 0.3; every planted centre is a 3x3 bump (centre in [1, 4], ring 2 below).
    With ``edge_cases`` the image also contains: corner and border peaks, a 2-pixel plateau (tie ->
    two peaks), a value exactly equal to the -1 threshold (not a peak), omega peaks at bins
    0 / 29 / 30 / 59, exactly tied antipodal omega pairs (exercise the asymmetric ``<`` / ``<=`` of
    img2smiles.py:143-158), two omega peaks at one bond, and one bond whose omega column is flat.
    Returns (outs, info) with info = dict(atoms=[(x,y)], bonds=[(x,y)]).
    """
    k = detrand.key("planted", seed, H, W)
    chans = (1, 14, 3, 2, 1, n_types * n_omega, n_omega, n_omega)
    outs = []
    for i, c in enumerate(chans):
        noise = detrand.uniform(k + 101 * i, (c, H, W), -0.3, 0.3)
        if i in (0, 4, 7):
            outs.append(noise - np.float32(6.0))
        elif i == 6:
            outs.append(noise * np.float32(10.0))             # rho: signed, |.| is taken by the decoder
        else:
            outs.append(noise * np.float32(3.0))              # class maps: arbitrary but tie-free argmax
    za, zt, zc, zh, zb, zbt, zr, zw = outs
    if n_atoms is None:
        n_atoms = int(detrand.integers(k + 1, (), 10, 61))
    if n_bonds is None:
        n_bonds = int(detrand.integers(k + 2, (), 10, 71))

    def place(n, salt, occupied):
        pts = []
        cand = detrand.integers(k + salt, (n * 40, 2), 0, 1 << 30)
        for cx, cy in cand.tolist():
            x, y = 2 + cx % (H - 4), 2 + cy % (W - 4)
            if all(abs(x - a) > 3 or abs(y - b) > 3 for a, b in occupied + pts):
                pts.append((x, y))
                if len(pts) == n:
                    break
        return pts

    def bump(z, x, y, v):
        x0, x1, y0, y1 = max(x - 1, 0), min(x + 2, H), max(y - 1, 0), min(y + 2, W)
        z[x0:x1, y0:y1] = np.float32(v - 2.0)
        z[x, y] = np.float32(v)

    corners = [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (0, W // 2), (H // 2, 0)] if edge_cases else []
    atoms = place(n_atoms, 11, corners)
    bonds = place(n_bonds, 12, corners)
    vals_a = detrand.uniform(k + 21, (len(atoms) + 16,), 1.0, 4.0)
    vals_b = detrand.uniform(k + 22, (len(bonds) + 16,), 1.0, 4.0)
    cls = detrand.integers(k + 23, (len(atoms) + 16, 3), 0, 1 << 20)
    for i, (x, y) in enumerate(atoms):
        bump(za[0], x, y, vals_a[i])
        zt[cls[i, 0] % 14, x, y] = 5.0
        zc[cls[i, 1] % 3, x, y] = 5.0
        zh[cls[i, 2] % 2, x, y] = 5.0
    wsel = detrand.integers(k + 24, (len(bonds) + 16, 4), 0, 1 << 20)
    special_w = [0, 29, 30, 59, 28, 31, 1, 58]
    for i, (x, y) in enumerate(bonds):
        bump(zb[0], x, y, vals_b[i])
        w = special_w[i] if (edge_cases and i < len(special_w)) else int(wsel[i, 0] % n_omega)
        kind = int(wsel[i, 1] % 4)
        col = zw[:, x, y]
        col[w] = 3.0
        col[(w - 1) % n_omega] = 1.0
        col[(w + 1) % n_omega] = 1.0
        anti = (w + n_omega // 2) % n_omega
        if kind == 0:                      # undirected bond: exactly tied antipodal peak
            col[anti] = 3.0
        elif kind == 1:                    # antipodal slightly weaker
            col[anti] = 2.5
        elif kind == 2:                    # a second, unrelated omega peak (ring-fusion crossing)
            w2 = (w + 7 + int(wsel[i, 2] % 10)) % n_omega
            col[w2] = 2.0
        t = int(wsel[i, 3] % n_types)
        zbt[t * n_omega + w, x, y] = 6.0
        zbt[t * n_omega + anti, x, y] = 6.0
    info = dict(atoms=list(atoms), bonds=list(bonds))
    if edge_cases:
        for j, (x, y) in enumerate(corners):
            bump(za[0], x, y, 2.0 + 0.1 * j)
            info["atoms"].append((x, y))
        bump(zb[0], 0, 0, 2.5)
        zw[5, 0, 0] = 2.0
        info["bonds"].append((0, 0))
        # plateau: two horizontally adjacent equal maxima -> both are peaks
        px, py = H // 2 + 1, W // 2 + 1
        bump(za[0], px, py, 1.5)
        za[0, px, py + 1] = za[0, px, py]
        # exactly-at-threshold centre: value == -1 is NOT a peak (strict >)
        qx, qy = 3, W - 6
        za[0, qx - 1:qx + 2, qy - 1:qy + 2] = -3.0
        za[0, qx, qy] = -1.0
        zb[0, qx - 1:qx + 2, qy - 1:qy + 2] = -3.0
        zb[0, qx, qy] = np.nextafter(np.float32(-1.0), np.float32(0.0))      # just above -> peak
        zw[:, qx, qy] = -1.0                                                 # flat column at threshold: no candidate
        info["bonds"].append((qx, qy))
    return [za, zt, zc, zh, zb, zbt, zr, zw], info


def random_logits(seed: int, B: int, H: int = 128, W: int = 128, heads=V2_HEADS, scale: float = 2.0):
    k = detrand.key("logits", seed, B, H, W)
    return [detrand.normalish(k + i, (B, h, H, W), scale) for i, h in enumerate(heads)]


# Pseudo-molecule drawings with labels known by construction live in synthdata.molecules (a plain synthetic-input generator, also
# used by tools/shard_infer.py); the target rasterisation below restates reference rules and stays here.
from synthdata.molecules import ATOM_LETTERS, pseudo_molecules  # noqa: E402,F401


def rasterise_targets(labels, H4: int = 128, W4: int = 128, n_omega: int = 60, f64: bool = True):
    """Dense targets at stride 4 from labels, restating the rules of utils.py:94-228: 3x3 neighbourhoods at 0.8 / 0.5 with
    the exact centre at 1, bond angle bin floor((atan(dy/dx) + pi/2) / (pi/30)) after flipping the vector to dx >= 0,
    undirected bonds marked at omega and omega + 30, wedge bonds (types 4, 5) only at the directed bin, wrap-around of the
    neighbouring bin at 0 / 59, rho = half bond length in map pixels; rho / omega arrays float64 like the reference
    (utils.py:91-92). Returns the 8 arrays in the order of train.py:86-87."""
    B = len(labels)
    half = n_omega // 2
    dt = np.float64 if f64 else np.float32
    ta = np.zeros((B, 1, H4, W4), np.float32)
    tt = np.zeros((B, 14, H4, W4), np.float32)
    tc = np.zeros((B, 3, H4, W4), np.float32)
    th = np.zeros((B, 2, H4, W4), np.float32)
    tb = np.zeros((B, 1, H4, W4), np.float32)
    tbt = np.zeros((B, 6, n_omega, H4, W4), np.float32)
    tr = np.zeros((B, n_omega, H4, W4), dt)
    tw = np.zeros((B, n_omega, H4, W4), dt)
    step = np.pi / half
    for b, lab in enumerate(labels):
        for (x, y, t, charge, hs) in lab["atoms"]:
            x, y = int(x) // 4, int(y) // 4
            x0, y0 = max(x - 1, 0), max(y - 1, 0)
            ta[b, 0, x0:x + 2, y0:y + 2] = 0.8
            ta[b, 0, x, y] = 1
            tt[b, t, x0:x + 2, y0:y + 2] = 0.5
            tt[b, t, x, y] = 1
            tc[b, charge, x0:x + 2, y0:y + 2] = 0.5
            tc[b, charge, x, y] = 1
            if hs in (0, 1):
                th[b, hs, x0:x + 2, y0:y + 2] = 0.5
                th[b, hs, x, y] = 1
        for (x, y, dx, dy, t, direction) in lab["bonds"]:
            x, y, dx, dy = int(x) // 4, int(y) // 4, dx / 4.0, dy / 4.0
            if dx < 0:
                dx, dy = -dx, -dy
                direction = 1 - direction if t >= 4 else direction
            elif dx == 0:
                if dy > 0:
                    direction = 1
                dy = -abs(dy)
            rho = float(np.sqrt(dx * dx + dy * dy))
            w = int(np.floor((np.arctan(dy / (dx + 1e-6)) + np.pi / 2) / step))
            w = min(max(w, 0), half - 1)
            x0, y0 = max(x - 1, 0), max(y - 1, 0)
            tb[b, 0, x0:x + 2, y0:y + 2] = 0.8
            tb[b, 0, x, y] = 1
            bins = [w + half * (direction % 2)] if t >= 4 else [w, w + half]
            for wi in bins:
                w0 = max(wi - 1, 0)
                tr[b, w0:wi + 2, x0:x + 2, y0:y + 2] = rho
                tw[b, w0:wi + 2, x0:x + 2, y0:y + 2] = 0.8
                tw[b, wi, x, y] = 1
                tbt[b, t, w0:wi + 2, x0:x + 2, y0:y + 2] = 0.5
                tbt[b, t, wi, x, y] = 1
                wrap = n_omega - 1 if wi == 0 else (0 if wi == n_omega - 1 else None)
                if wrap is not None:
                    tr[b, wrap, x0:x + 2, y0:y + 2] = rho
                    tw[b, wrap, x0:x + 2, y0:y + 2] = 0.8
                    tbt[b, t, wrap, x0:x + 2, y0:y + 2] = 0.5
    return ta, tt, tc, th, tb, tbt, tr, tw
