"""Restatement of the reference's dense training-target rasterisation (test infrastructure only).

Follows ``/root/reference/src/utils.py:83-228`` (``MolDataset.__getitem__``): from the label strings of one drawing
(``atoms_string`` = ``"C:x,y,charge[,hs];..."``, ``bonds_string`` = ``"type:x,y,dx,dy,stereo,direction;..."``, full-resolution
pixel coordinates, ``rdkit_img_generate.py:136-137,178-180``) and the augmentation parameters (``scale_x/y``, ``ddx/ddy``,
utils.py:44-58) to the nine stride-4 maps. Pinned by ``tests/golden/target_cases.npz``, minted by exec'ing the reference's own
statements (``tests/golden/make_golden.py::golden_targets``). Items are stamped in string order; later stamps overwrite.
"""
from __future__ import annotations

import math

import numpy as np

ATOM_VOCAB = {'<unkonw>': 0, 'C': 1, 'N': 2, 'O': 3, 'P': 4, 'F': 5, 'Cl': 6, 'S': 7, 'Br': 8, 'B': 9, 'Se': 10, 'I': 11, 'H': 12,
              'Si': 13}                                      # utils.py:12-13
CHARGE_VOCAB = {0: 0, 1: 1, -1: 2}                           # utils.py:14
BOND_VOCAB = {1: 0, 2: 1, 3: 2, 4: 3}                        # utils.py:15


def parse_atoms(atoms_string, scale_x=1, scale_y=1, ddx=0, ddy=0):
    """-> list of (x, y, type index, charge index, hs) at stride 4 (utils.py:94-108); hs = -1 when the label has none."""
    out = []
    for item in atoms_string.split(';')[:-1]:
        atom, position = item.split(':')
        if len(atom) == 1:
            atom = atom.upper()
        f = position.split(',')
        x, y, charge = int(int(f[0]) * scale_x + ddx) // 4, int(int(f[1]) * scale_y + ddy) // 4, int(f[2])
        hs = int(f[3]) if len(f) == 4 else -1
        out.append((x, y, ATOM_VOCAB.get(atom, 0), CHARGE_VOCAB.get(charge, 0), hs))
    return out


def parse_bonds(bonds_string, scale_x=1, scale_y=1, ddx=0, ddy=0, n_omega=60):
    """-> list of (x, y, type index, [omega bins], rho) (utils.py:126-160, bins as stamped at :171-228)."""
    out = []
    half, step = n_omega // 2, math.pi / (n_omega // 2)
    for item in bonds_string.split(';')[:-1]:
        bond, position = item.split(':')
        t = BOND_VOCAB.get(int(bond), 0)
        f = position.split(',')
        x, y = int(int(f[0]) * scale_x + ddx) // 4, int(int(f[1]) * scale_y + ddy) // 4
        dx, dy = (int(f[2]) * scale_x) / 4, (int(f[3]) * scale_y) / 4
        stereo, direction = int(f[4]), int(f[5])
        if stereo == 5 or stereo == 1:
            t = 4
        elif stereo == 6:
            t = 5
        if dx < 0:
            dx, dy = -dx, -dy
        elif dx == 0:
            if dy > 0:
                direction = 1
            dy = -abs(dy)
        rho = np.sqrt(dx * dx + dy * dy)
        w = int(np.floor((math.atan(dy / (dx + 1e-6)) + np.pi / 2) / step))
        bins = [w + half if direction == 1 else w] if t in (4, 5) else [w, w + half]
        out.append((x, y, t, bins, float(rho)))
    return out


def rasterise(atoms_string, bonds_string, scale_x=1, scale_y=1, ddx=0, ddy=0, H4=128, W4=128, n_omega=60):
    """The nine arrays of utils.py:83-92 in the order of the reference's return statement minus the image:
    atom_target, atom_type, atom_charge, atom_hs, bond_target, bond_type [6, n_omega, H, W], bond_rho, bond_omega_type."""
    ta = np.zeros((1, H4, W4), np.float32)
    tt = np.zeros((14, H4, W4), np.float32)
    tc = np.zeros((3, H4, W4), np.float32)
    th = np.zeros((2, H4, W4), np.float32)
    tb = np.zeros((1, H4, W4), np.float32)
    tbt = np.zeros((6, n_omega, H4, W4), np.float32)
    tr = np.zeros((n_omega, H4, W4), np.float64)             # utils.py:91-92: float64
    tw = np.zeros((n_omega, H4, W4), np.float64)
    for x, y, t, c, hs in parse_atoms(atoms_string, scale_x, scale_y, ddx, ddy):
        x0, y0 = (0 if x == 0 else x - 1), (0 if y == 0 else y - 1)
        ta[0, x0:x + 2, y0:y + 2] = 0.8
        ta[0, x, y] = 1
        tt[t, x0:x + 2, y0:y + 2] = 0.5
        tt[t, x, y] = 1
        tc[c, x0:x + 2, y0:y + 2] = 0.5
        tc[c, x, y] = 1
        if hs == 0 or hs == 1:
            th[hs, x0:x + 2, y0:y + 2] = 0.5
            th[hs, x, y] = 1
    for x, y, t, bins, rho in parse_bonds(bonds_string, scale_x, scale_y, ddx, ddy, n_omega):
        x0, y0 = (0 if x == 0 else x - 1), (0 if y == 0 else y - 1)
        tb[0, x0:x + 2, y0:y + 2] = 0.8
        tb[0, x, y] = 1
        for wi in bins:
            w0 = 0 if wi == 0 else wi - 1
            tr[w0:wi + 2, x0:x + 2, y0:y + 2] = rho
            tw[w0:wi + 2, x0:x + 2, y0:y + 2] = 0.8
            tw[wi, x, y] = 1
            tbt[t, w0:wi + 2, x0:x + 2, y0:y + 2] = 0.5
            tbt[t, wi, x, y] = 1
            wrap = n_omega - 1 if wi == 0 else (0 if wi == n_omega - 1 else None)
            if wrap is not None:
                tr[wrap, x0:x + 2, y0:y + 2] = rho
                tw[wrap, x0:x + 2, y0:y + 2] = 0.8
                tbt[t, wrap, x0:x + 2, y0:y + 2] = 0.5
    return ta, tt, tc, th, tb, tbt, tr, tw


def label_strings(seed, n_atoms=30, n_bonds=34, size=512):
    """Deterministic label strings exercising every branch of utils.py:94-228: border positions, labels with and without the
    H count, lower-case and unknown symbols, unknown charges, dx < 0 / dx == 0 with either sign of dy, wedge bonds
    (stereo 1 / 5 / 6) with both directions, bins 0 / 29 / 30 / 59 (wrap-around)."""
    from . import detrand
    k = detrand.key("labels", seed)
    ra = detrand.integers(k + 1, (n_atoms, 6), 0, 1 << 30)
    rb = detrand.integers(k + 2, (n_bonds, 8), 0, 1 << 30)
    syms = ['C', 'N', 'O', 'P', 'F', 'Cl', 'S', 'Br', 'B', 'Se', 'I', 'H', 'Si', 'c', 'n', 'Xx', 'se']
    a = []
    for i, (x, y, s, c, h, f) in enumerate(ra.tolist()):
        x, y = x % size, y % size
        if i == 0:
            x, y = 0, 0
        elif i == 1:
            x, y = size - 1, size - 1
        elif i == 2:
            x, y = 0, size - 2
        charge = (0, 0, 1, -1, 2)[c % 5]
        pos = f"{x},{y},{charge}" + (f",{h % 3}" if f % 4 else "")
        a.append(f"{syms[s % len(syms)]}:{pos}")
    b = []
    for i, (x, y, t, dx, dy, st, d, e) in enumerate(rb.tolist()):
        x, y = x % size, y % size
        dx, dy = dx % 121 - 60, dy % 121 - 60
        if i % 7 == 0:
            dx = 0
        if i % 11 == 0:
            dy = 0 if dx else 17
        if i == 3:
            x, y, dx, dy = 0, 5, 40, -1                # bin 29 / 59 side
        if i == 4:
            x, y, dx, dy = 7, 0, 1, -60                # bin 0 / 30 side
        if i == 5:
            dx, dy = 0, 25                             # dx == 0, dy > 0: direction forced to 1
        bt = (1, 2, 3, 4, 7)[t % 5]
        stereo = (0, 0, 0, 1, 5, 6)[st % 6]
        b.append(f"{bt}:{x},{y},{dx},{dy},{stereo},{d % 2}")
    return ";".join(a) + ";", ";".join(b) + ";"
