"""numpy restatement of the reference heat-map decoding stage (test infrastructure).

Follows ``/root/reference/src/img2smiles.py``:
  * :62-68   centre peaks: ``(max_pool2d(z, 3, 1, 1) == z) * (z > -1)`` on the RAW logits of
             output 0 (atoms) and output 4 (bonds); -inf padding at the border; plateau ties all count.
  * :70      rho = |z_rho|
  * :72,121  bond type per (omega, x, y) = argmax over the 6 types of channel ``t*60 + omega``
  * :74-80   circular 3-tap NMS over the 60 omega bins + ``z > -1``
  * :116-118 atom type / charge / H-count = argmax over 14 / 3 / 2 channels (first max wins)
  * :134-171 per bond peak (row-major ``nonzero`` order), candidates in ascending omega, the
             half-circle rejection test :143-158 (asymmetric ``<`` / ``<=``), emit position, type, delta
  * :177-193 per atom peak, greedy de-duplication (squared distance < 4 to an accepted atom)
``omega_mode='raw'`` reproduces ``img2smiles2.py:139`` (candidates = every omega whose raw logit != 0).

Two levels are provided: ``decode_records`` (the compact per-image records the CUDA kernel
emits) and ``records_to_lists`` (the Python lists the reference builds at img2smiles.py:131-193,
which the unchanged host assembly consumes).
"""
from __future__ import annotations

import math

import numpy as np

NEG_INF = np.float32(-np.inf)


def _peaks2d(z: np.ndarray, thr: float) -> np.ndarray:
    """z [H,W] float32 -> bool peak map (img2smiles.py:62-64)."""
    H, W = z.shape
    p = np.full((H + 2, W + 2), NEG_INF, np.float32)
    p[1:-1, 1:-1] = z
    m = z.copy()
    for dy in range(3):
        for dx in range(3):
            m = np.maximum(m, p[dy:dy + H, dx:dx + W])
    return (m == z) & (z > np.float32(thr))


def omega_candidates(zw: np.ndarray, thr: float, omega_mode: str = "nms") -> np.ndarray:
    """zw [n_omega] logits at one pixel -> bool candidate mask (img2smiles.py:74-80 / img2smiles2.py:139)."""
    if omega_mode == "raw":
        return zw != 0
    left = np.roll(zw, 1)
    right = np.roll(zw, -1)
    return (np.maximum(np.maximum(left, zw), right) == zw) & (zw > np.float32(thr))


def omega_survives(zw: np.ndarray, w: int) -> bool:
    """Half-circle test of img2smiles.py:143-158 for n_omega = 60 (generalised as n/2 = 30)."""
    n = zw.shape[0]
    h = n // 2
    v = zw[w]
    if w <= h - 2:
        if v < max(zw[w + h - 1], zw[w + h]):
            return False
    elif w == h - 1:
        if v < zw[w + h - 1] or v < zw[0]:
            return False
    elif w == h:
        if v <= zw[0] or v <= zw[n - 1]:
            return False
    else:
        if v <= max(zw[w - h - 1], zw[w - h]):
            return False
    return True


def centre_probability(z: np.ndarray) -> np.ndarray:
    """train.py:95,100: ``torch.clamp(torch.sigmoid(z), 1e-5, 1-1e-5)`` -- executed by torch itself (fp32)."""
    import torch
    return torch.clamp(torch.sigmoid(torch.from_numpy(np.ascontiguousarray(z, np.float32))), 1e-5, 1 - 1e-5).numpy()


def decode_records(outs, thr: float = -1.0, omega_mode: str = "nms", apply_sigmoid: bool = False, thr_omega: float = -1.0):
    """``apply_sigmoid``: centre peaks by the training-time metric rule (train.py:145-151: NMS and ``> thr`` on the
    clamped probabilities, thr = 0.25 there); the omega NMS keeps the logit threshold ``thr_omega``.

    outs: the 8 logit maps of ONE image as float32 arrays
    (atom[1,H,W], type[14,H,W], charge[3,H,W], hs[2,H,W], bond[1,H,W], btype[6*n_w,H,W], rho[n_w,H,W], omega[n_w,H,W]).

    Returns (atoms, bonds): atoms int array [n_a, 5] = (x, y, type, charge, hs) in row-major peak
    order (NOT yet de-duplicated); bonds structured as (int array [n_s, 4] = (x, y, omega, type),
    float32 array [n_s] = rho) for the surviving (peak, omega) pairs in reference enumeration order.
    """
    za, zt, zc, zh, zb, zbt, zr, zw = [np.asarray(o, np.float32) for o in outs]
    n_w = zw.shape[0]
    n_t = zbt.shape[0] // n_w
    H, W = za.shape[-2:]
    if apply_sigmoid:
        apk = _peaks2d(centre_probability(za.reshape(H, W)), thr)
        bpk = _peaks2d(centre_probability(zb.reshape(H, W)), thr)
    else:
        apk = _peaks2d(za.reshape(H, W), thr)
        bpk = _peaks2d(zb.reshape(H, W), thr)
        thr_omega = thr
    atoms = []
    for x, y in zip(*np.nonzero(apk)):
        atoms.append((x, y, int(zt[:, x, y].argmax()), int(zc[:, x, y].argmax()), int(zh[:, x, y].argmax())))
    b_int, b_rho = [], []
    zbt5 = zbt.reshape(n_t, n_w, H, W)
    for x, y in zip(*np.nonzero(bpk)):
        col = zw[:, x, y]
        cand = omega_candidates(col, thr_omega, omega_mode)
        for w in np.nonzero(cand)[0]:
            if not omega_survives(col, int(w)):
                continue
            b_int.append((x, y, int(w), int(zbt5[:, w, x, y].argmax())))
            b_rho.append(abs(zr[w, x, y]))
    atoms = np.asarray(atoms, np.int32).reshape(-1, 5)
    return atoms, (np.asarray(b_int, np.int32).reshape(-1, 4), np.asarray(b_rho, np.float32))


# vocabularies: utils.py:12-14, inverted at img2smiles.py:24-26 (index 0 -> 'C')
ATOM_SYMBOLS = ['C', 'C', 'N', 'O', 'P', 'F', 'Cl', 'S', 'Br', 'B', 'Se', 'I', 'H', 'Si']
CHARGE_VALUES = [0, 1, -1]


def records_to_lists(atoms, bonds, n_omega: int = 60):
    """Build the lists of img2smiles.py:131-193 from compact records.

    Returns dict with bonds_position_list, bonds_property_list, bonds_delta_list,
    atoms_position_list, atoms_type_list, atoms_charge_list, atoms_hs_list; or None when either
    peak set is empty (img2smiles.py:126-129 -- judged on the PEAK maps, so callers pass
    ``n_bond_peaks`` separately if they need the distinction; here emptiness of records is used).
    """
    b_int, b_rho = bonds
    bp, bt, bd = [], [], []
    for (x, y, w, t), rho in zip(b_int.tolist(), b_rho.tolist()):
        omega = w * (np.pi / (n_omega // 2)) + np.pi / n_omega - np.pi / 2   # :160
        bp.append([x, y])
        bt.append(t)
        bd.append([rho * np.cos(omega), rho * np.sin(omega)])                 # :164 (float64 math)
    ap, at, ac, ah = [], [], [], []
    for x, y, t, c, h in atoms.tolist():
        if ap:
            d = np.sum(np.square(np.array(ap) - np.array([[x, y]])), axis=-1).min()
            if d < 4:                                                         # :186
                continue
        ap.append([x, y])
        at.append(ATOM_SYMBOLS[t])
        ac.append(CHARGE_VALUES[c])
        ah.append(h)
    return dict(bonds_position_list=bp, bonds_property_list=bt, bonds_delta_list=bd,
                atoms_position_list=ap, atoms_type_list=at, atoms_charge_list=ac, atoms_hs_list=ah)


def count_peaks(outs, thr: float = -1.0):
    za, zb = np.asarray(outs[0], np.float32), np.asarray(outs[4], np.float32)
    H, W = za.shape[-2:]
    return int(_peaks2d(za.reshape(H, W), thr).sum()), int(_peaks2d(zb.reshape(H, W), thr).sum())


__all__ = ["decode_records", "records_to_lists", "omega_candidates", "omega_survives", "count_peaks",
           "ATOM_SYMBOLS", "CHARGE_VALUES", "math"]
