"""Pseudo-molecule drawings with labels known by construction (synthetic inputs (ii)) -- a plain
synthetic-input generator (no reference arithmetic): shared by the tests (through oracle.synth) and tools/shard_infer.py.
``label_strings`` writes the labels in the text format of the reference's data generator
(``/root/reference/src/rdkit_img_generate.py:136-137,178-180``), which ``abcnet_b200.parse_labels`` consumes."""
from __future__ import annotations

import numpy as np

from . import detrand

# RDKit / Indigo rendering (rdkit_img_generate.py, indigo_img_generator.py) is not available offline; these line drawings
# play the same role: binarised 1 x H x W images whose atoms / bonds are known, so that the dense targets of
# utils.py:83-228 can be rasterised and a network can be trained on them for the end-to-end peak-set / SMILES parity test.
# class index = position in the reference vocabulary (utils.py:12-13; index 0 = unknown, 1 = carbon, drawn without a letter)
ATOM_LETTERS = ("C", "C", "N", "O", "P", "F", "Cl", "S", "Br", "B", "Se", "I", "H", "Si")


def pseudo_molecules(seed: int, B: int, H: int = 512, W: int = 512):
    """Returns (images [B,1,H,W] float32 in {0,1}, labels) with labels[b] = dict(
         atoms=[(x, y, type, charge, hs)]           x = row, y = column in full-resolution pixels,
         bonds=[(x, y, dx, dy, type, direction)]    centre, half-extent vector (row, col), type 0..5).
    Drawn with cv2 like the binarised renderings of utils_for_test.py:22-27 (ink = 1)."""
    import cv2
    k = detrand.key("pseudo", seed, B, H, W)
    imgs = np.zeros((B, 1, H, W), np.float32)
    labels = []
    margin, min_d, max_bond = 40, 56, 120
    for b in range(B):
        n_want = int(detrand.integers(k + 7 * b + 1, (), 6, 19))
        cand = detrand.integers(k + 7 * b + 2, (n_want * 30, 2), 0, 1 << 30)
        pts = []
        for cx, cy in cand.tolist():
            x, y = margin + cx % (H - 2 * margin), margin + cy % (W - 2 * margin)
            if all((x - a) ** 2 + (y - c) ** 2 >= min_d * min_d for a, c in pts):
                pts.append((x, y))
                if len(pts) == n_want:
                    break
        n = len(pts)
        r = detrand.integers(k + 7 * b + 3, (n, 4), 0, 1 << 30)
        atoms = []
        canvas = np.zeros((H, W), np.uint8)
        for i, (x, y) in enumerate(pts):
            t = 1 if r[i, 0] % 10 < 6 else int(2 + r[i, 1] % 12)
            charge = (0, 0, 0, 0, 1, 2)[int(r[i, 2] % 6)] if t in (2, 3) else 0
            hs = int(r[i, 3] % 2) if t in (2, 3, 7) else 0
            atoms.append((x, y, t, charge, hs))
        # bonds: every atom to its nearest neighbours within max_bond (no duplicates)
        pairs = set()
        for i, (x, y) in enumerate(pts):
            d = sorted((((x - a) ** 2 + (y - c) ** 2), j) for j, (a, c) in enumerate(pts) if j != i)
            for dd, j in d[:2 + int(r[i, 0] % 2)]:
                if dd <= max_bond * max_bond:
                    pairs.add((min(i, j), max(i, j)))
        rb = detrand.integers(k + 7 * b + 4, (max(len(pairs), 1), 3), 0, 1 << 30)
        bonds = []
        for q, (i, j) in enumerate(sorted(pairs)):
            (x0, y0), (x1, y1) = pts[i], pts[j]
            bt = (0, 0, 0, 0, 1, 1, 2, 3, 4, 5)[int(rb[q, 0] % 10)]
            cx, cy = (x0 + x1) // 2, (y0 + y1) // 2
            dx, dy = (x1 - x0) / 2.0, (y1 - y0) / 2.0
            direction = int(rb[q, 1] % 2)
            bonds.append((cx, cy, dx, dy, bt, direction))
            # shorten the stroke near labelled (non-carbon) atoms so that the letter stays readable
            def end(p, other, lab):
                if not lab:
                    return p
                vx, vy = other[0] - p[0], other[1] - p[1]
                nrm = max((vx * vx + vy * vy) ** 0.5, 1.0)
                return (p[0] + vx / nrm * 14, p[1] + vy / nrm * 14)
            a0 = end((x0, y0), (x1, y1), atoms[i][2] != 1)
            a1 = end((x1, y1), (x0, y0), atoms[j][2] != 1)
            vx, vy = a1[0] - a0[0], a1[1] - a0[1]
            nrm = max((vx * vx + vy * vy) ** 0.5, 1.0)
            ox, oy = -vy / nrm, vx / nrm                       # unit normal

            def line(p, q_, shift, thick=2):
                cv2.line(canvas, (int(round(p[1] + oy * shift)), int(round(p[0] + ox * shift))),
                         (int(round(q_[1] + oy * shift)), int(round(q_[0] + ox * shift))), 255, thick)
            if bt == 0:
                line(a0, a1, 0)
            elif bt == 1:
                line(a0, a1, -3)
                line(a0, a1, 3)
            elif bt == 2:
                line(a0, a1, -5)
                line(a0, a1, 0)
                line(a0, a1, 5)
            elif bt == 3:                                       # "aromatic": full line + short inner line
                line(a0, a1, 0)
                m0 = (a0[0] + vx * 0.25, a0[1] + vy * 0.25)
                m1 = (a0[0] + vx * 0.75, a0[1] + vy * 0.75)
                line(m0, m1, 5)
            else:                                               # wedge (4: solid, 5: hashed), narrow end at the tail
                tail, head = (a0, a1) if direction == 0 else (a1, a0)
                tri = np.array([[tail[1], tail[0]], [head[1] + oy * 6, head[0] + ox * 6], [head[1] - oy * 6, head[0] - ox * 6]])
                if bt == 4:
                    cv2.fillPoly(canvas, [np.round(tri).astype(np.int32)], 255)
                else:
                    for s in range(1, 8):
                        f = s / 8.0
                        px_, py_ = tail[0] + (head[0] - tail[0]) * f, tail[1] + (head[1] - tail[1]) * f
                        cv2.line(canvas, (int(round(py_ + oy * 6 * f)), int(round(px_ + ox * 6 * f))),
                                 (int(round(py_ - oy * 6 * f)), int(round(px_ - ox * 6 * f))), 255, 1)
        for (x, y, t, charge, hs) in atoms:
            if t != 1:
                txt = ATOM_LETTERS[t] + ("H" if hs else "") + {0: "", 1: "+", 2: "-"}[charge]
                (tw, th), _ = cv2.getTextSize(txt, cv2.FONT_HERSHEY_SIMPLEX, 0.6, 2)
                cv2.rectangle(canvas, (y - tw // 2 - 2, x - th // 2 - 3), (y + tw // 2 + 2, x + th // 2 + 3), 0, -1)
                cv2.putText(canvas, txt, (y - tw // 2, x + th // 2), cv2.FONT_HERSHEY_SIMPLEX, 0.6, 255, 2)
        imgs[b, 0] = (canvas > 51).astype(np.float32)           # utils_for_test.py:23: img / 255 > 0.2
        labels.append(dict(atoms=atoms, bonds=bonds))
    return imgs, labels




def label_strings(label):
    """(atoms_string, bonds_string) of one drawing in the reference generator's format: ``sym:x,y,charge,hs;`` and
    ``order:x,y,dx,dy,stereo,direction;`` (x = row, y = column, full-resolution pixels; dx, dy = half-extent vector).
    Bond classes 0..3 -> orders 1..4 without stereo; 4 / 5 -> single bond with stereo 1 / 6 (utils.py:136-141)."""
    a = "".join(f"{ATOM_LETTERS[t]}:{x},{y},{(0, 1, -1)[charge]},{hs};" for (x, y, t, charge, hs) in label["atoms"])
    b = ""
    for (x, y, dx, dy, t, direction) in label["bonds"]:
        order, stereo = (t + 1, 0) if t < 4 else (1, 1 if t == 4 else 6)
        b += f"{order}:{x},{y},{int(dx)},{int(dy)},{stereo},{direction};"
    return a, b
