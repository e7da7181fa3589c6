"""Restatement of the reference's HOST assembly stage (test infrastructure).

The assembly is out of the accelerated path (north_star: "the existing RDKit SMILES assembly in
img2smiles*.py is unchanged"); it is restated here only so that tests can check that the records
produced by the CUDA decoder lead to the same MOL-block text (hence the same SMILES) as the
reference's own statements:

  * ``assemble``  follows ``/root/reference/src/img2smiles.py:195-318`` (bond -> atom assignment
    by anisotropic distance :195-212, pair de-duplication :219-236, valence repair :249-274,
    re-indexing :276-300, aromatic hetero-atom implicit-H list :302-314).
  * ``molblock``  follows ``/root/reference/src/generate_smiles.py:18-105`` (V2000 text).
    RDKit (``generate_smiles.py:115-118``) is not installed in this image; identical text implies
    identical SMILES.
"""
from __future__ import annotations

import numpy as np

MAX_VALENCE = {'<unkonw>': 4, 'O': 2, 'C': 4, 'N': 3, 'F': 1, 'H': 1, 'S': 6, 'Cl': 1, 'P': 5, 'Br': 1,
               'B': 3, 'I': 1, 'Si': 4, 'Se': 6, 'Te': 6, 'As': 3, 'Al': 3, 'Zn': 2, 'Ca': 2, 'Ag': 1}
_REPAIR = {2: 'O', 3: 'N', 4: 'C', 5: 'P', 6: 'S', 7: 'Cl'}


def _lrelu_half(v):
    return np.maximum(v, 0.5 * v)


def assemble(lists, aromatic_like=(4, 5, 6)):
    """lists: dict from decode_ref.records_to_lists. Returns the six arguments of sdf2smiles
    (atoms, bonds(1-based pairs), charges, bond orders, positions, implicit-H atoms) or None."""
    bp = np.array(lists["bonds_position_list"])
    bd = np.array(lists["bonds_delta_list"])
    ap_list = lists["atoms_position_list"]
    if len(bp) == 0 or len(ap_list) == 0:
        return None
    types = list(lists["atoms_type_list"])
    charges = lists["atoms_charge_list"]
    hs = lists["atoms_hs_list"]
    props = lists["bonds_property_list"]

    end_a = (bp + bd)[:, None, :]
    end_b = (bp - bd)[:, None, :]
    atoms = np.array(ap_list)[None, :, :]
    with np.errstate(invalid="ignore", divide="ignore"):
        u = bd / np.sqrt((bd ** 2).sum(-1, keepdims=True))
    v = u[:, ::-1].copy()
    v[:, 0] = -v[:, 0]
    u = u[:, None, :]
    v = v[:, None, :]
    d_a = np.abs(_lrelu_half(((end_a - atoms) * u).sum(-1))) + np.abs((2 * (end_a - atoms) * v).sum(-1))
    d_b = np.abs(_lrelu_half(-((end_b - atoms) * u).sum(-1))) + np.abs((2 * (end_b - atoms) * v).sum(-1))
    first = d_b.argmin(-1)            # img2smiles.py:211 (sic: index1 from distance2)
    second = d_a.argmin(-1)           # img2smiles.py:212

    pairs, orders = [], []
    for i in range(len(bp)):
        a, b = first[i], second[i]
        if a == b:
            continue
        if [a, b] in pairs or [b, a] in pairs:
            continue
        pairs.append([a, b])
        orders.append(props[i] + 1)   # bond_type_devocab, img2smiles.py:28

    used = set()
    for a, b in pairs:
        used.add(a)
        used.add(b)

    load = [-c for c in charges]
    for (a, b), o in zip(pairs, orders):
        n = 1 if o in aromatic_like else o
        load[a] += n
        load[b] += n
    for i, n in enumerate(load):
        if MAX_VALENCE[types[i]] < n and n in _REPAIR:
            types[i] = _REPAIR[n]

    remap, k = [], 1
    f_types, f_charges, f_pos, f_hs = [], [], [], []
    for i in range(len(ap_list)):
        remap.append(k)
        if i in used:
            f_types.append(types[i])
            f_charges.append(charges[i])
            f_pos.append(list(ap_list[i]))
            f_hs.append(hs[i])
            k += 1
    f_pairs = [[remap[a], remap[b]] for a, b in pairs]

    implicit = []
    for (a, b), o in zip(f_pairs, orders):
        if o == 4:
            for e in (a, b):
                if f_types[e - 1] != 'C' and f_hs[e - 1] != 0 and e not in implicit:
                    implicit.append(e)
    return f_types, f_pairs, f_charges, orders, f_pos, implicit


_TAIL = "0" + "  0" * 11 + "\n"


def molblock(atom_list, bond_list, charge_list, order_list, positions=None, implicit_h=()):
    """V2000 text exactly as generate_smiles.py:18-105 builds it."""
    t = "\n     RDKit\n\n"
    t += "{}{}  0  0  0  0  0  0  0  0999 V2000\n".format(str(len(atom_list)).rjust(3), str(len(bond_list)).rjust(3))
    for i, sym in enumerate(atom_list):
        sym4 = sym + " " * (4 - len(sym))
        if positions is None:
            t += "    0.0000    0.0000    0.0000 " + sym4 + _TAIL
            continue
        px = positions[i][0] / 60 - 1
        py = positions[i][1] / 60 - 1
        fx = "   {:2.4f}".format(px) if px < 0 else "    {:.4f}".format(px)
        fy = "   {:2.4f}".format(py) if py < 0 else "    {:.4f}".format(py)
        t += fx + fy + "    0.0000 " + sym4 + _TAIL
    for (a, b), o in zip(bond_list, order_list):
        o = int(o)
        if o <= 4:
            ot, st = str(o), "0"
        else:
            ot, st = "1", ("1" if o == 5 else "6")
        t += str(int(a)).rjust(3) + str(int(b)).rjust(3) + ot.rjust(3) + st.rjust(3) + "\n"
    n_chg, line = 0, ""
    for i, c in enumerate(charge_list):
        if c != 0:
            n_chg += 1
            cs = str(c)
            line += str(i + 1).rjust(4) + " " * (4 - len(cs)) + str(int(c))
    t += "M  CHG" + str(n_chg).rjust(3) + line + "\n"
    n = len(implicit_h)
    if n > 0:
        t += "M  STY  {}".format(n) + "".join("   {} DAT".format(k + 1) for k in range(n)) + "\n"
        t += "M  SLB  {}".format(n) + "".join("   {}   {}".format(k + 1, k + 1) for k in range(n)) + "\n"
        for k in range(n):
            t += "M  SAL   {}  1  {}  \n".format(k + 1, implicit_h[k])
            t += "M  SDT   {} MRV_IMPLICIT_H    \n".format(k + 1)
            t += "M  SDD   {}     0.0000    0.0000    DA    ALL  1       1    \n".format(k + 1)
            t += "M  SED   {} IMPL_H1\n".format(k + 1)
    t += "M  END\n$$$$"
    return t


def records_to_molblock(lists):
    r = assemble(lists)
    if r is None:
        return None
    return molblock(*r)



def molecule_graph(lists, with_positions=True):
    """Canonical, order-free description of the molecule that ``assemble`` builds from one image's lists, or None.

    Two record sets with the same graph give the same molecule to RDKit (``generate_smiles.py:115-118`` parses the MOL
    block and writes a canonical SMILES, which depends on the graph only): atoms with symbol / charge / position, the
    implicit-H atom set, and the bonds -- as UNORDERED atom pairs for the plain orders 1..4 (a V2000 bond line ``a b o``
    and ``b a o`` describe the same bond) and as ORDERED (begin, end) pairs for the two wedge orders 5 / 6, whose
    direction carries the stereo information. Used to decide whether an omega / omega + 30 flip of an undirected bond
    (which swaps the two end atoms of that bond, ``img2smiles.py:160-164, 195-212``) changes the molecule: it does not.
    ``with_positions=False`` drops the drawing coordinates (stride-4 pixel positions / 60 - 1 in the MOL block): the
    topology alone, which is all a SMILES string encodes for molecules without wedge bonds."""
    r = assemble(lists)
    if r is None:
        return None
    f_types, f_pairs, f_charges, orders, f_pos, implicit = r
    atoms = tuple((t, int(c)) + ((int(p[0]), int(p[1])) if with_positions else ()) for t, c, p in zip(f_types, f_charges, f_pos))
    bonds = []
    for (a, b), o in zip(f_pairs, orders):
        o = int(o)
        bonds.append((int(a), int(b), o) if o > 4 else (min(int(a), int(b)), max(int(a), int(b)), o))
    return atoms, tuple(sorted(bonds)), tuple(sorted(int(e) for e in implicit))
