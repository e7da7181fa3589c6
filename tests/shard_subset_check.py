#!/usr/bin/env python
"""CPU-oracle check of the subset dumped by `tools/shard_infer.py --dump-subset M` (BASELINE configs[2]: "SMILES exact-match vs
reference" on a fixed subset of the sharded 100 k-image run). Needs no GPU: run in the build container after the gpurun call.

    python tests/shard_subset_check.py gpurun_out/shard_subset.pt [--out profiles/r02_shard_subset_check.json]

For global images [0, M) of the run it regenerates the images (synthdata.molecules, deterministic), runs the fp32 CPU oracle
(oracle.unet_ref + decode_ref + assemble_ref = src/unet.py + img2smiles.py:62-318 + generate_smiles.py:18-105) with the dumped
weights and compares the MOL-block texts with the ones the sharded GPU run produced: identical text, identical molecular graph,
identical topology (see assemble_ref.molecule_graph). Test infrastructure (not collected by pytest: no test_ prefix)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assemble_ref, decode_ref, unet_ref  # noqa: E402
from synthdata import molecules  # noqa: E402


def parse_molblock(text):
    """V2000 text -> (atoms, bonds, implicit-H set) in the canonical form of assemble_ref.molecule_graph (with positions)."""
    lines = text.split("\n")
    na, nb = int(lines[3][0:3]), int(lines[3][3:6])
    atoms = []
    for ln in lines[4:4 + na]:
        f = ln.split()
        atoms.append([f[3], 0, int(round((float(f[0]) + 1) * 60)), int(round((float(f[1]) + 1) * 60))])
    bonds = []
    for ln in lines[4 + na:4 + na + nb]:
        a, b, o, st = int(ln[0:3]), int(ln[3:6]), int(ln[6:9]), int(ln[9:12])
        if st:
            bonds.append((a, b, 5 if st == 1 else 6))
        else:
            bonds.append((min(a, b), max(a, b), o))
    implicit = []
    for ln in lines[4 + na + nb:]:
        if ln.startswith("M  CHG"):
            f = ln.split()
            for i in range(int(f[2])):
                atoms[int(f[3 + 2 * i]) - 1][1] = int(f[4 + 2 * i])
        if ln.startswith("M  SAL"):
            implicit.append(int(ln.split()[4]))
    return tuple(tuple(a) for a in atoms), tuple(sorted(bonds)), tuple(sorted(implicit))


THR = -1.0


def _centre_need(z, x, y, present_in_ref):
    """Smallest per-comparison logit perturbation that turns the reference's decision about pixel (x, y) of a centre map
    (img2smiles.py:62-68) into the opposite one: a present peak disappears when ANY of its conditions fails (min margin); an
    absent one appears only when ALL currently failing conditions flip (max of their margins)."""
    H, W = z.shape
    m = [z[x, y] - THR] + [z[x, y] - z[i, j] for i in range(max(x - 1, 0), min(x + 2, H)) for j in range(max(y - 1, 0), min(y + 2, W))
                           if (i, j) != (x, y)]
    if present_in_ref:
        return float(min(abs(v) for v in m))
    failing = [abs(v) for k, v in enumerate(m) if (v <= 0 if k == 0 else v < 0)]
    return float(max(failing)) if failing else 0.0


def _omega_need(col, w, present_in_ref):
    n = len(col)
    h = n // 2
    conds = [(col[w] - THR, True), (col[w] - col[(w - 1) % n], False), (col[w] - col[(w + 1) % n], False)]
    if w <= h - 2:
        others, strict = (w + h - 1, w + h), False
    elif w == h - 1:
        others, strict = (n - 2, 0), False
    elif w == h:
        others, strict = (0, n - 1), True
    else:
        others, strict = (w - h - 1, w - h), True
    conds += [(col[w] - col[k], strict) for k in others]
    if present_in_ref:
        return float(min(abs(v) for v, _ in conds))
    failing = [abs(v) for v, st in conds if (v <= 0 if st else v < 0)]
    return float(max(failing)) if failing else 0.0


def record_differences(R, ra, rb, atoms, bonds):
    """Per differing decision between the reference records (ra, rb from logits R) and the product's records: the kind and the
    smallest logit perturbation that explains it (needed_error), computed on the REFERENCE logits."""
    out = []
    ga = {(int(a["x"]), int(a["y"])): (int(a["type"]), int(a["charge"]), int(a["hs"])) for a in atoms}
    A_r = {(a[0], a[1]): (a[2], a[3], a[4]) for a in ra.tolist()}
    for pos in sorted(set(A_r) ^ set(ga)):
        out.append(dict(kind="atom peak", pos=pos, in_ref=pos in A_r, needed_error=_centre_need(R[0][0], pos[0], pos[1], pos in A_r)))
    for pos in sorted(set(A_r) & set(ga)):
        for k, (name, head) in enumerate((("atom type", 1), ("atom charge", 2), ("atom hs", 3))):
            if A_r[pos][k] != ga[pos][k]:
                v = R[head][:, pos[0], pos[1]]
                out.append(dict(kind=name, pos=pos, needed_error=float(abs(v[A_r[pos][k]] - v[ga[pos][k]]))))
    gb = {(int(b["x"]), int(b["y"]), int(b["omega"])): int(b["type"]) for b in bonds}
    B_r = {(b[0], b[1], b[2]): b[3] for b in rb.tolist()}
    P_r = decode_ref._peaks2d(R[4][0], THR)
    P_o = {(k[0], k[1]) for k in gb}
    nw = R[7].shape[0]
    for key in sorted(set(B_r) ^ set(gb)):
        x, y, w = key
        in_ref = key in B_r
        if not in_ref and not P_r[x, y]:               # a bond-centre peak the reference does not have
            out.append(dict(kind="bond peak", pos=(x, y), omega=w, in_ref=False, needed_error=_centre_need(R[4][0], x, y, False)))
        elif in_ref and (x, y) not in P_o:             # no record at all at this pixel: the centre peak or this omega bin went away
            out.append(dict(kind="bond peak / omega", pos=(x, y), omega=w, in_ref=True,
                            needed_error=min(_centre_need(R[4][0], x, y, True), _omega_need(R[7][:, x, y], w, True))))
        else:
            out.append(dict(kind="bond omega", pos=(x, y), omega=w, in_ref=in_ref, needed_error=_omega_need(R[7][:, x, y], w, in_ref)))
    for key in sorted(set(B_r) & set(gb)):
        if B_r[key] != gb[key]:
            x, y, w = key
            v = R[5].reshape(-1, nw, *R[5].shape[1:])[:, w, x, y]
            out.append(dict(kind="bond type", pos=(x, y), omega=w, needed_error=float(abs(v[B_r[key]] - v[gb[key]]))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--out", default="")
    ap.add_argument("--limit", type=int, default=0)
    args = ap.parse_args()
    d = torch.load(args.dump, map_location="cpu")
    sd, got, P = d["state_dict"], d["molblocks"], d["pool"]
    M = len(got) if not args.limit else min(args.limit, len(got))
    assert d.get("trained"), "the subset check is meaningful for trained weights only"
    imgs, _ = molecules.pseudo_molecules(d["seed"], P, 512, 512)
    torch.set_num_threads(os.cpu_count() or 1)
    rep = dict(images=M, identical_text=0, identical_graph=0, identical_topology=0, both_none=0, differing=[], decisions=[],
               ref_atom_peaks=0, ref_bond_records=0, identical_records=0)
    have_records = "atoms" in d
    for c0 in range(0, M, 16):
        g = np.arange(c0, min(c0 + 16, M))
        x = np.stack([np.roll(imgs[i % P], (i // P) % 512, axis=-1) for i in g])
        with torch.no_grad():
            ref = [o.numpy() for o in unet_ref.forward(torch.from_numpy(x), sd)]
        for jj, i in enumerate(g):
            R = [r[jj] for r in ref]
            ra, (rb, rrho) = decode_ref.decode_records(R, -1.0, "nms")
            rep["ref_atom_peaks"] += len(ra)
            rep["ref_bond_records"] += len(rb)
            if have_records:
                diffs = record_differences(R, ra, rb, d["atoms"][i], d["bonds"][i])
                rep["identical_records"] += int(not diffs)
                for it in diffs:
                    it["image"] = int(i)
                rep["decisions"] += diffs
            L = decode_ref.records_to_lists(ra, (rb, rrho)) if (len(ra) and len(rb)) else None
            want = assemble_ref.records_to_molblock(L) if L is not None else None
            if want == got[i]:
                rep["identical_text"] += 1
                rep["identical_graph"] += 1
                rep["identical_topology"] += 1
                rep["both_none"] += int(want is None)
                continue
            gw = assemble_ref.molecule_graph(L) if L is not None else None
            gg = parse_molblock(got[i]) if got[i] is not None else None
            same_graph = gw == gg
            strip = lambda gr: None if gr is None else (tuple(a[:2] for a in gr[0]), gr[1], gr[2])  # noqa: E731
            same_topo = strip(gw) == strip(gg)
            rep["identical_graph"] += int(same_graph)
            rep["identical_topology"] += int(same_topo)
            rep["differing"].append(dict(image=int(i), graph_identical=bool(same_graph), topology_identical=bool(same_topo)))
    if have_records:
        need = sorted(it["needed_error"] for it in rep["decisions"])
        kinds = {}
        for it in rep["decisions"]:
            kinds[it["kind"]] = kinds.get(it["kind"], 0) + 1
        rep["decision_summary"] = dict(differing_decisions=len(need), by_kind=kinds, needed_error_max=need[-1] if need else 0.0,
                                       needed_error_p50=need[len(need) // 2] if need else 0.0,
                                       needed_error_p95=need[int(len(need) * 0.95)] if need else 0.0,
                                       logit_scale={str(k): float(np.abs(ref[k]).max()) for k in (0, 4, 7)})
    print(json.dumps({k: v for k, v in rep.items() if k not in ("differing", "decisions")}), f"({len(rep['differing'])} images listed)")
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
