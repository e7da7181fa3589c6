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
    rep = dict(images=M, identical_text=0, identical_graph=0, identical_topology=0, both_none=0, differing=[])
    for c0 in range(0, M, 16):
        g = np.arange(c0, min(c0 + 16, M))
        x = np.stack([np.roll(imgs[i % P], (i // P) % 512, axis=-1) for i in g])
        with torch.no_grad():
            ref = [o.numpy() for o in unet_ref.forward(torch.from_numpy(x), sd)]
        for jj, i in enumerate(g):
            ra, (rb, rrho) = decode_ref.decode_records([r[jj] for r in ref], -1.0, "nms")
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
    print(json.dumps({k: v for k, v in rep.items() if k != "differing"}), f"({len(rep['differing'])} images listed)")
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
