"""CPU tests of the native host assembler (abc_assemble_molblocks): records -> MOL-block text.
Pinned against (i) the MOL blocks minted by the reference's own statements (tests/golden/decode_cases.json, made by
tests/golden/make_golden.py from img2smiles.py + generate_smiles.py) and (ii) the oracle restatement on randomised records,
including the degenerate cases (rho = 0 -> NaN distances, duplicate atom peaks, over-valent atoms, aromatic hetero atoms)."""
import json
import os

import numpy as np
import pytest

from abcnet_b200.decode import ATOM_DT, BOND_DT, assemble_molblocks, records_to_lists
from oracle import assemble_ref, decode_ref, synth


def _pack(recs, atom_cap=None, bond_cap=None):
    """recs: list of (atoms int [na, 5], (bonds int [nb, 4], rho float32 [nb]), n_bond_peaks) -> uint8 arrays + counts."""
    N = len(recs)
    atom_cap = atom_cap or max(1, max(len(a) for a, _, _ in recs))
    bond_cap = bond_cap or max(1, max(len(b[0]) for _, b, _ in recs))
    A = np.zeros((N, atom_cap), ATOM_DT)
    B = np.zeros((N, bond_cap), BOND_DT)
    C = np.zeros((N, 4), np.int32)
    for i, (a, (bi, br), nbp) in enumerate(recs):
        a, bi = np.asarray(a, np.int64).reshape(-1, 5), np.asarray(bi, np.int64).reshape(-1, 4)
        for k, f in enumerate(("x", "y", "type", "charge", "hs")):
            A[f][i, :len(a)] = a[:, k]
        for k, f in enumerate(("x", "y", "omega", "type")):
            B[f][i, :len(bi)] = bi[:, k]
        B["rho"][i, :len(bi)] = np.asarray(br, np.float32)
        C[i] = (len(a), len(bi), nbp, 0)
    return A.view(np.uint8).reshape(N, atom_cap, 8), B.view(np.uint8).reshape(N, bond_cap, 12), C, A, B


def _python_path(A, B, C):
    out = []
    for i in range(len(C)):
        L = records_to_lists(A[i, :C[i, 0]], B[i, :C[i, 1]], int(C[i, 2]))
        out.append(None if L is None else assemble_ref.records_to_molblock(L))
    return out


@pytest.mark.parametrize("mode,suffix", [("nms", ""), ("raw", "_raw")])
def test_native_assembler_matches_reference_molblocks(golden_dir, mode, suffix):
    g = json.load(open(os.path.join(golden_dir, "decode_cases.json")))
    recs = []
    for seed in range(4):
        atoms, bonds = decode_ref.decode_records(synth.planted_logits(seed)[0], -1.0, mode)
        recs.append((atoms, bonds, max(1, len(bonds[0]))))
    a8, b8, C, _, _ = _pack(recs)
    for n_threads in (1, 0):
        got = assemble_molblocks(a8, b8, C, n_threads=n_threads)
        for seed in range(4):
            assert got[seed] == g[f"planted{seed}{suffix}"]["molblock"], seed


def test_native_assembler_matches_python_path_on_random_records():
    rng = np.random.default_rng(0)
    recs = []
    for i in range(300):
        na, nb = int(rng.integers(0, 40)), int(rng.integers(0, 60))
        pos = rng.integers(0, 128, size=(na, 2))
        if na > 3 and i % 3 == 0:
            pos[1] = pos[0] + (1, 0)                       # duplicate peak (< 2 px): dropped by the greedy de-duplication
        atoms = np.concatenate([pos, rng.integers(0, 14, (na, 1)), rng.integers(0, 3, (na, 1)), rng.integers(0, 2, (na, 1))], 1)
        if na >= 2 and nb:
            # bonds between random atom pairs: centre = midpoint, rho = half length, omega = direction bin -> realistic geometry
            i0, i1 = rng.integers(0, na, nb), rng.integers(0, na, nb)
            p0, p1 = pos[i0].astype(np.float64), pos[i1].astype(np.float64)
            mid = np.rint((p0 + p1) / 2).astype(np.int64)
            d = (p0 - p1) / 2
            ang = np.arctan2(d[:, 1], d[:, 0])
            ang = np.where(ang < -np.pi / 2, ang + np.pi, np.where(ang >= np.pi / 2, ang - np.pi, ang))
            w = np.clip(np.floor((ang + np.pi / 2) / (np.pi / 30)), 0, 59).astype(np.int64)
            w = np.where(rng.random(nb) < 0.3, (w + 30) % 60, w)
            rho = np.hypot(d[:, 0], d[:, 1]).astype(np.float32)
            rho[rng.random(nb) < 0.05] = 0.0                # degenerate: NaN distances -> both ends pick atom 0
            bi = np.concatenate([mid, w[:, None], rng.integers(0, 6, (nb, 1))], 1)
        else:
            bi = np.concatenate([rng.integers(0, 128, (nb, 2)), rng.integers(0, 60, (nb, 1)), rng.integers(0, 6, (nb, 1))], 1)
            rho = (rng.random(nb) * 20).astype(np.float32)
        recs.append((atoms, (bi, rho), 0 if i % 17 == 0 else max(nb, 1)))
    a8, b8, C, A, B = _pack(recs, atom_cap=48, bond_cap=64)
    want = _python_path(A, B, C)
    got = assemble_molblocks(a8, b8, C)
    assert sum(w is not None for w in want) > 200 and any(w is None for w in want)
    assert sum("M  STY" in w for w in want if w) > 5                      # implicit-H blocks exercised
    for i, (g_, w_) in enumerate(zip(got, want)):
        assert g_ == w_, f"image {i}"


def test_native_assembler_rejects_bad_input():
    a8, b8, C, _, _ = _pack([(np.zeros((2, 5)), (np.zeros((1, 4)), np.ones(1)), 1)])
    C[0, 0] = 99
    with pytest.raises(RuntimeError, match="exceed"):
        assemble_molblocks(a8, b8, C)
    with pytest.raises(ValueError):
        assemble_molblocks(a8.astype(np.int16), b8, C)
