"""CPU test of the host half of the GPU target rasteriser (abcnet_b200.parse_labels) against the
oracle's parse, which tests/test_oracle_golden.py pins to the reference's own statements (utils.py:83-228)."""
import numpy as np
import pytest

from abcnet_b200.targets import parse_labels
from oracle import targets_ref


@pytest.mark.parametrize("aug", [(1, 1, 0, 0), (0.87, 1, 33, 0), (1, 0.93, 0, 17)])
def test_parse_labels_matches_oracle(aug):
    for seed in range(6):
        a, b = targets_ref.label_strings(seed)
        atoms, bonds, rho = parse_labels(a, b, *aug)
        ra = targets_ref.parse_atoms(a, *aug)
        rb = targets_ref.parse_bonds(b, *aug)
        assert atoms.tolist() == [list(t) for t in ra]
        assert [(x, y, t, bins[:n]) for (x, y, t, n, b0, b1), bins in zip(bonds.tolist(), [[r[4], r[5]] for r in bonds.tolist()])] == \
               [(x, y, t, bins) for (x, y, t, bins, _) in rb]
        assert np.array_equal(rho, np.array([r[4] for r in rb], np.float64))           # same float64 bits


def test_parse_labels_rejects_out_of_range():
    with pytest.raises(ValueError):
        parse_labels("C:600,10,0,0;", "")
    with pytest.raises(ValueError):
        parse_labels("", "1:10,700,5,5,0,0;")
