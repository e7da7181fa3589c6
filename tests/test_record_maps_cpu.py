"""tools/gpu_comparator.record_level_maps -- the decode rules of img2smiles.py:62-80, :134-182 as dense tensor ops, used by
`tools/shard_infer.py --ref-check` to compare all 102 400 images of the sharded run with the fp32 torch forward -- pinned against
the oracle decode (itself pinned by the reference-minted goldens): same atom peaks and classes, same emitted (peak, omega)
pairs and bond types, on planted maps with every edge case (borders, plateau ties, value == threshold, omega bins 0 / 29 / 30 /
59, tied antipodal pairs) and on dense random logits. CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import decode_ref, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _check(maps):
    import gpu_comparator as gc
    ra, (rb, _) = decode_ref.decode_records(maps)
    outs = [torch.from_numpy(np.ascontiguousarray(m)).unsqueeze(0) for m in maps]
    atom_pk, a_cls, bond_pk, emitted, b_type = gc.record_level_maps(outs)
    assert torch.nonzero(atom_pk[0, 0]).tolist() == [[a[0], a[1]] for a in ra.tolist()]
    for a in ra.tolist():
        assert a_cls[0, :, a[0], a[1]].tolist() == a[2:5]
    assert torch.nonzero(emitted[0].permute(1, 2, 0)).tolist() == [[b[0], b[1], b[2]] for b in rb.tolist()]
    for b in rb.tolist():
        assert int(b_type[0, b[2], b[0], b[1]]) == b[3]
    return len(ra), len(rb)


@pytest.mark.parametrize("seed", range(6))
def test_record_level_maps_match_oracle_decode_on_planted_maps(seed):
    na, nb = _check(synth.planted_logits(seed)[0])
    assert na > 10 and nb > 10


def test_record_level_maps_match_oracle_decode_on_dense_random_logits():
    r = synth.random_logits(3, 1, 64, 64)
    na, nb = _check([x[0] for x in r])
    assert na > 100 and nb > 1000
