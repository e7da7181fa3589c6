"""GPU parity of the device-side target rasteriser (abc_rasterise_targets) against dense maps minted by the reference's own
statements (tests/golden/target_cases.npz <- utils.py:83-228 exec'd by tests/golden/make_golden.py): all eight tensors of a
batch, values and dtypes, bit-exact; and the fused losses on those targets equal the losses on the oracle's host arrays."""
import numpy as np
import pytest
import torch

from oracle import targets_ref
from test_oracle_golden import TARGET_NAMES, target_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("f64", [True, False])
def test_rasteriser_matches_reference_maps(golden_dir, f64):
    import abcnet_b200
    cases = list(target_cases(golden_dir))
    parsed = [abcnet_b200.parse_labels(a, b, *aug) for _, (a, b), aug, _ in cases]
    rast = abcnet_b200.TargetRasteriser(len(cases) + 2, f64=f64)
    for rep in range(2):                                       # second call: the maps are cleared again, nothing leaks
        order = list(range(len(cases))) if rep == 0 else list(reversed(range(len(cases))))
        maps = rast([parsed[i] for i in order])
        torch.cuda.synchronize()
        for k, name in enumerate(TARGET_NAMES):
            got = maps[k].cpu().numpy()
            for j, i in enumerate(order):
                want = cases[i][3][name]
                if not f64:
                    want = want.astype(np.float32)
                assert got[j].dtype == want.dtype and got[j].shape == want.shape, (name, got[j].shape, want.shape)
                assert np.array_equal(got[j], want), (rep, i, name)


def test_losses_on_device_targets_equal_losses_on_host_targets(golden_dir):
    import abcnet_b200
    from oracle import synth
    labels = [targets_ref.label_strings(s) for s in (11, 12)]
    host = [np.stack(x) for x in zip(*[targets_ref.rasterise(a, b) for a, b in labels])]
    dev_t = abcnet_b200.TargetRasteriser(2)([abcnet_b200.parse_labels(a, b) for a, b in labels])
    logits = [torch.from_numpy(z).cuda() for z in synth.random_logits(3, 2)]
    s = torch.zeros(10, device="cuda")
    crit = abcnet_b200.HeatmapLoss(class_weights=True)
    la = crit(logits, [torch.from_numpy(t).cuda().contiguous() for t in host], s)
    lb = crit(logits, [t.contiguous() for t in dev_t], s)
    assert abs(la.item() - lb.item()) <= 1e-12 * abs(la.item())      # identical targets; fp64 atomic sums are equal to rounding only
