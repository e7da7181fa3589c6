"""GPU parity of the sparse-heads inference + decode path (abcnet_b200.SparseHeadsPipeline) against
the dense path (UNet.infer + PeakDecoder): identical records bit for bit (positions, classes, omega survivors, rho bits,
counts) -- the class / offset heads evaluated only at the peaks run the same MMA sequence per output element."""
import numpy as np
import pytest
import torch

from oracle import synth, unet_ref

pytestmark = pytest.mark.gpu
HEADS = list(unet_ref.V2_HEADS)


def _model(seed):
    import abcnet_b200
    m = abcnet_b200.UNet(1, HEADS).cuda().eval()
    m.load_state_dict(unet_ref.make_state_dict(seed=seed, variant="W1"))
    return m


def _calibrate(m, x, q):
    """Constant offsets on the centre / omega heads so that a fraction (1 - q) of the pixels exceeds the -1 threshold."""
    outs = m(x)
    with torch.no_grad():
        for k in (0, 4, 7):
            m.out_modules[k].conv2.bias += -1.0 - torch.quantile(outs[k].flatten()[:2_000_000].float(), q)


@pytest.mark.parametrize("B,H,W,q,mode", [(2, 512, 512, 0.997, "nms"), (3, 256, 384, 0.99, "raw"), (2, 512, 512, 0.994, "raw")])
def test_sparse_heads_records_equal_dense_path(B, H, W, q, mode):
    import abcnet_b200
    m = _model(21)
    x = torch.from_numpy(synth.binary_images(5, B, H, W, 0.05)).cuda()
    _calibrate(m, x, q)
    dense = abcnet_b200.PeakDecoder(B, atom_cap=1024, bond_cap=8192)
    want = dense(m.infer(x, layout="p8f"), omega_mode=mode)
    pipe = abcnet_b200.SparseHeadsPipeline(m, B, peak_cap=512, bond_cap=8192)
    got = pipe.fetch(pipe.launch(x, omega_mode=mode))
    n_atoms = sum(len(a) for a, _, _ in want)
    n_bonds = sum(len(b) for _, b, _ in want)
    assert n_atoms > 5 * B and n_bonds > 0, (n_atoms, n_bonds)            # the comparison is not vacuous
    for i, ((wa, wb, wn), (ga, gb, gn)) in enumerate(zip(want, got)):
        assert wn == gn, (i, wn, gn)
        assert np.array_equal(wa, ga), f"image {i}: atom records differ"
        assert np.array_equal(wb, gb), f"image {i}: bond records differ"    # structured compare: includes the rho bits
    assert pipe.molblocks(B) == dense.molblocks(B)
    # a second batch through the same buffers (stale slots from the first one must not leak)
    x2 = torch.from_numpy(synth.binary_images(6, B, H, W, 0.04)).cuda()
    want2 = dense(m.infer(x2, layout="p8f"), omega_mode=mode)
    got2 = pipe.fetch(pipe.launch(x2, omega_mode=mode))
    for (wa, wb, wn), (ga, gb, gn) in zip(want2, got2):
        assert wn == gn and np.array_equal(wa, ga) and np.array_equal(wb, gb)


def test_sparse_heads_probability_mode_equals_dense_path():
    """The training-metric peak rule (train.py:145-151: NMS / threshold on the clamped sigmoid) through the sparse path."""
    import abcnet_b200
    m = _model(23)
    x = torch.from_numpy(synth.binary_images(8, 2, 256, 256, 0.05)).cuda()
    _calibrate(m, x, 0.99)
    dense = abcnet_b200.PeakDecoder(2, atom_cap=1024, bond_cap=8192)
    want = dense(m.infer(x, layout="p8f"), thr=0.25, apply_sigmoid=True, thr_omega=-1.0)
    pipe = abcnet_b200.SparseHeadsPipeline(m, 2, peak_cap=1024, bond_cap=8192)
    got = pipe.fetch(pipe.launch(x, thr=0.25, apply_sigmoid=True, thr_omega=-1.0))
    assert sum(len(a) for a, _, _ in want) > 10
    for (wa, wb, wn), (ga, gb, gn) in zip(want, got):
        assert wn == gn and np.array_equal(wa, ga) and np.array_equal(wb, gb)


def test_sparse_heads_capacity_overflow_raises():
    import abcnet_b200
    m = _model(22)
    x = torch.from_numpy(synth.binary_images(7, 2, 256, 256, 0.05)).cuda()
    _calibrate(m, x, 0.9)                                                 # ~10 % of the pixels above the threshold
    pipe = abcnet_b200.SparseHeadsPipeline(m, 2, peak_cap=64)
    with pytest.raises(RuntimeError, match="capacity|peak_cap"):
        pipe.fetch(pipe.launch(x))
