"""GPU tests of the host-side paths added in round 2: the CUDA-graph inference wrapper, the graph-safe optimiser state
operations and the partial-batch fallback of TrainStep."""
import numpy as np
import pytest
import torch

from oracle import synth, unet_ref

pytestmark = pytest.mark.gpu
HEADS = list(unet_ref.V2_HEADS)


def _model(seed=3):
    import abcnet_b200
    m = abcnet_b200.UNet(1, HEADS).cuda()
    m.load_state_dict(unet_ref.make_state_dict(seed=seed, variant="W1"))
    return m


def test_infer_graph_replay_equals_eager_and_follows_weight_changes():
    import abcnet_b200
    m = _model().eval()
    B, H, W = 4, 64, 96
    x = torch.from_numpy(synth.binary_images(5, B, H, W, 0.08)).cuda()
    x8 = (x > 0).to(torch.uint8)
    with torch.no_grad():                                   # calibrate the centre heads so that peaks exist
        outs = m(x)
        for k in (0, 4, 7):
            m.out_modules[k].conv2.bias += -1.0 - torch.quantile(outs[k].flatten().float(), 0.99)
    m.invalidate_packed()
    g = abcnet_b200.InferGraph(m, B, H, W, atom_cap=4096, bond_cap=16384, dtype=torch.uint8)
    got = g(x8)
    dec = abcnet_b200.PeakDecoder(B, atom_cap=4096, bond_cap=16384)
    want = dec(m.infer(x, layout="p8f"))
    assert sum(len(a) for a, _, _ in want) > 0
    for (ga, gb, gn), (wa, wb, wn) in zip(got, want):
        assert gn == wn and np.array_equal(ga, wa) and np.array_equal(gb, wb)
    assert g.launches_per_replay >= 30
    got2 = g(x8.cpu().pin_memory())                         # pinned host input: copied asynchronously, same records
    for (ga, gb, gn), (wa, wb, wn) in zip(got2, want):
        assert gn == wn and np.array_equal(ga, wa) and np.array_equal(gb, wb)
    with torch.no_grad():                                   # weights change -> the graph is re-captured, not replayed stale
        m.out_modules[0].conv2.bias += 0.5
    got3 = g(x8)
    want3 = dec(m.infer(x, layout="p8f"))
    assert any(len(a) != len(b) for (a, _, _), (b, _, _) in zip(want, want3)), "the bias change should move the peak set"
    for (ga, gb, gn), (wa, wb, wn) in zip(got3, want3):
        assert gn == wn and np.array_equal(ga, wa) and np.array_equal(gb, wb)


def test_fused_adam_reset_and_load_state_keep_graph_pointers():
    """train.py:84-85 re-creates Adam at the LR drop; FusedAdam.reset_state does the same in place. load_state_dict copies
    into the existing moment buffers. Both must leave a captured TrainStep graph valid."""
    import abcnet_b200
    m = _model(7).train()
    m.dropout_p = 0.0
    B, H, W = 2, 64, 64
    x = torch.from_numpy(synth.binary_images(7, B, H, W, 0.08)).cuda()
    tg = [torch.from_numpy(t).cuda().contiguous() for t in synth.dense_targets(7, B, H // 4, W // 4)]
    opt = abcnet_b200.make_optimizer(m, lr=1e-3)
    step = abcnet_b200.TrainStep(m, opt, use_graph=True)
    for _ in range(3):
        step(x, tg)
    torch.cuda.synchronize()
    p0 = next(iter(opt.state))
    ptr = opt.state[p0]["exp_avg"].data_ptr()
    assert float(opt.state[p0]["step"]) == 3.0 and opt.state[p0]["exp_avg"].abs().sum() > 0
    import copy
    saved = copy.deepcopy(opt.state_dict())                  # state_dict() returns the live tensors
    opt.reset_state(lr=1e-4)
    assert float(opt.state[p0]["step"]) == 0.0 and opt.state[p0]["exp_avg"].abs().sum() == 0
    assert opt.state[p0]["exp_avg"].data_ptr() == ptr and opt.param_groups[0]["lr"] == 1e-4
    l1 = step(x, tg)                                         # the captured graph keeps working after the reset
    torch.cuda.synchronize()
    assert float(opt.state[p0]["step"]) == 1.0 and torch.isfinite(l1)
    opt.load_state_dict(saved)                               # values land in the SAME buffers
    assert opt.state[p0]["exp_avg"].data_ptr() == ptr and float(opt.state[p0]["step"]) == 3.0
    l2 = step(x, tg)
    torch.cuda.synchronize()
    assert float(opt.state[p0]["step"]) == 4.0 and torch.isfinite(l2)
    assert l1.data_ptr() != l2.data_ptr()                    # a fresh loss tensor per step


def test_train_step_partial_batch_runs_eagerly_and_graph_recovers():
    import abcnet_b200
    m = _model(9).train()
    B, H, W = 4, 64, 64
    x = torch.from_numpy(synth.binary_images(9, B, H, W, 0.08)).cuda()
    tg = [torch.from_numpy(t).cuda().contiguous() for t in synth.dense_targets(9, B, H // 4, W // 4)]
    opt = abcnet_b200.make_optimizer(m, lr=1e-3)
    step = abcnet_b200.TrainStep(m, opt, use_graph=True)
    a = step(x, tg)
    b = step(x[:3].contiguous(), [t[:3].contiguous() for t in tg])      # the DataLoader's last, partial batch (no drop_last)
    c = step(x, tg)                                                      # re-captured for the regular shape
    torch.cuda.synchronize()
    assert all(torch.isfinite(v) for v in (a, b, c)) and step.graph is not None
    assert all(torch.isfinite(p).all() for p in m.parameters())
