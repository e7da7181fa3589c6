"""GPU parity tests of the training-mode kernels (C-ABI via ctypes) against torch autograd in fp64 on bf16-rounded
operands. Tolerances: outputs stored in bf16 -> 2^-7 relative; fp32 accumulations -> 1e-3 relative to the max."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from test_kernels_gpu import TAPS3, _lib, assert_close, bf16_round, from_p8, rnd, to_p8

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


def bn_forward(z, gamma, beta, act, pool, rm=None, rv=None):
    L = _lib()
    N, Cc, H, W = z.shape
    dev = "cuda"
    zp = to_p8(z).to(dev)
    s = torch.zeros(Cc, dtype=torch.float64, device=dev)
    q = torch.zeros_like(s)
    L.check(L.lib.abc_bn_stats(zp.data_ptr(), N, H, W, Cc // 8, 0, Cc, s.data_ptr(), q.data_ptr(), _st()))
    bufs = [torch.empty(Cc, device=dev) for _ in range(4)]
    g, b = gamma.to(dev), beta.to(dev)
    L.check(L.lib.abc_bn_finalize(s.data_ptr(), q.data_ptr(), Cc, float(N * H * W), g.data_ptr(), b.data_ptr(), 1e-5, 0.1,
                                  rm.data_ptr() if rm is not None else None, rv.data_ptr() if rv is not None else None,
                                  *[t.data_ptr() for t in bufs], _st()))
    d = L.AbcBnActDesc()
    d.z, d.z_planes, d.z_plane_off = zp.data_ptr(), Cc // 8, 0
    out = torch.empty_like(zp)
    d.out, d.out_planes, d.out_plane_off = out.data_ptr(), Cc // 8, 0
    po = None
    if pool:
        po = torch.empty((N, Cc // 8, H // 2, W // 2, 8), dtype=torch.bfloat16, device=dev)
        d.pool, d.pool_planes, d.pool_plane_off = po.data_ptr(), Cc // 8, 0
    d.N, d.H, d.W, d.C = N, H, W, Cc
    d.scale, d.shift = bufs[0].data_ptr(), bufs[1].data_ptr()
    d.act, d.drop_p, d.seed = act, 0.0, 0
    L.check(L.lib.abc_bn_act(C.byref(d), _st()))
    torch.cuda.synchronize()
    return zp, bufs, from_p8(out).cpu(), (from_p8(po).cpu() if pool else None)


@pytest.mark.parametrize("act", [1, 2])
def test_bn_train_forward_and_backward(act):
    N, Cc, H, W = 3, 32, 12, 20
    z = bf16_round(rnd(1, (N, Cc, H, W)) * 2 + rnd(2, (1, Cc, 1, 1)))
    gamma, beta = rnd(3, (Cc,), 0.5, 1.5), rnd(4, (Cc,), -0.3, 0.3)
    rm = torch.zeros(Cc, device="cuda")
    rv = torch.ones(Cc, device="cuda")
    zp, bufs, out, pooled = bn_forward(z, gamma, beta, act, True, rm, rv)
    # reference (fp64)
    zz = z.double().requires_grad_(True)
    gg, bb = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = F.batch_norm(zz, None, None, gg, bb, True, 0.1, 1e-5)
    a = F.relu(y) if act == 1 else F.leaky_relu(y, 0.01)
    assert_close(out, a.detach().float(), 2 ** -7, 1e-3, "bn_act out")
    assert_close(pooled, F.max_pool2d(bf16_round(out), 2), 0, 0, "bn_act pool")
    mean, var = z.double().mean((0, 2, 3)), z.double().var((0, 2, 3), unbiased=True)
    assert_close(rm.cpu(), (0.1 * mean).float(), 1e-4, 1e-5, "running_mean")
    assert_close(rv.cpu(), (0.9 + 0.1 * var).float(), 1e-4, 1e-5, "running_var")
    # backward: loss = <dA, a> + <dP, maxpool(a)>  (pool routing evaluated on the bf16 activation, as the kernel stores it)
    dA = bf16_round(rnd(5, (N, Cc, H, W)))
    dP = bf16_round(rnd(6, (N, Cc, H // 2, W // 2)))
    a_b = a + (bf16_round(a.detach().float()).double() - a.detach())        # value = bf16(a), gradient = identity
    loss = (a * dA.double()).sum() + (F.max_pool2d(a_b, 2) * dP.double()).sum()
    loss.backward()
    L = _lib()
    d = L.AbcBnActBwdDesc()
    dev = "cuda"
    dAp, dPp = to_p8(dA).to(dev), to_p8(dP).to(dev)
    dz = torch.empty_like(zp)
    s1 = torch.zeros(Cc, dtype=torch.float64, device=dev)
    s2 = torch.zeros_like(s1)
    d.z, d.z_planes, d.z_plane_off = zp.data_ptr(), Cc // 8, 0
    d.dA, d.dA_planes, d.dA_plane_off = dAp.data_ptr(), Cc // 8, 0
    d.dP, d.dP_planes, d.dP_plane_off = dPp.data_ptr(), Cc // 8, 0
    d.dz, d.dz_planes, d.dz_plane_off = dz.data_ptr(), Cc // 8, 0
    d.N, d.H, d.W, d.C = N, H, W, Cc
    d.scale, d.shift, d.mean, d.invstd = [t.data_ptr() for t in bufs]
    d.act, d.drop_p, d.seed = act, 0.0, 0
    d.s1, d.s2 = s1.data_ptr(), s2.data_ptr()
    L.check(L.lib.abc_bn_act_backward(C.byref(d), _st()))
    torch.cuda.synchronize()
    ref = zz.grad.float()
    assert_close(from_p8(dz).cpu(), ref, 2 ** -6, 2e-3 * ref.abs().max().item(), "dz")
    assert_close(s1.cpu().float(), bb.grad.float(), 1e-3, 1e-3, "dbeta")
    assert_close(s2.cpu().float(), gg.grad.float(), 1e-3, 1e-3, "dgamma")


def run_wgrad(dz, a, taps, row_boxes=0):
    L = _lib()
    dev = "cuda"
    N, cout, H, W = dz.shape
    cin = a.shape[1]
    dzp, ap = to_p8(dz).to(dev), to_p8(a).to(dev)
    dw = torch.zeros((len(taps), cout, cin), dtype=torch.float32, device=dev)
    d = L.AbcWgradDesc()
    d.dz, d.dz_planes, d.dz_plane_off, d.cout = dzp.data_ptr(), cout // 8, 0, cout
    d.in_, d.in_planes, d.in_plane_off, d.cin = ap.data_ptr(), cin // 8, 0, cin
    d.N, d.H, d.W, d.ntaps = N, H, W, len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.dw = dw.data_ptr()
    d.row_boxes = row_boxes
    L.check(L.lib.abc_conv_wgrad(C.byref(d), _st()), "abc_conv_wgrad")
    torch.cuda.synchronize()
    return dw.cpu()


@pytest.mark.parametrize("cin,cout,N,H,W", [
    (16, 16, 2, 32, 32), (32, 64, 2, 32, 24), (128, 128, 2, 32, 32), (256, 256, 2, 16, 16), (128, 1024, 1, 32, 16),
    (512, 256, 1, 16, 16), (64, 128, 3, 6, 10), (16, 32, 2, 20, 12), (32, 32, 1, 37, 21),
])
def test_wgrad_conv3x3(cin, cout, N, H, W):
    a = bf16_round(rnd(cin, (N, cin, H, W)))
    dz = bf16_round(rnd(cout + 1, (N, cout, H, W)))
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(a.double(), w, padding=1) * dz.double()).sum().backward()
    ref = torch.stack([w.grad[:, :, dy + 1, dx + 1] for dy, dx in TAPS3]).float()
    for row_boxes in ((0, 1) if cin <= 32 else (0,)):             # the two tap-folded variants (AbcWgradDesc.row_boxes)
        got = run_wgrad(dz, a, TAPS3, row_boxes)                   # [9][cout][cin]
        assert_close(got, ref, 1e-3, 1e-3 * ref.abs().max().item(), f"wgrad {cin}->{cout} row_boxes={row_boxes}")
    if cin <= 32:                                                  # any order of the nine taps
        perm = [TAPS3[i] for i in (4, 0, 8, 2, 6, 1, 3, 5, 7)]
        got = run_wgrad(dz, a, perm, 1)
        assert_close(got, torch.stack([w.grad[:, :, dy + 1, dx + 1] for dy, dx in perm]).float(), 1e-3, 1e-3 * ref.abs().max().item(), "permuted taps")


def test_wgrad_1x1_and_padded_cout():
    a = bf16_round(rnd(7, (2, 128, 32, 24)))
    for cout in (8, 16, 64, 360):
        dz = bf16_round(rnd(8 + cout, (2, cout, 32, 24)))
        got = run_wgrad(dz, a, [(0, 0)])[0]
        ref = torch.einsum("nohw,nchw->oc", dz.double(), a.double()).float()
        assert_close(got, ref, 1e-3, 1e-3 * ref.abs().max().item(), f"wgrad 1x1 cout={cout}")


def test_first_conv_wgrad():
    L = _lib()
    N, H, W = 2, 40, 72
    img = (rnd(1, (N, 1, H, W), 0, 1) < 0.3).float()
    dz = bf16_round(rnd(2, (N, 16, H, W)))
    w = torch.zeros(16, 1, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(img.double(), w, padding=1) * dz.double()).sum().backward()
    d_img, dzp = img.cuda(), to_p8(dz).cuda()
    dw = torch.zeros(144, device="cuda")
    L.check(L.lib.abc_conv3x3_c1_wgrad(d_img.data_ptr(), 0, dzp.data_ptr(), 2, 0, N, H, W, dw.data_ptr(), _st()))
    torch.cuda.synchronize()
    ref = w.grad.reshape(144).float()
    assert_close(dw.cpu(), ref, 1e-3, 1e-3 * ref.abs().max().item(), "c1 wgrad")


@pytest.mark.parametrize("crop_first", [True, False])
def test_upsampling_conv_backward(crop_first):
    """dgrad (one K-segmented igemm over the de-interleaved phases) and wgrad (per phase) of ConvTranspose2d + crop."""
    import abcnet_b200
    from abcnet_b200 import train as T
    cin, cout, N, H, W = 128, 64, 2, 16, 8
    x = bf16_round(rnd(91, (N, cin, H, W)))
    w = bf16_round(rnd(92, (cin, cout, 3, 3)) * 0.05)
    du = bf16_round(rnd(93, (N, cout, 2 * H, 2 * W)))
    xx = x.double().requires_grad_(True)
    ww = w.double().requires_grad_(True)
    U = F.conv_transpose2d(xx, ww, None, stride=2)
    kept = U[:, :, 1:, 1:] if crop_first else U[:, :, :-1, :-1]
    (kept * du.double()).sum().backward()
    dev = "cuda"
    dx, dw = T.upconv_backward(to_p8(du).to(dev), 0, cout, to_p8(x).to(dev), w.to(dev), crop_first)
    torch.cuda.synchronize()
    assert_close(from_p8(dx).cpu(), xx.grad.float(), 2 ** -6, 2e-3 * xx.grad.abs().max().item(), "upconv dgrad")
    assert_close(dw.cpu(), ww.grad.float(), 1e-3, 1e-3 * ww.grad.abs().max().item(), "upconv wgrad")


def test_fused_loss_mode_and_scaled_p8_conversion():
    """abc_loss_partials in fused mode (sums + unscaled gradient in one pass) times head_scale equals the two-pass
    gradient bit for bit up to one fp32 rounding, and abc_nchw_to_p8_ex applies that scale, zero-pads the K planes and
    yields the conv2 bias gradient (sum over batch and pixels)."""
    from abcnet_b200.loss import ATOM_TYPE_WEIGHTS, loss_forward_backward
    from oracle import synth
    L = _lib()
    dev = "cuda"
    tg = [torch.from_numpy(t).to(dev).contiguous() for t in synth.dense_targets(7, 2, 32, 32)]
    logits = [torch.from_numpy(o).to(dev) for o in synth.random_logits(7, 2, 32, 32)]
    s = rnd(9, (10,), -0.3, 0.3).to(dev)
    tw = torch.tensor(ATOM_TYPE_WEIGHTS, device=dev)
    t1, p1, ds1, d1, hs1 = loss_forward_backward(s, tw, tg, logits, scaled=True)
    t2, p2, ds2, d2, hs2 = loss_forward_backward(s, tw, tg, logits, scaled=False)
    assert hs1 is None and hs2.shape == (8,)
    # fp64 atomics: the accumulation order differs from launch to launch -> equal to ~1e-15 relative, not bitwise
    assert abs(t1.item() - t2.item()) <= 1e-12 * abs(t1.item())
    assert_close(ds2.cpu(), ds1.cpu(), 1e-12, 1e-15, "ds")
    for i in range(8):
        assert_close((d2[i] * hs2[i]).cpu(), d1[i].cpu(), 2e-6, 1e-12, f"dlogits {i}")
    for i, g in enumerate(d2):
        N, Cc, H, W = g.shape
        planes = (Cc + 15) // 16 * 2
        out = torch.full((N, planes, H, W, 8), float("nan"), dtype=torch.bfloat16, device=dev)
        db = torch.empty(Cc, dtype=torch.float64, device=dev)
        L.check(L.lib.abc_nchw_to_p8_ex(g.data_ptr(), out.data_ptr(), N, Cc, H, W, planes, hs2[i:i + 1].data_ptr(), db.data_ptr(), _st()))
        torch.cuda.synchronize()
        ref = (g * hs2[i]).cpu()
        got = from_p8(out).cpu()
        assert torch.equal(got[:, :Cc], bf16_round(ref)), i
        assert (got[:, Cc:] == 0).all()
        assert_close(db.cpu().float(), ref.double().sum((0, 2, 3)).float(), 1e-5, 1e-9, f"dbias {i}")


def test_loss_writes_p8_gradient_operands_directly():
    """abc_loss_partials_p8: the same losses as abc_loss_partials and the unscaled gradient written directly as bf16 P8 operands
    (zero channel padding and zero padding planes over NaN-filled buffers), bit-identical to rounding the fp32 gradient of the
    fused fp32 pass, plus the per-channel sums (conv2 bias gradient / per-loss factor). train.py:95-137."""
    from abcnet_b200.loss import ATOM_TYPE_WEIGHTS, head_grad_planes, loss_forward_backward, loss_forward_p8, p8_loss_supported
    from oracle import synth
    dev = "cuda"
    for seed, B, H, W in ((7, 2, 32, 32), (8, 3, 16, 24)):
        tg = [torch.from_numpy(t).to(dev).contiguous() for t in synth.dense_targets(seed, B, H, W)]
        logits = [torch.from_numpy(o).to(dev) for o in synth.random_logits(seed, B, H, W)]
        assert p8_loss_supported(logits)
        s = rnd(9, (10,), -0.3, 0.3).to(dev)
        tw = torch.tensor(ATOM_TYPE_WEIGHTS, device=dev)
        t2, p2, ds2, d2, hs2 = loss_forward_backward(s, tw, tg, logits, scaled=False)
        dz = [torch.full((B, head_grad_planes(z.shape[1]), H, W, 8), float("nan"), dtype=torch.bfloat16, device=dev) for z in logits]
        db = [torch.full((z.shape[1],), float("nan"), dtype=torch.float64, device=dev) for z in logits]
        t3, p3, ds3, hs3 = loss_forward_p8(s, tw, tg, logits, dz, db)
        torch.cuda.synchronize()
        assert abs(t3.item() - t2.item()) <= 1e-6 * abs(t2.item())          # fp32 per-thread partial sums in a different order
        assert_close(p3.cpu(), p2.cpu(), 1e-6, 1e-12, "parts")
        assert_close(ds3.cpu(), ds2.cpu(), 1e-6, 1e-12, "ds")
        assert_close(hs3.cpu(), hs2.cpu(), 1e-6, 1e-12, "head_scale")
        nonzero = 0
        for i, g in enumerate(d2):
            Cc = g.shape[1]
            got = from_p8(dz[i]).cpu()
            assert not torch.isnan(got).any(), i
            assert torch.equal(got[:, :Cc], bf16_round(g.cpu())), i
            assert (got[:, Cc:] == 0).all(), i
            ref = g.double().sum((0, 2, 3)).cpu()
            assert_close(db[i].cpu(), ref, 1e-6, 1e-9 * max(1.0, ref.abs().max().item()), f"dbias {i}")
            nonzero += int((g != 0).sum())
        assert nonzero > 1000


def test_bn_backward_channel_factor():
    """AbcBnActBwdDesc.gscale: dz, dbeta, dgamma equal those of the same call on gscale[c] * dA (everything is linear in dA), with the
    factor applied in fp32 inside the kernel instead of being rounded into the bf16 gradient."""
    L = _lib()
    dev = "cuda"
    N, Cc, H, W = 2, 32, 12, 20
    z = bf16_round(rnd(1, (N, Cc, H, W)) * 2 + rnd(2, (1, Cc, 1, 1)))
    gamma, beta = rnd(3, (Cc,), 0.5, 1.5), rnd(4, (Cc,), -0.3, 0.3)
    zp, bufs, out, _ = bn_forward(z, gamma, beta, 2, False)
    dA = bf16_round(rnd(5, (N, Cc, H, W)))
    gs = torch.tensor([2.0 ** (i % 5 - 6) for i in range(Cc)])           # powers of two: gs * dA is exact in bf16
    res = []
    for mode in ("factor", "prescaled"):
        dAp = to_p8(dA if mode == "factor" else dA * gs.view(1, -1, 1, 1)).to(dev)
        dz = torch.empty_like(zp)
        s1 = torch.zeros(Cc, dtype=torch.float64, device=dev)
        s2 = torch.zeros_like(s1)
        gsd = gs.to(dev)
        d = L.AbcBnActBwdDesc()
        d.z, d.z_planes, d.z_plane_off = zp.data_ptr(), Cc // 8, 0
        d.dA, d.dA_planes, d.dA_plane_off = dAp.data_ptr(), Cc // 8, 0
        d.dz, d.dz_planes, d.dz_plane_off = dz.data_ptr(), Cc // 8, 0
        d.N, d.H, d.W, d.C = N, H, W, Cc
        d.scale, d.shift, d.mean, d.invstd = [t.data_ptr() for t in bufs]
        d.act, d.drop_p, d.seed = 2, 0.2, 1234
        d.s1, d.s2 = s1.data_ptr(), s2.data_ptr()
        d.gscale = gsd.data_ptr() if mode == "factor" else None
        L.check(L.lib.abc_bn_act_backward(C.byref(d), _st()))
        torch.cuda.synchronize()
        res.append((from_p8(dz).cpu(), s1.cpu(), s2.cpu()))
    (dz_a, s1_a, s2_a), (dz_b, s1_b, s2_b) = res
    assert_close(s1_a, s1_b, 1e-5, 1e-9, "dbeta")
    assert_close(s2_a, s2_b, 1e-5, 1e-9, "dgamma")
    assert_close(dz_a, dz_b, 2 ** -7, 1e-6 * dz_b.abs().max().item(), "dz")
    assert dz_b.abs().max().item() > 0


def test_stored_dropout_mask_equals_regenerated_mask():
    """AbcBnActDesc.drop_mask: the keep bits abc_bn_act stores (1 byte per P8 vector) are exactly the mask the backward pass would
    regenerate from the counter-based hash -- dz, dbeta, dgamma identical with and without them -- and they describe the zeros of
    the forward output (LeakyReLU never yields an exact zero by itself). unet.py:69."""
    L = _lib()
    dev = "cuda"
    N, Cc, H, W = 2, 32, 12, 20
    z = bf16_round(rnd(1, (N, Cc, H, W)) * 2 + rnd(2, (1, Cc, 1, 1)))
    gamma, beta = rnd(3, (Cc,), 0.5, 1.5), rnd(4, (Cc,), -0.3, 0.3)
    zp, bufs, _, _ = bn_forward(z, gamma, beta, 2, False)                # statistics / scale / shift
    seed_dev = torch.tensor([77], dtype=torch.int64, device=dev)
    mask = torch.full((N * (Cc // 8) * H * W,), 255, dtype=torch.uint8, device=dev)
    out = torch.empty_like(zp)
    f = L.AbcBnActDesc()
    f.z, f.z_planes, f.z_plane_off = zp.data_ptr(), Cc // 8, 0
    f.out, f.out_planes, f.out_plane_off = out.data_ptr(), Cc // 8, 0
    f.N, f.H, f.W, f.C = N, H, W, Cc
    f.scale, f.shift = bufs[0].data_ptr(), bufs[1].data_ptr()
    f.act, f.drop_p, f.seed = 2, 0.2, 1234
    f.seed_dev = seed_dev.data_ptr()
    f.drop_mask = mask.data_ptr()
    L.check(L.lib.abc_bn_act(C.byref(f), _st()))
    torch.cuda.synchronize()
    a = from_p8(out).cpu()                                                 # [N, C, H, W]
    bits = mask.view(N, Cc // 8, H, W).cpu()
    kept = torch.stack([(bits >> i) & 1 for i in range(8)], 2).reshape(N, Cc, H, W).bool()      # channel = plane * 8 + i
    assert torch.equal(kept, a != 0)
    assert 0.7 < kept.float().mean().item() < 0.9
    dA = bf16_round(rnd(5, (N, Cc, H, W)))
    dAp = to_p8(dA).to(dev)
    res = []
    for use_mask in (True, False):
        dz = torch.empty_like(zp)
        s1 = torch.zeros(Cc, dtype=torch.float64, device=dev)
        s2 = torch.zeros_like(s1)
        d = L.AbcBnActBwdDesc()
        d.z, d.z_planes, d.z_plane_off = zp.data_ptr(), Cc // 8, 0
        d.dA, d.dA_planes, d.dA_plane_off = dAp.data_ptr(), Cc // 8, 0
        d.dz, d.dz_planes, d.dz_plane_off = dz.data_ptr(), Cc // 8, 0
        d.N, d.H, d.W, d.C = N, H, W, Cc
        d.scale, d.shift, d.mean, d.invstd = [t.data_ptr() for t in bufs]
        d.act, d.drop_p, d.seed = 2, 0.2, 1234
        d.seed_dev = seed_dev.data_ptr()
        d.s1, d.s2 = s1.data_ptr(), s2.data_ptr()
        d.drop_mask = mask.data_ptr() if use_mask else None
        L.check(L.lib.abc_bn_act_backward(C.byref(d), _st()))
        torch.cuda.synchronize()
        res.append((dz.clone(), s1.cpu(), s2.cpu()))
    assert torch.equal(res[0][0].view(torch.int16), res[1][0].view(torch.int16))
    assert_close(res[0][1], res[1][1], 1e-12, 1e-12, "dbeta")             # fp64 atomics: order only
    assert_close(res[0][2], res[1][2], 1e-12, 1e-12, "dgamma")


def test_gather_pack_kernel():
    """abc_gather_pack: out[i] = params[code >> 22][code & 0x3FFFFF] (bf16 or fp32), 0xFFFFFFFF -> 0."""
    import ctypes as C
    from abcnet_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    params = [torch.randn(n, generator=g).cuda() for n in (5, 4096, 1 << 20, 77)]
    ptrs = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device="cuda")
    n = 8 * 4001
    pid = torch.randint(0, len(params), (n,), generator=g)
    off = (torch.rand(n, generator=g) * torch.tensor([p.numel() for p in params])[pid]).long()
    codes = (pid << 22) | off
    zero = torch.rand(n, generator=g) < 0.1
    codes[zero] = -1
    want = torch.stack([params[i].cpu()[o] for i, o in zip(pid.tolist(), off.tolist())])
    want[zero] = 0
    dcodes = codes.to(torch.int32).cuda()
    for bf16 in (1, 0):
        out = torch.full((n,), 7.0, dtype=torch.bfloat16 if bf16 else torch.float32, device="cuda")
        L.check(L.lib.abc_gather_pack(ptrs.data_ptr(), dcodes.data_ptr(), out.data_ptr(), n, bf16, torch.cuda.current_stream().cuda_stream),
                "abc_gather_pack")
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), want.to(out.dtype))


def test_fused_adam_matches_torch_adam():
    """abcnet_b200.FusedAdam == torch.optim.Adam(lr, weight_decay) (train.py:55) over several steps, including an lr change
    (train.py:84-85) and tensors smaller / larger than one chunk."""
    import abcnet_b200
    g = torch.Generator().manual_seed(4)
    shapes = [(10,), (128, 128, 3, 3), (16, 1, 3, 3), (4097,), (360, 128, 1, 1)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = abcnet_b200.FusedAdam(pa, lr=2.5e-4, weight_decay=1e-8)
    ob = torch.optim.Adam(pb, lr=2.5e-4, weight_decay=1e-8)
    for it in range(6):
        if it == 3:
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 2.5e-5
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.0 if it == 4 else 1.0)     # it 4: pure weight-decay step
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max().item()
        sa, sb = oa.state[a], ob.state[b]
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-4, atol=1e-7)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-4, atol=1e-9)
    assert float(oa.state[pa[0]]["step"]) == 6.0
