"""GPU parity tests of the individual kernels, called through the C-ABI (ctypes), against CPU fp32/fp64 references.

Tolerances: operands are rounded to bf16 before the reference computes in fp64, so the only differences are
fp32 accumulation order (~1e-6 relative) and, for P8 outputs, the final bf16 rounding (2^-9 relative).
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import detrand

pytestmark = pytest.mark.gpu


def _lib():
    import abcnet_b200
    from abcnet_b200 import _lib
    _lib.require_device()
    return _lib


def to_p8(x):
    """NCHW fp32 -> P8 bf16 [N][C/8][H][W][8]"""
    N, Cc, H, W = x.shape
    return x.view(N, Cc // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def from_p8(t):
    N, P, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(N, P * 8, H, W)


def rnd(seed, shape, lo=-1.0, hi=1.0):
    return torch.from_numpy(detrand.uniform(detrand.key("k", seed), shape, lo, hi))


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def run_conv(x, w_taps, bias, taps, n_tile, act=0, out_mode=0, pool=False, out_planes_extra=0, out_plane_off=0,
             in_plane_off=0, in_planes_extra=0, out_scale=(1, 0, 1, 0), out_hw=None, want_full=True, fold=1, pair=False,
             swap=False, fold_swap=False):
    """x: NCHW fp32 (bf16-representable). w_taps: [ntaps, cout, cin]. Returns (out NCHW fp32 or None, pooled or None)."""
    L = _lib()
    from abcnet_b200.unet import _Packed
    dev = torch.device("cuda")
    N, cin, H, W = x.shape
    cout = w_taps.shape[1]
    pk = _Packed(w_taps.to(dev), bias.to(dev), taps, n_tile, cout, fold=fold, pair=pair, fold_swap=fold_swap)
    n_tile = pk.n_tile
    xin = x
    if in_planes_extra or in_plane_off:
        full = torch.full((N, cin + 8 * in_planes_extra, H, W), 7.0)
        full[:, 8 * in_plane_off:8 * in_plane_off + cin] = x
        xin = full
    src = to_p8(xin).to(dev)
    d = L.AbcConvDesc()
    d.in_, d.N, d.H, d.W = src.data_ptr(), N, H, W
    d.in_planes, d.in_plane_off, d.cin = src.shape[1], in_plane_off, cin
    d.wpack, d.bias = pk.w.data_ptr(), pk.bias.data_ptr()
    d.cout, d.n_tile, d.ntaps = cout, n_tile, len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.act, d.out_mode, d.row_fold, d.cta_pair, d.swap_mn = act, out_mode, fold, int(pair), int(swap)
    d.out_sy, d.out_oy, d.out_sx, d.out_ox = out_scale
    oH, oW = out_hw or (H, W)
    out = pooled = None
    if want_full:
        if out_mode == 0:
            out = torch.full((N, cout // 8 + out_planes_extra, oH, oW, 8), -5.0, dtype=torch.bfloat16, device=dev)
            d.out_planes, d.out_plane_off = out.shape[1], out_plane_off
        else:
            out = torch.full((N, cout, oH, oW), -5.0, dtype=torch.float32, device=dev)
        d.out = out.data_ptr()
    d.out_H, d.out_W = oH, oW
    if pool:
        pooled = torch.full((N, cout // 8, H // 2, W // 2, 8), -5.0, dtype=torch.bfloat16, device=dev)
        d.pool_out, d.pool_planes, d.pool_plane_off = pooled.data_ptr(), pooled.shape[1], 0
    L.check(L.lib.abc_conv_igemm(C.byref(d), torch.cuda.current_stream().cuda_stream), "abc_conv_igemm")
    torch.cuda.synchronize()
    o = None
    if out is not None:
        o = from_p8(out).cpu() if out_mode == 0 else out.cpu()
    return o, (from_p8(pooled).cpu() if pooled is not None else None)


def ref_conv3(x, w, b, act):
    y = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    if act == 1:
        y = F.relu(y)
    elif act == 2:
        y = F.leaky_relu(y, 0.01)
    return y.float()


def assert_close(got, ref, rtol, atol, what=""):
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = err > bound
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max err {err.max():.4g} "
                             f"(ref max {ref.abs().max():.4g}); first bad idx {idx}; got {got[bad][:8].tolist()} "
                             f"ref {ref[bad][:8].tolist()}")


TAPS3 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


def test_first_conv():
    L = _lib()
    N, H, W = 2, 40, 72
    img = (rnd(1, (N, 1, H, W), 0, 1) < 0.3).float()
    w = rnd(2, (16, 1, 3, 3))
    b = rnd(3, (16,))
    out = torch.full((N, 3, H, W, 8), -5.0, dtype=torch.bfloat16, device="cuda")
    d_img, d_w, d_b = img.cuda(), w.reshape(16, 9).contiguous().cuda(), b.cuda()      # keep the device copies alive
    L.check(L.lib.abc_conv3x3_c1(d_img.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), out.data_ptr(), N, H, W, 3, 1, 0), "c1")
    torch.cuda.synchronize()
    got = from_p8(out).cpu()
    ref = F.relu(F.conv2d(img, w, b, padding=1))
    assert_close(got[:, 8:24], ref, 2 ** -8, 1e-6, "conv3x3_c1")
    assert (got[:, :8] == -5.0).all()                      # plane 0 untouched (plane offset honoured)


def _first_conv_expected_bits(img, w, b, relu=True):
    """Bit-exact model of the kernels' arithmetic on a {0,1} image: a = bias; for tap k in row-major order
    a = fmaf(v_k, w_k, a), which for v in {0,1} is a single fp32 addition (or nothing); ReLU; round to bf16."""
    N, _, H, W = img.shape
    pad = F.pad(img, (1, 1, 1, 1)).numpy().astype(np.float32)
    wn, acc = w.reshape(16, 9).numpy().astype(np.float32), np.empty((N, 16, H, W), np.float32)
    acc[:] = b.numpy().astype(np.float32)[None, :, None, None]
    for k in range(9):
        v = pad[:, :, k // 3:k // 3 + H, k % 3:k % 3 + W]
        acc = np.where(v != 0, (acc + wn[None, :, k, None, None]).astype(np.float32), acc)
    if relu:
        acc = np.maximum(acc, np.float32(0))
    return torch.from_numpy(acc).to(torch.bfloat16)


@pytest.mark.parametrize("shape", [(2, 40, 72), (1, 17, 132), (3, 64, 516), (2, 9, 70)])
@pytest.mark.parametrize("u8", [False, True])
def test_first_conv_binary_bit_exact(shape, u8):
    """Table path (W % 4 == 0), its arithmetic fallback and the one-pixel kernel (W % 4 != 0) give the same bits on binary
    images, through both entry points (fp32 image of utils.py:80-81 and the uint8 transport format)."""
    L = _lib()
    N, H, W = shape
    img = (rnd(4, (N, 1, H, W), 0, 1) < 0.3).float()
    w, b = rnd(5, (16, 1, 3, 3)), rnd(6, (16,))
    want = _first_conv_expected_bits(img, w, b)
    d_w, d_b = w.reshape(16, 9).contiguous().cuda(), b.cuda()
    out = torch.zeros((N, 2, H, W, 8), dtype=torch.bfloat16, device="cuda")
    d_img = img.to(torch.uint8).cuda() if u8 else img.cuda()
    fn = L.lib.abc_conv3x3_c1_u8 if u8 else L.lib.abc_conv3x3_c1
    L.check(fn(d_img.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), out.data_ptr(), N, H, W, 2, 0, 0), "c1")
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 1, 4, 2, 3).reshape(N, 16, H, W)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    if not u8 and W % 4 == 0:
        # one non-binary pixel sends its warp down the arithmetic path; pixels whose 3x3 window misses it keep their bits
        img2 = img.clone()
        img2[0, 0, H // 2, W // 2] = 0.5
        d_img2 = img2.cuda()
        out2 = torch.zeros_like(out)
        L.check(fn(d_img2.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), out2.data_ptr(), N, H, W, 2, 0, 0), "c1")
        torch.cuda.synchronize()
        got2 = out2.cpu().permute(0, 1, 4, 2, 3).reshape(N, 16, H, W)
        far = torch.ones(N, 1, H, W, dtype=torch.bool)
        far[0, 0, H // 2 - 1:H // 2 + 2, W // 2 - 1:W // 2 + 2] = False
        assert torch.equal(got2.view(torch.int16)[far.expand(-1, 16, -1, -1)], want.view(torch.int16)[far.expand(-1, 16, -1, -1)])
        ref2 = F.relu(F.conv2d(img2, w, b, padding=1))
        assert_close(got2.float(), ref2, 2 ** -8, 1e-6, "conv3x3_c1 arithmetic path")


@pytest.mark.parametrize("cin,cout,n_tile", [(16, 16, 16), (64, 32, 32), (128, 64, 64)])
def test_igemm_1x1_single_tile(cin, cout, n_tile):
    """Smallest possible GEMM: one 16x8 tile, one tap, no halo -> isolates TMA box, descriptors, TMEM epilogue."""
    x = bf16_round(rnd(10 + cin, (1, cin, 16, 8)))
    w = bf16_round(rnd(20 + cin, (cout, cin)) * 0.25)
    b = rnd(30, (cout,))
    got, _ = run_conv(x, w.unsqueeze(0), b, [(0, 0)], n_tile)
    ref = (torch.einsum("nchw,oc->nohw", x.double(), w.double()) + b.double().view(1, -1, 1, 1)).float()
    assert_close(got, ref, 2 ** -7, 1e-3, f"1x1 cin={cin}")


@pytest.mark.parametrize("cin,cout,n_tile,N,H,W,act", [
    (16, 16, 16, 1, 16, 8, 0),          # single tile, halo reads only padding
    (16, 16, 16, 2, 32, 32, 1),         # many tiles, resident weights
    (32, 64, 64, 1, 48, 40, 1),
    (64, 64, 64, 2, 32, 24, 2),
    (128, 128, 128, 2, 32, 32, 1),      # weight ring (non-resident), two K chunks
    (128, 128, 64, 1, 32, 32, 1),       # two N tiles, resident weights
    (256, 256, 256, 2, 16, 16, 1),      # n_tile 256
    (512, 256, 128, 1, 16, 16, 1),      # 8 K chunks, 2 N tiles
    (64, 128, 128, 3, 6, 10, 1),        # partial tiles (H, W not multiples of the tile)
])
def test_igemm_conv3x3(cin, cout, n_tile, N, H, W, act):
    x = bf16_round(rnd(cin + H, (N, cin, H, W)))
    w = bf16_round(rnd(cout + W, (cout, cin, 3, 3)) * (2.0 / (cin * 9) ** 0.5))
    b = rnd(5, (cout,))
    wt = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3])
    got, _ = run_conv(x, wt, b, TAPS3, n_tile, act=act)
    ref = ref_conv3(x, w, b, act)
    assert_close(got, ref, 2 ** -7, 2e-3, f"conv3x3 {cin}->{cout}")


def test_igemm_pool_and_concat_slot():
    cin, cout, N, H, W = 64, 64, 2, 32, 16
    x = bf16_round(rnd(77, (N, cin, H, W)))
    w = bf16_round(rnd(78, (cout, cin, 3, 3)) * 0.05)
    b = rnd(79, (cout,))
    wt = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3])
    got, pooled = run_conv(x, wt, b, TAPS3, 64, act=1, pool=True, out_planes_extra=8, out_plane_off=3,
                           in_plane_off=2, in_planes_extra=5)
    ref = ref_conv3(x, w, b, 1)
    assert_close(got[:, 24:24 + cout], ref, 2 ** -7, 2e-3, "concat slot")
    assert (got[:, :24] == -5.0).all() and (got[:, 24 + cout:] == -5.0).all()
    assert_close(pooled, F.max_pool2d(bf16_round(ref), 2), 2 ** -7, 2e-3, "fused pool")
    # pooled-only launch (no full-resolution output)
    _, pooled2 = run_conv(x, wt, b, TAPS3, 64, act=1, pool=True, want_full=False)
    assert torch.equal(pooled2, pooled)


@pytest.mark.parametrize("cin,cout,fold,N,H,W", [
    (16, 16, 4, 2, 128, 64),           # inc1.3 / inc2.x: four rows folded into N = 64
    (16, 16, 4, 1, 40, 24),            # partial tiles in both directions (tile = 64 x 8 pixels)
    (16, 32, 2, 2, 64, 32),            # down1.0
    (32, 32, 2, 1, 96, 40),            # down1.3
    (32, 16, 4, 1, 64, 16),            # data gradient of down1.0
])
def test_igemm_row_folded_conv3x3(cin, cout, fold, N, H, W):
    """Row-folded variant (AbcConvDesc.row_fold) == the plain implicit GEMM, bit for bit (same products, same fp32
    accumulation order per output: taps in (ky, kx) order), full-resolution and fused max-pool outputs, concat slot."""
    x = bf16_round(rnd(cin + H, (N, cin, H, W)))
    w = bf16_round(rnd(cout + W, (cout, cin, 3, 3)) * (2.0 / (cin * 9) ** 0.5))
    b = rnd(5, (cout,))
    wt = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3])
    plain, plain_pool = run_conv(x, wt, b, TAPS3, cout, act=1, pool=True, out_planes_extra=3, out_plane_off=1)
    got, pooled = run_conv(x, wt, b, TAPS3, cout, act=1, pool=True, out_planes_extra=3, out_plane_off=1, fold=fold)
    ref = ref_conv3(x, w, b, 1)
    assert_close(got[:, 8:8 + cout], ref, 2 ** -7, 2e-3, f"folded conv3x3 {cin}->{cout} J={fold}")
    assert (got[:, :8] == -5.0).all() and (got[:, 8 + cout:] == -5.0).all()
    assert_close(pooled, F.max_pool2d(bf16_round(ref), 2), 2 ** -7, 2e-3, "folded fused pool")
    assert torch.equal(got, plain) and torch.equal(pooled, plain_pool)
    _, pooled2 = run_conv(x, wt, b, TAPS3, cout, act=1, pool=True, want_full=False, fold=fold)
    assert torch.equal(pooled2, pooled)


@pytest.mark.parametrize("cin,cout,n_tile,N,H,W,act", [
    (128, 128, 128, 2, 32, 32, 1),      # two tiles per stage (mt = 2), even group count
    (128, 128, 128, 3, 16, 8, 1),       # 3 groups of one tile: odd count -> the peer's last group is a dummy
    (128, 1024, 256, 2, 32, 16, 2),     # the 8-head conv1: four n-tiles, N = 256 per pair, LeakyReLU
    (256, 256, 256, 2, 16, 16, 1),
    (512, 256, 256, 1, 16, 16, 1),      # 8 K chunks
    (128, 128, 128, 1, 48, 40, 1),      # partial tiles
    (64, 64, 64, 5, 32, 24, 2),         # small layer forced into pair mode
])
def test_igemm_cta_pair_matches_single_cta(cin, cout, n_tile, N, H, W, act):
    """CTA-pair mode (cta_group::2, M = 256 per instruction, AbcConvDesc.cta_pair) == the single-CTA kernel bit for bit
    (same products, same K order), and both within tolerance of the fp64 reference."""
    x = bf16_round(rnd(cin + H, (N, cin, H, W)))
    w = bf16_round(rnd(cout + W, (cout, cin, 3, 3)) * (2.0 / (cin * 9) ** 0.5))
    b = rnd(5, (cout,))
    wt = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3])
    single, single_pool = run_conv(x, wt, b, TAPS3, n_tile, act=act, pool=True, out_planes_extra=2, out_plane_off=1)
    got, pooled = run_conv(x, wt, b, TAPS3, n_tile, act=act, pool=True, out_planes_extra=2, out_plane_off=1, pair=True)
    ref = ref_conv3(x, w, b, act)
    assert_close(got[:, 8:8 + cout], ref, 2 ** -7, 2e-3, f"pair conv3x3 {cin}->{cout}")
    assert (got[:, :8] == -5.0).all() and (got[:, 8 + cout:] == -5.0).all()
    assert torch.equal(got, single) and torch.equal(pooled, single_pool)


@pytest.mark.parametrize("cin,cout,N,H,W,act,taps", [
    (128, 128, 2, 64, 32, 1, "3x3"),      # dconv / up3.conv class: weights streamed, full 32 x 8 tiles
    (128, 128, 1, 48, 40, 2, "3x3"),      # partial tiles in y (48 = 32 + 16) and LeakyReLU
    (64, 128, 3, 32, 24, 1, "3x3"),       # down3.0 class: one K chunk
    (256, 128, 1, 32, 16, 0, "3x3"),      # up2.conv.0 class: four K chunks, no activation
    (1024, 128, 1, 32, 16, 0, "3x3"),     # data gradient of the 8-head conv1: K = 9216
    (32, 128, 2, 40, 8, 0, "1x1"),        # 1x1 data gradient of a head conv2 (weights resident, no halo), H % 32 != 0
    (384, 256, 1, 32, 24, 1, "1x1"),      # two n-tiles of 128
])
def test_igemm_operand_swap_matches_unswapped(cin, cout, N, H, W, act, taps):
    """Operand-swap mode (AbcConvDesc.swap_mn: M = 128 output channels, N = 256 pixels, transposing epilogue) == the
    unswapped kernel bit for bit (same products, same K order per output element), plane offsets honoured, and both within
    tolerance of the fp64 reference."""
    x = bf16_round(rnd(cin + H, (N, cin, H, W)))
    b = rnd(5, (cout,))
    if taps == "3x3":
        w = bf16_round(rnd(cout + W, (cout, cin, 3, 3)) * (2.0 / (cin * 9) ** 0.5))
        wt, tp = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3]), TAPS3
        ref = ref_conv3(x, w, b, act)
    else:
        w = bf16_round(rnd(cout + W, (cout, cin)) * (2.0 / cin ** 0.5))
        wt, tp = w.unsqueeze(0), [(0, 0)]
        ref = torch.einsum("nchw,oc->nohw", x.double(), w.double()) + b.double().view(1, -1, 1, 1)
        ref = (F.relu(ref) if act == 1 else ref).float()
    plain, _ = run_conv(x, wt, b, tp, 128, act=act, out_planes_extra=3, out_plane_off=2)
    got, _ = run_conv(x, wt, b, tp, 128, act=act, out_planes_extra=3, out_plane_off=2, swap=True)
    assert_close(got[:, 16:16 + cout], ref, 2 ** -7, 2e-3, f"swapped conv {cin}->{cout}")
    assert (got[:, :16] == -5.0).all() and (got[:, 16 + cout:] == -5.0).all()
    assert torch.equal(got, plain)


@pytest.mark.parametrize("cin,cout,J,N,H,W,act", [
    (64, 64, 2, 2, 128, 32, 1),          # down2.3 / inc3.0 class: two rows folded, weights streamed (3-block ring)
    (32, 64, 2, 1, 72, 40, 1),           # down2.0 class, partial tiles in y (72 = 64 + 8) and x
    (16, 32, 4, 2, 128, 24, 1),          # down1.0 class: four rows folded
    (32, 32, 4, 1, 136, 16, 0),          # 32 -> 32 (train-mode down1.3, data gradients), partial tile, no activation
    (64, 64, 2, 3, 64, 64, 2),           # LeakyReLU
])
def test_igemm_swap_with_row_fold_matches_plain(cin, cout, J, N, H, W, act):
    """Operand swap combined with row folding (GEMM rows = (folded row j, channel), J * cout = 128) == the plain implicit
    GEMM bit for bit (zero Toeplitz taps add exact zeros, same K order), plane offsets honoured."""
    x = bf16_round(rnd(cin + H, (N, cin, H, W)))
    w = bf16_round(rnd(cout + W, (cout, cin, 3, 3)) * (2.0 / (cin * 9) ** 0.5))
    b = rnd(5, (cout,))
    wt = torch.stack([w[:, :, dy + 1, dx + 1] for dy, dx in TAPS3])
    plain, _ = run_conv(x, wt, b, TAPS3, cout, act=act, out_planes_extra=3, out_plane_off=2)
    got, _ = run_conv(x, wt, b, TAPS3, cout, act=act, out_planes_extra=3, out_plane_off=2, fold=J, fold_swap=True, swap=True)
    ref = ref_conv3(x, w, b, act)
    assert_close(got[:, 16:16 + cout], ref, 2 ** -7, 2e-3, f"swap + fold conv3x3 {cin}->{cout} J={J}")
    assert (got[:, :16] == -5.0).all() and (got[:, 16 + cout:] == -5.0).all()
    assert torch.equal(got, plain)


def test_igemm_operand_swap_strided_output():
    """swap_mn with the output mapping of an up-sampling phase (pixel (y, x) -> (2y + 1, 2x), concat slot)."""
    cin, cout, N, H, W = 256, 128, 2, 32, 16
    x = bf16_round(rnd(7, (N, cin, H, W)))
    w = bf16_round(rnd(8, (cout, cin)) * (2.0 / cin ** 0.5))
    b = rnd(9, (cout,))
    kw = dict(out_planes_extra=16, out_plane_off=16, out_scale=(2, 1, 2, 0), out_hw=(2 * H, 2 * W))
    plain, _ = run_conv(x, w.unsqueeze(0), b, [(0, 0)], 128, **kw)
    got, _ = run_conv(x, w.unsqueeze(0), b, [(0, 0)], 128, swap=True, **kw)
    ref = (torch.einsum("nchw,oc->nohw", x.double(), w.double()) + b.double().view(1, -1, 1, 1)).float()
    assert_close(got[:, 128:256, 1::2, 0::2], ref, 2 ** -7, 2e-3, "swapped strided")
    assert torch.equal(got, plain)


def test_igemm_nchw_fp32_heads():
    cin, N, H, W = 128, 2, 32, 24
    x = bf16_round(rnd(81, (N, 256, H, W)))
    for cout, n_tile, off in ((1, 16, 0), (14, 16, 16), (360, 128, 0), (360, 192, 0), (60, 64, 16)):
        w = bf16_round(rnd(82 + cout, (cout, cin)) * 0.1)
        b = rnd(83, (cout,))
        xs = x[:, off * 8: off * 8 + cin]
        got, _ = run_conv(xs, w.unsqueeze(0), b, [(0, 0)], n_tile, out_mode=1, in_plane_off=off, in_planes_extra=16)
        ref = (torch.einsum("nchw,oc->nohw", xs.double(), w.double()) + b.double().view(1, -1, 1, 1)).float()
        assert_close(got, ref, 1e-4, 1e-4, f"1x1 NCHW cout={cout}")


@pytest.mark.parametrize("crop_first", [True, False])
def test_upsampling_conv_phases(crop_first):
    """ConvTranspose2d(k3, s2) + crop as 4 sub-pixel GEMMs writing a concat slot (unet.py:44-59)."""
    import abcnet_b200
    cin, cout, N, H, W = 128, 64, 2, 16, 8
    x = bf16_round(rnd(91, (N, cin, H, W)))
    w = bf16_round(rnd(92, (cin, cout, 3, 3)) * 0.05)
    b = rnd(93, (cout,))
    U = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2).float()
    ref = U[:, :, 1:, 1:] if crop_first else U[:, :, :-1, :-1]
    m = abcnet_b200.UNet(1, [1], crop_first=crop_first)
    out = torch.zeros(N, cout, 2 * H, 2 * W)
    for py in (0, 1):
        for px in (0, 1):
            ys, xs = m._phase_taps(py), m._phase_taps(px)
            taps = [(dy, dx) for (ky, dy) in ys for (kx, dx) in xs]
            wt = torch.stack([w[:, :, ky, kx].t() for (ky, dy) in ys for (kx, dx) in xs]).contiguous()
            got, _ = run_conv(x, wt, b, taps, 64, act=0, out_scale=(2, py, 2, px), out_hw=(2 * H, 2 * W))
            out[:, :, py::2, px::2] = got[:, :, py::2, px::2]
    assert_close(out, ref, 2 ** -7, 2e-3, "up-sampling conv")


@pytest.mark.parametrize("crop_first", [True, False])
@pytest.mark.parametrize("cin,cout,H,W", [(128, 64, 16, 8), (256, 128, 8, 24), (512, 256, 2, 2)])
def test_upsampling_conv_single_launch(crop_first, cin, cout, H, W):
    """The same operator as ONE launch: the four sub-pixel phases as blocks of the GEMM N axis (AbcConvDesc.subpixel), written
    into the upper half of a concat buffer whose lower half (the skip) must stay untouched. Against fp64 conv_transpose2d."""
    import abcnet_b200
    N = 2
    x = bf16_round(rnd(191, (N, cin, H, W)))
    w = bf16_round(rnd(192, (cin, cout, 3, 3)) * 0.05)
    b = rnd(193, (cout,))
    U = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2).float()
    ref = U[:, :, 1:, 1:] if crop_first else U[:, :, :-1, :-1]
    m = abcnet_b200.UNet(1, [1], crop_first=crop_first)
    dev = torch.device("cuda")
    pk = m._pack_subpixel(w.to(dev), b.to(dev))
    src = to_p8(x).to(dev)
    cat = torch.full((N, 2 * cout // 8, 2 * H, 2 * W, 8), -5.0, dtype=torch.bfloat16, device=dev)
    m._conv(pk, src, 0, cat, out_plane_off=cout // 8, act=0, out_scale=(2, 0, 2, 0), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = from_p8(cat).cpu()
    assert (got[:, :cout] == -5.0).all(), "the skip half of the concat buffer was overwritten"
    assert_close(got[:, cout:], ref, 2 ** -7, 2e-3, "single-launch up-sampling conv")
