"""CPU tests of the gather-map derivation behind abcnet_b200.train.PackArena: running the packing
code on element codes and gathering must reproduce the directly packed bf16 blocks / fp32 biases bit for bit, for every
layout variant the training pass uses (plain 3x3, row-folded, zero-padded 1x1, transposed data-gradient, K-segmented)."""
import torch

from abcnet_b200.train import TAPS3, Packed, phase_taps
from abcnet_b200.unet import row_fold_for, swap_fold_for


def _codes(t, pid):
    return (torch.arange(t.numel(), dtype=torch.float64) + float((pid << 22) + 1)).view(t.shape)


def _gather(code_t, params, dtype):
    c = code_t.reshape(-1).to(torch.int64) - 1
    out = torch.zeros(c.numel(), dtype=torch.float32)
    for pid, p in enumerate(params):
        sel = (c >= 0) & ((c >> 22) == pid)
        out[sel] = p.reshape(-1)[c[sel] & ((1 << 22) - 1)]
    return out.to(dtype).view(code_t.shape)


def _check(make, params):
    direct = make([p.float() for p in params])
    coded = make([_codes(p, i) for i, p in enumerate(params)])
    assert coded.w.dtype == torch.float64 and direct.w.dtype == torch.bfloat16
    assert torch.equal(_gather(coded.w, params, torch.bfloat16).view(torch.int16), direct.w.view(torch.int16))
    assert torch.equal(_gather(coded.bias, params, torch.float32), direct.bias)
    assert (coded.n_tile, coded.cout, coded.cin, coded.fold) == (direct.n_tile, direct.cout, direct.cin, direct.fold)


def test_plain_and_folded_3x3():
    g = torch.Generator().manual_seed(0)
    for cin, cout in ((16, 16), (32, 32), (16, 32), (32, 64), (64, 64), (64, 128), (128, 128)):
        w, b = torch.randn(cout, cin, 3, 3, generator=g), torch.randn(cout, generator=g)

        def make(ps, cin=cin, cout=cout):
            wt, bias = ps
            return Packed(torch.stack([wt[:, :, dy + 1, dx + 1] for dy, dx in TAPS3]), bias, TAPS3, fold=row_fold_for(cin, cout))
        _check(make, [w, b])

        def make_s(ps, cin=cin, cout=cout):                       # row folding in the operand-swap order (32 / 64 channels)
            wt, bias = ps
            js = swap_fold_for(cin, cout)
            return Packed(torch.stack([wt[:, :, dy + 1, dx + 1] for dy, dx in TAPS3]), bias, TAPS3, fold=js or 1, fold_swap=bool(js))
        _check(make_s, [w, b])

        def make_d(ps, cin=cin, cout=cout):                       # data gradient: transposed, flipped taps, zero bias
            wt = ps[0]
            mats = torch.stack([wt[:, :, dy + 1, dx + 1].t() for dy, dx in TAPS3]).contiguous()
            return Packed(mats, wt.new_zeros(cin), [(-dy, -dx) for dy, dx in TAPS3], fold=row_fold_for(cout, cin))
        _check(make_d, [w, b])


def test_padded_1x1_heads_and_concatenated_conv1():
    g = torch.Generator().manual_seed(1)
    for h in (1, 14, 360, 60):
        w, b = torch.randn(h, 128, 1, 1, generator=g), torch.randn(h, generator=g)
        n_tile = 16 if h <= 16 else (64 if h <= 64 else 128)
        _check(lambda ps, h=h, n_tile=n_tile: Packed(ps[0].reshape(h, -1).unsqueeze(0).contiguous(), ps[1], [(0, 0)], n_tile=n_tile), [w, b])
        c16 = (h + 15) // 16 * 16 if h <= 64 else (h + 63) // 64 * 64

        def make_d(ps, h=h, c16=c16):
            w2 = ps[0].reshape(h, 128)
            w2p = torch.cat([w2, w2.new_zeros(c16 - h, 128)], 0)
            return Packed(w2p.t().contiguous().unsqueeze(0), w2.new_zeros(128), [(0, 0)], n_tile=128)
        _check(make_d, [w, b])
    ws = [torch.randn(128, 128, 3, 3, generator=g) for _ in range(4)]
    bs = [torch.randn(128, generator=g) for _ in range(4)]

    def make_h1(ps):
        w1, b1 = torch.cat(ps[:4], 0), torch.cat(ps[4:])
        return Packed(torch.stack([w1[:, :, dy + 1, dx + 1] for dy, dx in TAPS3]), b1, TAPS3, n_tile=256)
    _check(make_h1, ws + bs)


def test_upsampling_phases_and_segmented_backward():
    g = torch.Generator().manual_seed(2)
    cin, cout = 128, 64
    w, b = torch.randn(cin, cout, 3, 3, generator=g), torch.randn(cout, generator=g)
    for crop_first in (True, False):
        for py in (0, 1):
            for px in (0, 1):
                ys, xs = phase_taps(py, crop_first), phase_taps(px, crop_first)
                taps = [(dy, dx) for (ky, dy) in ys for (kx, dx) in xs]
                _check(lambda ps, ys=ys, xs=xs, taps=taps: Packed(
                    torch.stack([ps[0][:, :, ky, kx].t() for (ky, dy) in ys for (kx, dx) in xs]).contiguous(), ps[1], taps), [w, b])
        taps, segments, sel = [], [], []
        for py in (0, 1):
            for px in (0, 1):
                t0 = len(taps)
                for (ky, dy) in phase_taps(py, crop_first):
                    for (kx, dx) in phase_taps(px, crop_first):
                        taps.append((-dy, -dx))
                        sel.append((ky, kx))
                segments.append((t0, len(taps) - t0))
        _check(lambda ps, sel=sel, taps=taps, segments=segments: Packed(
            torch.stack([ps[0][:, :, ky, kx] for ky, kx in sel]).contiguous(), ps[0].new_zeros(cin), taps, segments=segments), [w, b])
