#!/usr/bin/env python
"""Stand-alone timing of single conv_igemm launches (CUDA events, device-resident synthetic data).
Usage: python tools/layer_bench.py  -> prints one line per (layer, variant): ms, TFLOP/s, GB/s (algorithmic bytes).

Note (round 2): the library reads its experiment knobs (ABCNET_MT, ABCNET_NACC, ABCNET_MT256, ABCNET_PAIR_DBG) ONCE per process --
the launch path no longer calls getenv. The mt / nacc / MT256 sweeps below therefore need one process per setting: run e.g.
`ABCNET_MT=1 python tools/layer_bench.py shallow`; values set through os.environ after the first launch are ignored."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
from abcnet_b200 import _lib  # noqa: E402
from abcnet_b200.unet import _Packed  # noqa: E402

TAPS3 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


def bench(name, B, cin, cout, H, W, n_tile, taps=TAPS3, out_mode=0, pool=False, mt=None, iters=5, act=1, nacc=None, fold=1, want_full=True, pair=False):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    src = (torch.rand((B, cin // 8, H, W, 8), device=dev, generator=g) - 0.5).to(torch.bfloat16)
    w = (torch.rand((len(taps), cout, cin), device=dev, generator=g) - 0.5) * 0.05
    pk = _Packed(w, torch.zeros(cout, device=dev), taps, n_tile, cout, fold=fold, pair=pair)
    n_tile = pk.n_tile
    d = _lib.AbcConvDesc()
    d.in_, d.N, d.H, d.W = src.data_ptr(), B, H, W
    d.in_planes, d.in_plane_off, d.cin = cin // 8, 0, cin
    d.wpack, d.bias = pk.w.data_ptr(), pk.bias.data_ptr()
    d.cout, d.n_tile, d.ntaps = cout, n_tile, len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.act, d.out_mode, d.row_fold, d.cta_pair = act, out_mode, fold, int(pair)
    d.out_sy, d.out_oy, d.out_sx, d.out_ox = 1, 0, 1, 0
    d.out_H, d.out_W = H, W
    if out_mode in (0, 2):
        out = torch.empty((B, (cout + 7) // 8, H, W, 8), dtype=torch.bfloat16 if out_mode == 0 else torch.float32, device=dev)
        d.out_planes = out.shape[1]
        out_bytes = out.numel() * (2 if out_mode == 0 else 4)
    else:
        out = torch.empty((B, cout, H, W), dtype=torch.float32, device=dev)
        out_bytes = out.numel() * 4
    d.out = out.data_ptr() if want_full else None
    if not want_full:
        out_bytes = 0
    if pool:
        po = torch.empty((B, cout // 8, H // 2, W // 2, 8), dtype=torch.bfloat16, device=dev)
        d.pool_out, d.pool_planes = po.data_ptr(), po.shape[1]
        out_bytes += po.numel() * 2
    if mt is not None:
        os.environ["ABCNET_MT"] = str(mt)
    else:
        os.environ.pop("ABCNET_MT", None)
    if nacc is not None:
        os.environ["ABCNET_NACC"] = str(nacc)
    else:
        os.environ.pop("ABCNET_NACC", None)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        _lib.check(_lib.lib.abc_conv_igemm(C.byref(d), st))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        _lib.check(_lib.lib.abc_conv_igemm(C.byref(d), st))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    flops = 2.0 * B * H * W * cout * cin * len(taps)
    byts = src.numel() * 2 + out_bytes
    print(f"{name:28s} B={B:3d} {cin:4d}->{cout:4d} @{H}x{W} n_tile={n_tile:3d} mt={mt} nacc={nacc} fold={fold} pair={int(pair)}  {ms:8.3f} ms  "
          f"{flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s", flush=True)
    return ms


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "shallow"):
        for mt in (1, 2, 4, 8):
            bench("16->16@512", 64, 16, 16, 512, 512, 16, mt=mt)
        bench("16->16@512 pool", 64, 16, 16, 512, 512, 16, pool=True)
        for mt in (1, 2, 4, 8):
            bench("32->32@256", 64, 32, 32, 256, 256, 32, mt=mt)
        for mt in (1, 2, 4):
            bench("64->64@128", 256, 64, 64, 128, 128, 64, mt=mt)
    if which == "fold":                       # A/B of the row-folded variant (same box, same clocks)
        for fold in (1, 4, 1, 4):
            bench("16->16@512", 64, 16, 16, 512, 512, 16, fold=fold, iters=10)
        for fold in (1, 4):
            for mt in (1, 2):
                bench("16->16@512", 64, 16, 16, 512, 512, 16, fold=fold, mt=mt, iters=10)
        for fold in (1, 4):
            bench("16->16@512 pool only", 64, 16, 16, 512, 512, 16, pool=True, want_full=False, fold=fold, iters=10)
        for fold in (1, 2):
            bench("16->32@256", 64, 16, 32, 256, 256, 32, fold=fold, iters=10)
            bench("32->32@256 pool only", 64, 32, 32, 256, 256, 32, pool=True, want_full=False, fold=fold, iters=10)
    if which == "pair":                       # A/B of the CTA-pair (cta_group::2) variant
        for pair in (False, True, False, True):
            bench("heads 128->1024@128", 256, 128, 1024, 128, 128, 256, act=2, pair=pair)
        for pair in (False, True, False, True):
            bench("128->128@128", 256, 128, 128, 128, 128, 128, pair=pair)
        for pair in (False, True):
            bench("256->256@32", 256, 256, 256, 32, 32, 256, pair=pair)
            bench("512->512@16", 256, 512, 512, 16, 16, 256, pair=pair)
            bench("512->256@32", 256, 512, 256, 32, 32, 256, pair=pair)
            bench("256->128@64", 256, 256, 128, 64, 64, 128, pair=pair)
            bench("128->128@64", 256, 128, 128, 64, 64, 128, pair=pair)
    if which == "pairdbg":
        for dbg in ("0", "1"):
            os.environ["ABCNET_PAIR_DBG"] = dbg
            bench(f"heads pair dbg={dbg}", 256, 128, 1024, 128, 128, 256, act=2, pair=True)
            bench(f"128->128 pair dbg={dbg}", 256, 128, 128, 128, 128, 128, pair=True)
        os.environ.pop("ABCNET_PAIR_DBG")
    if which == "heads":
        for mt256 in (None, "2", None, "2"):
            if mt256:
                os.environ["ABCNET_MT256"] = mt256
            else:
                os.environ.pop("ABCNET_MT256", None)
            bench(f"heads 128->1024@128 MT256={mt256}", 256, 128, 1024, 128, 128, 256, act=2, iters=5)
        os.environ.pop("ABCNET_MT256", None)
    if which == "conv2":                      # the 1x1 head convolutions (HBM class): n_tile / mt variants, stand-alone
        one = [(0, 0)]
        for nt in (128, 192, 256, 96, 64):
            for mt in (None, 1):
                bench("head 128->360 p8f", 256, 128, 360, 128, 128, nt, taps=one, out_mode=2, act=0, mt=mt, iters=10)
        for nt in (64, 128):
            bench("head 128->60 p8f", 256, 128, 60, 128, 128, nt, taps=one, out_mode=2, act=0, iters=10)
        for cout in (14, 3, 2):
            for nt in (16, 32):
                bench(f"head 128->{cout} p8f", 256, 128, cout, 128, 128, nt, taps=one, out_mode=2, act=0, iters=10)
        for nt in (16, 32):
            bench("head 128->1 nchw", 256, 128, 1, 128, 128, nt, taps=one, out_mode=1, act=0, iters=10)
    if which == "one16f":
        bench("16->16@512", 64, 16, 16, 512, 512, 16, fold=4, iters=1)
    if which == "one16":                      # single configuration for ncu captures: layer_bench.py one16 [mt]
        bench("16->16@512", 64, 16, 16, 512, 512, 16, mt=int(sys.argv[2]) if len(sys.argv) > 2 else None, iters=1)
    if which == "one128":
        bench("128->128@128", 256, 128, 128, 128, 128, 128, iters=1)
    if which in ("sweep",):
        for mt in (2, 4, 8):
            for nacc in (2, 4, 8):
                bench("16->16@512", 64, 16, 16, 512, 512, 16, mt=mt, nacc=nacc)
        for mt in (2, 4, 8):
            for nacc in (2, 4, 8):
                bench("32->32@256", 64, 32, 32, 256, 256, 32, mt=mt, nacc=nacc)
        for mt in (1, 2, 4):
            for nacc in (2, 4, 8):
                bench("64->64@128", 256, 64, 64, 128, 128, 64, mt=mt, nacc=nacc)
        for mt in (1, 2):
            for nacc in (2, 4):
                bench("128->128@128", 256, 128, 128, 128, 128, 128, mt=mt, nacc=nacc)
    if which in ("all", "mid"):
        for nt, mt in ((128, 1), (128, 2), (64, 1), (64, 2)):
            bench("128->128@128", 256, 128, 128, 128, 128, nt, mt=mt)
        for nt, mt in ((128, 1), (128, 2), (64, 1), (256, 1)):
            bench("heads 128->1024@128", 128, 128, 1024, 128, 128, nt, mt=mt, act=2)
        bench("256->256@32", 256, 256, 256, 32, 32, 256)
        bench("256->256@32", 256, 256, 256, 32, 32, 128)
        bench("512->512@16", 256, 512, 512, 16, 16, 256)
        bench("512->256@32", 256, 512, 256, 32, 32, 256)
    if which in ("all", "heads2"):
        for cout, nt in ((1, 16), (14, 16), (360, 128), (60, 64)):
            bench(f"1x1 128->{cout} NCHW", 128, 128, cout, 128, 128, nt, taps=[(0, 0)], out_mode=1, act=0)
        for cout, nt in ((14, 16), (360, 128), (360, 192), (60, 64)):
            bench(f"1x1 128->{cout} P8F", 128, 128, cout, 128, 128, nt, taps=[(0, 0)], out_mode=2, act=0)
