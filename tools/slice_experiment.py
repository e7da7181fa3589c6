#!/usr/bin/env python
"""Experiment (tooling only): are the four full-resolution layers (stem, inc1.3, inc2.0, inc2.3 + pool) limited by HBM or by the
tensor pipe's operand fetch? Run them on batches small enough for every intermediate map to stay in the 126 MB L2 and compare the
time per image with the 256-image batch. If small slices are much faster per image, running the shallow layers slice by slice
(inside one CUDA graph) would pay; if not, they are not HBM-bound and slicing cannot help.
    python tools/slice_experiment.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
from abcnet_b200 import _lib  # noqa: E402
from abcnet_b200._lib import check, lib  # noqa: E402
import synthdata  # noqa: E402

dev = torch.device("cuda", 0)
m = abcnet_b200.UNet(1, list(synthdata.V2_HEADS)).to(dev).eval()
m.load_state_dict(synthdata.make_state_dict(0))
m.prepare()
P = m._packed
pool = torch.from_numpy(synthdata.binary_images(0, 8, 512, 512, 0.05)).to(dev)


def shallow(x, bufs, upto=4):
    B, _, H, W = x.shape
    st = _lib.current_stream_ptr()
    a, b, p1 = bufs
    w0, b0 = P["inc1.0"]
    check(lib.abc_conv3x3_c1(x.data_ptr(), w0.data_ptr(), b0.data_ptr(), a.data_ptr(), B, H, W, 2, 0, st))
    if upto > 1:
        m._conv(P["inc1.3"], a, 0, b, stream=st)
    if upto > 2:
        m._conv(P["inc2.0"], b, 0, a, stream=st)
    if upto > 3:
        m._conv(P["inc2.3"], a, 0, None, pool=p1, stream=st)


for B in (2, 4, 8, 16, 32, 256):
    x = pool.repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous()
    bufs = (torch.empty((B, 2, 512, 512, 8), dtype=torch.bfloat16, device=dev), torch.empty((B, 2, 512, 512, 8), dtype=torch.bfloat16, device=dev),
            torch.empty((B, 2, 256, 256, 8), dtype=torch.bfloat16, device=dev))
    reps = max(1, 256 // B)
    for _ in range(2):
        shallow(x, bufs)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            shallow(x, bufs)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"batch {B:4d} x {reps:3d} slices: {ms:7.3f} ms per 256 images (stem + inc1.3 + inc2.0 + inc2.3/pool), {ms / (B * reps) * 1e3:6.2f} us / image")
