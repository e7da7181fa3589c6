"""Thin Python caller of the whole-network C entry points (``abc_unet_create`` / ``abc_unet_forward_infer``).

``abcnet_b200.UNet`` keeps the launch plan in Python because it also serves training, timing hooks and experiments; a host
without Python (the reference's maintainers may serve ``/root/reference/src/unet.py`` from C++, Go, Java ...) gets the same
forward pass from ``include/abcnet_b200.h`` alone. ``NativeUNet`` is that path driven from Python: it hands the raw fp32
``state_dict`` to the library (BatchNorm fold + packing happen in C++ on the host), owns the device memory the library asks for,
and returns the same list of logits as ``UNet.forward`` -- bit-identical (``tests/test_native_gpu.py``). No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AbcNamedTensor, AbcUNetConfig, check, lib
from .unet import HeadMaps


class NativeUNet:
    """net = NativeUNet(state_dict, in_channels=1, heads=[1, 14, 3, 2, 1, 360, 60, 60]); outs = net(x)  # x: CUDA [B,C,H,W]"""

    def __init__(self, state_dict, in_channels=1, heads=(1, 14, 3, 2, 1, 360, 60, 60), crop_first=True, device=None, act_dtype="bf16"):
        _lib.require_device()
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.heads = list(heads)
        cfg = self.cfg = AbcUNetConfig()
        if act_dtype not in ("bf16", "fp16"):
            raise ValueError("act_dtype must be 'bf16' or 'fp16'")
        cfg.in_channels, cfg.n_heads, cfg.crop_first = int(in_channels), len(self.heads), int(bool(crop_first))
        cfg.act_fp16 = int(act_dtype == "fp16")
        for i, h in enumerate(self.heads):
            cfg.heads[i] = int(h)
        n = lib.abc_unet_wpack_bytes(C.byref(cfg))
        if n < 0:
            check(-1, "abc_unet_wpack_bytes")
        self.wpack = torch.empty(n, dtype=torch.uint8, device=self.dev)
        # host copies of the fp32 tensors (the library reads host pointers: this is what a checkpoint reader hands over)
        keep, arr = [], []
        for k, v in state_dict.items():
            if not torch.is_floating_point(v):
                continue                                     # num_batches_tracked
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append((k.encode(), t))
        arr = (AbcNamedTensor * len(keep))()
        for i, (k, t) in enumerate(keep):
            arr[i].name, arr[i].data, arr[i].numel = k, t.data_ptr(), t.numel()
        handle = C.c_void_p()
        with torch.cuda.device(self.dev):
            check(lib.abc_unet_create(C.byref(cfg), arr, len(keep), self.wpack.data_ptr(), n, _lib.current_stream_ptr(), C.byref(handle)),
                  "abc_unet_create")
        self._h = handle
        self._ws = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib.abc_unet_destroy(h)
            except Exception:                                # interpreter shutdown: the library may already be gone
                pass
            self._h = None

    @torch.no_grad()
    def __call__(self, x, layout="nchw", outs=None):
        if not x.is_cuda:
            raise RuntimeError("abcnet_b200.NativeUNet needs a CUDA tensor (no CPU fallback)")
        if layout not in ("nchw", "p8f"):
            raise ValueError("layout must be 'nchw' or 'p8f'")
        u8 = x.dtype in (torch.uint8, torch.bool)
        x = x.contiguous().view(torch.uint8) if u8 else x.contiguous().float()
        B, _, H, W = x.shape
        need = lib.abc_unet_workspace_bytes(C.byref(self.cfg), B, H, W)
        if need < 0:
            check(-1, "abc_unet_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        p8f = layout == "p8f"
        if outs is None:
            outs = HeadMaps(self.heads)
            for h in self.heads:
                shape = (B, (h + 7) // 8, H // 4, W // 4, 8) if (p8f and h > 1) else (B, h, H // 4, W // 4)
                outs.append(torch.empty(shape, dtype=torch.float32, device=x.device))
        ptrs = (C.c_void_p * len(self.heads))(*[o.data_ptr() for o in outs])
        check(lib.abc_unet_forward_infer(self._h, x.data_ptr(), int(u8), B, H, W, self._ws.data_ptr(), self._ws.numel(), ptrs,
                                         2 if p8f else 1, _lib.current_stream_ptr()), "abc_unet_forward_infer")
        return outs
