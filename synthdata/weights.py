"""Deterministic synthetic weights of the v2 U-Net as a plain ``state_dict`` (names / shapes of
``/root/reference/src/unet.py:83-98``): inputs for tests, bench.py and the profiling tools. Not part of the oracle (no
reference arithmetic here) and not part of the product package."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import detrand

V2_HEADS = (1, 14, 3, 2, 1, 360, 60, 60)

# (state_dict prefix, Cin, Cout) of every DoubleConv in forward order -- unet.py:83-95
_DOUBLE_CONVS = OrderedDict([
    ("inc1.double_conv", None),                       # Cin = in_channels
    ("inc2.double_conv", (16, 16)),
    ("down1.maxpool_conv.1.double_conv", (16, 32)),
    ("down2.maxpool_conv.1.double_conv", (32, 64)),
    ("inc3.double_conv", (64, 64)),
    ("down3.maxpool_conv.1.double_conv", (64, 128)),
    ("down4.maxpool_conv.1.double_conv", (128, 256)),
    ("down5.maxpool_conv.1.double_conv", (256, 512)),
    ("up1.conv.double_conv", (512, 256)),
    ("up2.conv.double_conv", (256, 128)),
    ("up3.conv.double_conv", (128, 128)),
    ("dconv1.double_conv", (128, 128)),
    ("dconv2.double_conv", (128, 128)),
])
_UPS = OrderedDict([("up1.up", 512), ("up2.up", 256), ("up3.up", 128)])


def param_shapes(in_channels: int = 1, heads=V2_HEADS) -> "OrderedDict[str, tuple]":
    """All state_dict entries of UNet(in_channels, heads) with their shapes (261 for v2)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out["s"] = (10,)

    def bn(prefix, c):
        out[prefix + ".weight"] = (c,)
        out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,)
        out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    def dc(prefix, cin, cout):
        out[prefix + ".0.weight"] = (cout, cin, 3, 3)
        out[prefix + ".0.bias"] = (cout,)
        bn(prefix + ".1", cout)
        out[prefix + ".3.weight"] = (cout, cout, 3, 3)
        out[prefix + ".3.bias"] = (cout,)
        bn(prefix + ".4", cout)

    dcs = dict(_DOUBLE_CONVS)
    dcs["inc1.double_conv"] = (in_channels, 16)
    for name in list(dcs)[:8]:                       # encoder, unet.py:83-90
        dc(name, *dcs[name])
    for up, c in _UPS.items():                       # decoder, unet.py:91-93
        out[up + ".weight"] = (c, c // 2, 3, 3)      # ConvTranspose2d layout [Cin, Cout, kh, kw]
        out[up + ".bias"] = (c // 2,)
        name = up.split(".")[0] + ".conv.double_conv"
        dc(name, *dcs[name])
    dc("dconv1.double_conv", 128, 128)
    dc("dconv2.double_conv", 128, 128)
    for i, h in enumerate(heads):
        p = f"out_modules.{i}"
        out[p + ".conv1.weight"] = (128, 128, 3, 3)
        out[p + ".conv1.bias"] = (128,)
        bn(p + ".bn", 128)
        out[p + ".conv2.weight"] = (h, 128, 1, 1)
        out[p + ".conv2.bias"] = (h,)
    return out


def make_state_dict(seed: int = 0, in_channels: int = 1, heads=V2_HEADS, variant: str = "W1"):
    """Deterministic weights (platform exact).

    W0: He-uniform convs, BN at its torch default (identity in eval mode).
    W1: additionally randomised BN affine / running statistics so that BN folding is
        exercised, and biases of the centre / omega heads (outputs 0, 4, 7) shifted by -3 so
        that peak density is moderate.
    """
    sd = OrderedDict()
    for name, shape in param_shapes(in_channels, heads).items():
        k = detrand.key(name, seed)
        if name == "s":
            v = detrand.normalish(k, shape, 0.01)
        elif name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        elif name.endswith("running_mean"):
            v = detrand.uniform(k, shape, -0.2, 0.2) if variant == "W1" else np.zeros(shape, np.float32)
        elif name.endswith("running_var"):
            v = detrand.uniform(k, shape, 0.5, 2.0) if variant == "W1" else np.ones(shape, np.float32)
        elif len(shape) == 4:
            if ".up." in name:
                fan_in = shape[0] * 9 / 4.0          # stride-2 transposed conv: ~9/4 taps per output
                gain = 1.0
            else:
                fan_in = shape[1] * shape[2] * shape[3]
                gain = 2.0 if shape[2] == 3 else 1.0
            a = float(np.sqrt(3.0 * gain / fan_in))
            v = detrand.uniform(k, shape, -a, a)
        elif ".1.weight" in name or ".4.weight" in name or ".bn.weight" in name:
            v = detrand.uniform(k, shape, 0.5, 1.5) if variant == "W1" else np.ones(shape, np.float32)
        else:                                           # conv / BN biases
            v = detrand.uniform(k, shape, -0.1, 0.1)
            if variant == "W1" and name in ("out_modules.0.conv2.bias", "out_modules.4.conv2.bias",
                                            "out_modules.7.conv2.bias"):
                v = v - np.float32(3.0)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd
