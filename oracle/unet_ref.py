"""Functional fp32 restatement of the reference U-Net (test infrastructure).

Follows ``/root/reference/src/unet.py``:
  * DoubleConv  :6-21   (conv3x3 pad1 + bias -> BatchNorm2d(eps 1e-5) -> ReLU) x2
  * Down        :24-35  MaxPool2d(2) then DoubleConv
  * Up          :38-60  ConvTranspose2d(C, C/2, k3, s2) -> crop to the skip size -> cat([skip, up]) -> DoubleConv
  * OutConv     :63-74  conv3x3 + bias -> BN -> LeakyReLU(0.01) -> Dropout(0.2) -> conv1x1
  * UNet        :77-119 wiring; forward returns a list of len(heads) tensors at stride 4

The arithmetic itself lives in PyTorch (unpinned third-party dependency of the reference, see
SURVEY.md section 8c); this module calls the same ATen CPU ops through ``torch.nn.functional`` on a
plain ``state_dict`` so that it needs neither the reference sources nor the product package.
It is pinned against the imported reference by ``tests/golden/make_golden.py``.

Crop side (``unet.py:51-55``): ``diff // 2`` on a tensor is floor division under torch 2.x, so
for diff = -1 the pad list is [-1, 0, -1, 0] = drop the FIRST row and column.
``crop_first=False`` gives the pre-1.13 truncating behaviour (drop the last row/column).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import detrand

V2_HEADS = (1, 14, 3, 2, 1, 360, 60, 60)
BN_EPS = 1e-5
LEAKY = 0.01

from synthdata.weights import _DOUBLE_CONVS, _UPS, make_state_dict, param_shapes  # noqa: E402,F401  (synthetic weights live outside the oracle)


def _strip(sd):
    """Accept DataParallel / DDP checkpoints (``module.`` prefix, train.py:435)."""
    if any(k.startswith("module.") for k in sd):
        return OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
    return sd


def _bn(x, sd, p, training, stats_out=None):
    if training:
        y = F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, BN_EPS)
        if stats_out is not None:
            stats_out[p] = (x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False))
        return y
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.1, BN_EPS)


def bf16_ste(t):
    """Round to bf16 (value) with a straight-through gradient -- used to EMULATE the storage precision of the CUDA path
    (bf16 conv outputs / activations / weights, fp32 accumulation) inside the fp32 oracle. Train-mode BatchNorm makes a
    randomly initialised net amplify perturbations ~370x from the first layer to the trunk (fp32 vs fp64 oracle: 5e-8 ->
    2e-5), so bf16 storage noise (2e-3) cannot be compared end-to-end against the unrounded oracle; rounding at the same
    points removes the storage noise from the comparison and leaves accumulation-order effects only."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def _id(t):
    return t


def _double_conv(x, sd, p, training, acts=None, rnd=_id, first_fp32=False):
    # first_fp32: the CUDA path keeps the 1 -> 16 stem convolution's weights in fp32 (direct kernel, not tensor cores)
    w0 = sd[p + ".0.weight"] if first_fp32 else rnd(sd[p + ".0.weight"])
    x = rnd(F.conv2d(x, w0, sd[p + ".0.bias"], padding=1))
    x = rnd(F.relu(_bn(x, sd, p + ".1", training)))
    if acts is not None:
        acts[p + ".0"] = x
    x = rnd(F.conv2d(x, rnd(sd[p + ".3.weight"]), sd[p + ".3.bias"], padding=1))
    x = rnd(F.relu(_bn(x, sd, p + ".4", training)))
    if acts is not None:
        acts[p + ".3"] = x
    return x


def _up(x1, x2, sd, name, training, crop_first, acts=None, rnd=_id):
    x1 = rnd(F.conv_transpose2d(x1, rnd(sd[name + ".up.weight"]), sd[name + ".up.bias"], stride=2))
    dy = x1.shape[2] - x2.shape[2]
    dx = x1.shape[3] - x2.shape[3]
    if crop_first:
        x1 = x1[:, :, dy:, dx:]
    else:
        x1 = x1[:, :, : x1.shape[2] - dy, : x1.shape[3] - dx]
    if acts is not None:
        acts[name + ".up"] = x1
    x = torch.cat([x2, x1], dim=1)
    return _double_conv(x, sd, name + ".conv.double_conv", training, acts, rnd)


def trunk(x, sd, training=False, crop_first=True, acts=None, rnd=_id):
    """unet.py:101-115 -- everything before the heads."""
    sd = _strip(sd)
    x1 = _double_conv(x, sd, "inc1.double_conv", training, acts, rnd, first_fp32=True)
    x1 = _double_conv(x1, sd, "inc2.double_conv", training, acts, rnd)
    x2 = _double_conv(F.max_pool2d(x1, 2), sd, "down1.maxpool_conv.1.double_conv", training, acts, rnd)
    x3 = _double_conv(F.max_pool2d(x2, 2), sd, "down2.maxpool_conv.1.double_conv", training, acts, rnd)
    x3 = _double_conv(x3, sd, "inc3.double_conv", training, acts, rnd)
    x4 = _double_conv(F.max_pool2d(x3, 2), sd, "down3.maxpool_conv.1.double_conv", training, acts, rnd)
    x5 = _double_conv(F.max_pool2d(x4, 2), sd, "down4.maxpool_conv.1.double_conv", training, acts, rnd)
    x6 = _double_conv(F.max_pool2d(x5, 2), sd, "down5.maxpool_conv.1.double_conv", training, acts, rnd)
    x = _up(x6, x5, sd, "up1", training, crop_first, acts, rnd)
    x = _up(x, x4, sd, "up2", training, crop_first, acts, rnd)
    x = _up(x, x3, sd, "up3", training, crop_first, acts, rnd)
    x = _double_conv(x, sd, "dconv1.double_conv", training, acts, rnd)
    x = _double_conv(x, sd, "dconv2.double_conv", training, acts, rnd)
    return x


def forward(x, sd, heads=V2_HEADS, training=False, crop_first=True, acts=None, dropout_masks=None, emulate_bf16=False):
    """UNet.forward (unet.py:100-119). ``dropout_masks`` (list of 0/1 tensors [B,128,H,W] or None)
    replaces nn.Dropout(0.2) in training mode so that tests are deterministic (scale 1/0.8).
    ``emulate_bf16`` rounds weights, conv outputs and activations to bf16 (straight-through) exactly where the CUDA
    training path stores bf16 -- see ``bf16_ste``."""
    sd = _strip(sd)
    rnd = bf16_ste if emulate_bf16 else _id
    t = trunk(x, sd, training, crop_first, acts, rnd)
    return heads_forward(t, sd, heads, training, acts, dropout_masks, rnd)


def heads_forward(t, sd, heads=V2_HEADS, training=False, acts=None, dropout_masks=None, rnd=_id, round_hidden=None):
    """The eight OutConv heads (unet.py:63-74, :116-118) on a given trunk tensor [B,128,H/4,W/4]. Split out of ``forward``
    so that tests can feed the heads with the product's own trunk activations (error budget per stage).
    ``round_hidden``: optional rounding applied to the hidden map only (e.g. ``bf16_ste``)."""
    sd = _strip(sd)
    outs = []
    for i, _ in enumerate(heads):
        p = f"out_modules.{i}"
        h = rnd(F.conv2d(t, rnd(sd[p + ".conv1.weight"]), sd[p + ".conv1.bias"], padding=1))
        h = F.leaky_relu(_bn(h, sd, p + ".bn", training), LEAKY)
        if training and dropout_masks is not None:
            h = h * dropout_masks[i] / 0.8
        h = rnd(h)
        if round_hidden is not None:
            h = round_hidden(h)
        if acts is not None:
            acts[p + ".hidden"] = h
        outs.append(F.conv2d(h, rnd(sd[p + ".conv2.weight"]), sd[p + ".conv2.bias"]))
    return outs
