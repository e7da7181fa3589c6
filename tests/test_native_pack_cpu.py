"""The C++ weight packing of the whole-network entry (csrc/unet_plan.cu: BatchNorm fold + plain / row-folded / swap-folded /
sub-pixel / head layouts, abc_unet_pack_host) against the Python packing of abcnet_b200.UNet.prepare: the SAME BYTES, layer by
layer, for bf16 and fp16, in_channels 1 and 3, both crop sides. CPU only (pure data movement)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import unet_ref

HEADS = list(unet_ref.V2_HEADS)


def _align(buf):
    while len(buf) % 256:
        buf.append(0)


def _python_arena(m):
    """The packs of UNet.prepare() concatenated in the order of csrc/unet_plan.cu (stem, DoubleConvs, up-convs, heads), every
    piece 256-byte aligned."""
    P = m._packed
    buf = bytearray()

    def put(t):
        _align(buf)
        buf.extend(t.contiguous().view(torch.uint8).numpy().tobytes() if t.dtype != torch.float32 else t.contiguous().numpy().tobytes())
    w0, b0 = P["inc1.0"]
    put(w0)
    put(b0)
    order = ["inc1.3"] + [f"{n}.{h}" for n in ("inc2", "down1", "down2", "inc3", "down3", "down4", "down5", "up1.conv", "up2.conv", "up3.conv",
                                               "dconv1", "dconv2") for h in ("0", "3")]
    order += ["up1.up", "up2.up", "up3.up", "heads.conv1"] + [f"heads.{i}.conv2" for i in range(len(m.heads))]
    for name in order:
        put(P[name].w)
        put(P[name].bias)
    return bytes(buf)


@pytest.mark.parametrize("cin,crop_first,act", [(1, True, "bf16"), (3, False, "bf16"), (1, True, "fp16")])
def test_cpp_packing_equals_python_packing(cin, crop_first, act):
    import abcnet_b200
    from abcnet_b200._lib import AbcNamedTensor, AbcUNetConfig, lib
    sd = unet_ref.make_state_dict(seed=41, in_channels=cin, variant="W1")
    m = abcnet_b200.UNet(cin, HEADS, crop_first=crop_first, act_dtype=act).eval()
    m.load_state_dict(sd)
    m.prepare(_inspect_on_cpu=True)
    want = _python_arena(m)
    cfg = AbcUNetConfig()
    cfg.in_channels, cfg.n_heads, cfg.crop_first, cfg.act_fp16 = cin, len(HEADS), int(crop_first), int(act == "fp16")
    for i, h in enumerate(HEADS):
        cfg.heads[i] = h
    keep = [(k.encode(), v.detach().float().contiguous()) for k, v in sd.items() if torch.is_floating_point(v)]
    arr = (AbcNamedTensor * len(keep))()
    for i, (k, t) in enumerate(keep):
        arr[i].name, arr[i].data, arr[i].numel = k, t.data_ptr(), t.numel()
    cap = lib.abc_unet_wpack_bytes(C.byref(cfg))
    out = np.zeros(cap, np.uint8)
    used = C.c_int64(0)
    rc = lib.abc_unet_pack_host(C.byref(cfg), arr, len(keep), out.ctypes.data, cap, C.byref(used))
    assert rc == 0, lib.abc_last_error()
    assert used.value <= cap and used.value == len(want), (used.value, cap, len(want))
    got = out[:used.value].tobytes()
    if got != want:
        a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
        first = int(np.nonzero(a != b)[0][0])
        raise AssertionError(f"packs differ from byte {first} on ({int((a != b).sum())} bytes)")
    # a missing tensor is reported by name, without a GPU
    rc = lib.abc_unet_pack_host(C.byref(cfg), arr, len(keep) - 1, out.ctypes.data, cap, C.byref(used))
    assert rc != 0 and keep[-1][0] in lib.abc_last_error()
