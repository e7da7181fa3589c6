"""Same-box GPU comparator: the reference's graph executed by torch / cuDNN on the B200.

The reference (`/root/reference/src/unet.py:100-119`, `train.py:94-141`, `img2smiles.py:62-80,115-124`) is plain PyTorch, so
"the library path to beat" is that same graph on CUDA: cuDNN convolutions, ATen BatchNorm / pooling / losses, autograd.
`/root/reference` does not exist on the GPU box; the graph is therefore executed through the torch containers that
`abcnet_b200.UNet` keeps for its parameters (the same `nn.Conv2d` / `nn.BatchNorm2d` / `nn.ConvTranspose2d` objects, same
state_dict as `src/unet.py`), called in the order of `UNet.forward` -- stock torch ops only, none of this repo's kernels.
Three arithmetic modes, as SURVEY asks: fp32 with TF32 off (the CPU oracle's arithmetic), fp32 default (cuDNN TF32 allowed, what
a user of the reference gets on Ampere+), bf16 autocast + channels_last (the strongest library configuration).

Not product code and not the parity oracle: a measured baseline, imported by bench.py's `gpu_comparator` leg only.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

ATOM_TYPE_WEIGHTS = [1, 0.1, 0.1, 0.1, 1, 1, 1, 1, 1, 10, 10, 10, 10, 10]          # train.py:16


def torch_forward(m, x):
    """unet.py:100-119 on m's parameter containers (m: abcnet_b200.UNet; BN / dropout follow m.training)."""
    def dc(h, t):
        return h.double_conv(t)

    def down(h, t):
        return h.maxpool_conv[1].double_conv(F.max_pool2d(t, 2))

    def up(h, t, skip):
        u = h.up(t)
        dy, dx = u.shape[2] - skip.shape[2], u.shape[3] - skip.shape[3]
        u = u[:, :, dy:, dx:] if m.crop_first else u[:, :, :u.shape[2] - dy, :u.shape[3] - dx]      # unet.py:51-55
        return h.conv.double_conv(torch.cat([skip, u], 1))

    x1 = dc(m.inc2, dc(m.inc1, x))
    x2 = down(m.down1, x1)
    x3 = dc(m.inc3, down(m.down2, x2))
    x4 = down(m.down3, x3)
    x5 = down(m.down4, x4)
    x6 = down(m.down5, x5)
    t = up(m.up3, up(m.up2, up(m.up1, x6, x5), x4), x3)
    t = dc(m.dconv2, dc(m.dconv1, t))
    outs = []
    for om in m.out_modules:
        h = F.leaky_relu(om.bn(om.conv1(t)), 0.01)
        outs.append(om.conv2(F.dropout(h, m.dropout_p, m.training)))
    return outs


def dense_decode_ops(outs):
    """The tensor statements of img2smiles.py:62-80, :115-124 (NMS masks, dense arg-max maps) -- the part of the reference's
    decode that runs on the device; its per-peak Python loop (:134-193, one .item() per scalar) is not timed here."""
    za, zt, zc, zh, zb, zbt, zr, zw = [o.float() for o in outs]
    atom_pk = (F.max_pool2d(za, 3, 1, 1) == za) * (za > -1)
    bond_pk = (F.max_pool2d(zb, 3, 1, 1) == zb) * (zb > -1)
    B, nw, H, W = zw.shape
    col = zw.permute(0, 2, 3, 1).reshape(-1, 1, nw)
    pad = torch.cat([col[..., -1:], col, col[..., :1]], -1)
    omega_pk = ((F.max_pool1d(pad, 3, 1) == col) * (col > -1)).view(B, H, W, nw)
    return (atom_pk, bond_pk, omega_pk, zt.argmax(1), zc.argmax(1), zh.argmax(1),
            zbt.view(B, -1, nw, H, W).argmax(1), zr.abs())


def record_level_maps(outs, thr=-1.0):
    """Everything the decoded records depend on, as dense tensors (img2smiles.py:62-80, :115-124 and the per-peak rules of :134-182
    vectorised): atom peaks + their three classes, and for every bond-centre peak the EMITTED omega bins (3-bin circular NMS, the
    > thr test and the asymmetric half-circle rule of :143-158) + their bond types. Two logit sets give identical records (up to
    the float rho) iff these maps agree -- used by tools/shard_infer.py --ref-check to compare the product with the fp32 torch
    forward on every image of the 100 k-image run. outs: 8 NCHW fp32 tensors."""
    za, zt, zc, zh, zb, zbt, zr, zw = [o.float() for o in outs]
    B, nw, H, W = zw.shape
    h = nw // 2
    atom_pk = (F.max_pool2d(za, 3, 1, 1) == za) & (za > thr)
    bond_pk = (F.max_pool2d(zb, 3, 1, 1) == zb) & (zb > thr)
    left, right = torch.roll(zw, 1, 1), torch.roll(zw, -1, 1)
    cand = (zw >= left) & (zw >= right) & (zw > thr)
    surv = torch.zeros_like(cand)
    lo = zw[:, :h - 1]                                                     # omega <= h - 2: dropped if z < max(z[w + h - 1], z[w + h])
    surv[:, :h - 1] = (lo >= zw[:, h - 1:2 * h - 2]) & (lo >= zw[:, h:2 * h - 1])
    surv[:, h - 1] = (zw[:, h - 1] >= zw[:, nw - 2]) & (zw[:, h - 1] >= zw[:, 0])      # sic: bins n - 2 and 0
    surv[:, h] = (zw[:, h] > zw[:, 0]) & (zw[:, h] > zw[:, nw - 1])
    hi = zw[:, h + 1:]                                                     # omega >= h + 1: dropped if z <= max(z[w - h - 1], z[w - h])
    surv[:, h + 1:] = (hi > zw[:, :h - 1]) & (hi > zw[:, 1:h])
    emitted = cand & surv & bond_pk
    a_cls = torch.stack([zt.argmax(1), zc.argmax(1), zh.argmax(1)], 1) * atom_pk
    b_type = zbt.view(B, -1, nw, H, W).argmax(1) * emitted
    return atom_pk, a_cls, bond_pk, emitted, b_type


def reference_loss(outs, targets, s, type_w):
    """train.py:95-137 as torch statements (same formulas as SURVEY App. C); rho / omega targets may be float64 (utils.py:91-92)."""
    za, zt, zc, zh, zb, zbt, zr, zw = [o.float() for o in outs]
    ta, tt, tc, th, tb, tbt, tr, tw = targets
    lo, hi = 1e-5, 1 - 1e-5
    B, nw, H, W = zw.shape

    def focal(p, t):
        return torch.sum(-(t == 1).to(p.dtype) * (1 - p) ** 2 * torch.log(p) - (1 - t) ** 4 * p ** 2 * torch.log(1 - p)) / torch.sum(t == 1)

    def ce(z, t, w=1.0, extra=0.0):
        p = torch.clamp(torch.softmax(z, 1), lo, hi)
        return torch.sum(-w * t * (1 - p) ** 2 * torch.log(p)) / (torch.sum(t) + extra)

    pa, pb, pw = (torch.clamp(torch.sigmoid(z), lo, hi) for z in (za, zb, zw))
    L = [focal(pa, ta) * (torch.exp(-s[0]) + s[0]), focal(pb, tb) * (torch.exp(-s[1]) + s[1]),
         ce(zt, tt, type_w.view(1, -1, 1, 1)) * (torch.exp(-s[2]) + s[2]), ce(zc, tc) * (torch.exp(-s[3]) + s[3]),
         ce(zbt.view(B, -1, nw, H, W), tbt) * (torch.exp(-s[4]) + s[4]),
         torch.sum(torch.abs(zr.abs() - tr) * tbt.sum(1)) / tbt.sum() * (0.5 * torch.exp(-s[6]) + s[6]),
         -torch.sum(tw.sum(1, keepdim=True) * ((tw == 1) * (1 - pw) ** 2 * torch.log(pw) + (1 - tw) ** 4 * pw ** 2 * torch.log(1 - pw)))
         / tw.sum() * (torch.exp(-s[7]) + s[7]),
         ce(zh, th, extra=0.1) * (torch.exp(-s[9]) + s[9])]
    return sum(L)


@contextlib.contextmanager
def _mode(name):
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = name != "fp32_tf32_off"
    torch.backends.cudnn.benchmark = True
    try:
        if name == "bf16_autocast_channels_last":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                yield
        else:
            yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


MODES = ("fp32_tf32_off", "fp32_default_tf32", "bf16_autocast_channels_last")


def _time(fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(model, x_infer, x_train, targets, iters=3, warmup=2):
    """model: abcnet_b200.UNet on CUDA (its containers are executed by torch); x_infer [B,1,H,W] fp32, x_train + dense targets
    for the training leg. Returns {mode: {"infer_img_s", "infer_ms", "train_img_s", "train_ms"}} (CUDA events)."""
    res = {}
    type_w = torch.tensor(ATOM_TYPE_WEIGHTS, dtype=torch.float32, device=x_infer.device)
    was_training = model.training
    for mode in MODES:
        cl = mode == "bf16_autocast_channels_last"
        r = {}
        try:
            if cl:
                model.to(memory_format=torch.channels_last)
            xi = x_infer.contiguous(memory_format=torch.channels_last) if cl else x_infer
            model.eval()

            def infer():
                with torch.no_grad():
                    return dense_decode_ops(torch_forward(model, xi))
            with _mode(mode):
                ms = _time(infer, iters, warmup)
            r.update(infer_ms=ms, infer_img_s=x_infer.shape[0] / (ms * 1e-3))
            if x_train is not None:
                model.train()
                nb = x_train.shape[0]
                while True:                               # autograd keeps ~1.7 GB of fp32 activations per image: halve on OOM
                    xt = x_train[:nb].contiguous(memory_format=torch.channels_last) if cl else x_train[:nb]
                    tg = [t[:nb] for t in targets]

                    def train():
                        for p in model.parameters():
                            p.grad = None
                        loss = reference_loss(torch_forward(model, xt), tg, model.s, type_w)
                        loss.backward()
                    try:
                        torch.cuda.reset_peak_memory_stats()
                        with _mode(mode):
                            ms = _time(train, iters, warmup)
                        break
                    except torch.cuda.OutOfMemoryError:
                        for p in model.parameters():
                            p.grad = None
                        torch.cuda.empty_cache()
                        if nb <= 8:
                            raise
                        nb //= 2
                r.update(train_ms=ms, train_batch=nb, train_img_s=nb / (ms * 1e-3),
                         train_peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
        except RuntimeError as e:                         # e.g. out of memory in one mode: reported, the others still run
            r["error"] = str(e)[:200]
        finally:
            for p in model.parameters():
                p.grad = None
            if cl:
                model.to(memory_format=torch.contiguous_format)
            torch.cuda.empty_cache()
        res[mode] = r
    model.train(was_training)
    return res
