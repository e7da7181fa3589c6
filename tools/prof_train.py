#!/usr/bin/env python
"""Per-kernel breakdown of one training step (forward + fused losses + backward) with torch.profiler (CUPTI):
  python tools/prof_train.py [batch] [size]  ->  table of device time per kernel name, plus the host wall time of the step
(tooling only: finds where the step time goes; bench numbers are never taken under a profiler)."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
import synthdata as synth  # noqa: E402  (deterministic synthetic weights / images / targets; the oracle is not used here)
import synthdata as unet_ref  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
model = abcnet_b200.UNet(1, list(unet_ref.V2_HEADS)).to(dev)
model.load_state_dict(unet_ref.make_state_dict(0))
model.train()
crit = abcnet_b200.HeatmapLoss(class_weights=True)
x = torch.from_numpy(synth.binary_images(0, 8, S, S, 0.05)).repeat(B // 8, 1, 1, 1).to(dev)
tg = [torch.from_numpy(t).repeat(*([B // 8] + [1] * (t.ndim - 1))).to(dev).contiguous() for t in synth.dense_targets(0, 8, S // 4, S // 4)]


def step():
    for p in model.parameters():
        p.grad = None
    outs = model(x)
    loss = crit(outs, tg, model.s)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue time {t_host * 1e3:.1f} ms, step wall {t_all * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(e.name, [0, 0.0])
        r[0] += 1
        r[1] += e.device_time
tot = sum(v[1] for v in rows.values())
print(f"total device time {tot / 1e3:.2f} ms over {sum(v[0] for v in rows.values())} kernels")
for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 1e3:9.3f} ms  {n:5d}x  {name[:110]}")

# the BatchNorm / activation passes in launch order (forward: layer order, backward: reversed), one duration per launch
seq = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "bn_" in e.name and "finalize" not in e.name),
             key=lambda e: e.time_range.start)
print("BatchNorm passes in launch order (us):")
for kind in ("bn_act_kernel", "bn_stats_kernel", "bn_act_bwd_reduce_kernel", "bn_act_bwd_apply_kernel"):
    print(f"  {kind}: " + " ".join(f"{e.device_time:.0f}{'p' if '<true>' in e.name or '<(bool)1>' in e.name else ''}" for e in seq if kind in e.name))

# per-launch CUDA-event timing of the tensor-core kernels (conv fprop / dgrad, wgrad), grouped by shape
from abcnet_b200 import train as _tr  # noqa: E402
_tr.timing = []
step()
torch.cuda.synchronize()
agg = {}
for label, a, b in _tr.timing:
    r = agg.setdefault(label, [0, 0.0])
    r[0] += 1
    r[1] += a.elapsed_time(b)
_tr.timing = None
print("tensor-core launches by shape:")
for label, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  {n:3d}x  {label}")
