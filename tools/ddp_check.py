#!/usr/bin/env python
"""Two-or-more-rank check of the data-parallel training path on real GPUs (NCCL), launched under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py

Each rank runs one TrainStep on its own batch (a) without buckets -> local gradients, all-reduced (mean) by a plain NCCL
call, and (b) with GradBuckets (overlapped, bucketed all-reduce launched from inside the backward pass; eager and CUDA-graph
replay). (b) must equal (a); the replica weights after one Adam step must be identical on all ranks. Replaces what
DistributedDataParallel guarantees in /root/reference/src/multi_gpu_train2.py:89."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
from abcnet_b200.ddp import GradBuckets  # noqa: E402
import synthdata as synth  # noqa: E402  (deterministic synthetic weights / images / targets; the oracle is not used here)
import synthdata as unet_ref  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 4, 64
HEADS = list(unet_ref.V2_HEADS)
sd = unet_ref.make_state_dict(seed=1, variant="W1")
x = torch.from_numpy(synth.binary_images(10 + rank, B, S, S, 0.08)).to(dev)
tg = [torch.from_numpy(t).to(dev).contiguous() for t in synth.dense_targets(20 + rank, B, S // 4, S // 4)]


def fresh():
    m = abcnet_b200.UNet(1, HEADS).to(dev)
    m.load_state_dict(sd)
    m.train()
    m.dropout_p = 0.0          # identical masks are not the point here; keep the comparison deterministic
    return m


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


# (a) local gradients, reduced by hand
m0 = fresh()
abcnet_b200.TrainStep(m0, None, buckets=None, use_graph=False)(x, tg)
ref = []
for p in m0.parameters():
    g = p.grad.detach().clone()
    dist.all_reduce(g)
    ref.append(g / world)
worst = {}
for mode, graph in (("eager", False), ("graph", True)):
    m1 = fresh()
    gb = GradBuckets(list(m1.parameters()))
    step = abcnet_b200.TrainStep(m1, None, buckets=gb, use_graph=graph)
    step(x, tg)
    if graph:
        step(x, tg)            # a replay must give the same gradients again
    torch.cuda.synchronize()
    w = 0.0
    for (n, p), g in zip(m1.named_parameters(), ref):
        if g.abs().max() == 0:
            assert p.grad.abs().max() == 0, n
            continue
        w = max(w, rel(p.grad, g))
    worst[mode] = w
# replicas stay identical through an optimiser step
m2 = fresh()
gb2 = GradBuckets(list(m2.parameters()))
opt = abcnet_b200.make_optimizer(m2, capturable=True)
st2 = abcnet_b200.TrainStep(m2, opt, buckets=gb2, use_graph=True)
for _ in range(3):
    st2(x, tg)
torch.cuda.synchronize()
flat = torch.cat([p.detach().flatten() for p in m2.parameters()])
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
spread = (hi - lo).abs().max().item()
moved = (flat - torch.cat([p.detach().flatten() for p in m0.parameters()])).abs().max().item()
if rank == 0:
    print(f"ddp_check world={world} buckets={len(gb.buckets)} worst_rel_err={worst} replica_spread={spread:.3e} moved={moved:.3e}")
    # wgrad accumulates with fp32 atomics -> run-to-run rounding differences only
    assert worst["eager"] < 2e-3 and worst["graph"] < 2e-3, worst
    assert spread == 0.0, spread
    assert moved > 0
    print("ddp_check ok", flush=True)
# CUDA graphs that captured NCCL kernels must die before the communicator does (destroy_process_group blocks otherwise)
del step, st2, gb, gb2, opt
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
