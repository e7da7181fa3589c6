#!/usr/bin/env python
"""Training-step throughput (BASELINE.json configs[3] / [4]): forward (batch-statistics BN, dropout) + fused losses +
backward (+ bucketed gradient all-reduce under torchrun) at the reference's batch size 64 per GPU, 1x512x512 images.
  python tools/train_bench.py [--batch 64] [--steps 5]            # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py   # data parallel
The optimiser step is torch.optim.Adam (fused multi-tensor Adam is a 'next' row) and is timed separately."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
from abcnet_b200.ddp import GradBuckets  # noqa: E402
import synthdata as synth  # noqa: E402  (deterministic synthetic weights / images / targets; the oracle is not used here)
import synthdata as unet_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--mode", default="graph", choices=["graph", "eager", "autograd"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.size
    model = abcnet_b200.UNet(1, list(unet_ref.V2_HEADS)).to(dev)
    model.load_state_dict(unet_ref.make_state_dict(0))
    model.train()
    params = [p for p in model.parameters()]
    buckets = GradBuckets(params) if world > 1 else None
    model.grad_buckets = buckets
    x = torch.from_numpy(synth.binary_images(rank, 8, S, S, 0.05)).repeat(B // 8, 1, 1, 1).to(dev)
    tg = [torch.from_numpy(t).repeat(*([B // 8] + [1] * (t.ndim - 1))).to(dev).contiguous() for t in synth.dense_targets(rank, 8, S // 4, S // 4)]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    times = {"fwd": [], "loss": [], "bwd": [], "opt": [], "step": []}
    if args.mode in ("graph", "eager"):
        # the fast path: abcnet_b200.TrainStep (no autograd; one CUDA-graph replay per iteration in graph mode)
        opt = abcnet_b200.make_optimizer(model, capturable=args.mode == "graph")
        step = abcnet_b200.TrainStep(model, opt, class_weights=True, buckets=buckets, use_graph=args.mode == "graph")
        for it in range(args.warmup + args.steps):
            e0, e1 = ev(), ev()
            e0.record()
            loss = step(x, tg)
            e1.record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                times["step"].append(e0.elapsed_time(e1))
        for k in ("fwd", "loss", "bwd", "opt"):
            times[k].append(float("nan"))
    else:
        crit = abcnet_b200.HeatmapLoss(class_weights=True)
        opt = torch.optim.Adam(params, lr=2.5e-4, weight_decay=1e-8)
        for it in range(args.warmup + args.steps):
            e = [ev() for _ in range(5)]
            if buckets is not None:
                buckets.zero()
            else:
                opt.zero_grad(set_to_none=True)
            e[0].record()
            outs = model(x)
            e[1].record()
            loss = crit(outs, tg, model.s)
            e[2].record()
            loss.backward()
            if buckets is not None:
                buckets.finish()
            e[3].record()
            opt.step()
            e[4].record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                for k, a, b in (("fwd", 0, 1), ("loss", 1, 2), ("bwd", 2, 3), ("opt", 3, 4), ("step", 0, 4)):
                    times[k].append(e[a].elapsed_time(e[b]))
    ms = {k: sum(v) / len(v) for k, v in times.items()}
    if world > 1:
        t = torch.tensor([ms["step"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms["step"] = t.item()
    if rank == 0:
        print(json.dumps({"metric": "images_per_sec_train_step", "value": world * B / (ms["step"] * 1e-3), "unit": "images/s",
                          "n_gpus": world, "batch_per_gpu": B, "mode": args.mode, "ms": ms, "loss": float(loss.item()),
                          "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                          "train_tflops": 281.9e9 * world * B / (ms["step"] * 1e-3) / 1e12}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
