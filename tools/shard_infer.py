#!/usr/bin/env python
"""Image-sharded inference over a large synthetic set (BASELINE.json configs[2]: the multi_proc_img2smiles.py replacement).

    python tools/shard_infer.py --images 102400                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tools/shard_infer.py --images 102400 [--sparse]                             # N GPUs, one process each

Every rank takes a contiguous range of the global image list (abcnet_b200.shard.shard_bounds), runs U-Net forward + decode
on it in batches, assembles each image's records into MOL-block text with the native host assembler while the GPU works on
the next batch, and hashes the texts. There is NO data-path collective: only one 32-byte digest per batch and a molecule
count are gathered at the end (shard.gather_results). Image g of the global set is deterministic -- image g % 64 of a seeded
pool, rolled by (g // 64) % W pixels -- so the combined digest must not depend on the number of GPUs: run with N = 1 and
N = 2 and compare "digest". Rank 0 prints one JSON line.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
from abcnet_b200 import shard  # noqa: E402
import synthdata  # noqa: E402  (deterministic synthetic weights / images; the oracle is not used here)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=102400)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sparse", action="store_true", help="SparseHeadsPipeline instead of dense heads + PeakDecoder")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl" if not os.environ.get("SHARD_GLOO") else "gloo", device_id=dev)
    B = args.batch
    total = args.images // (B * world) * (B * world)   # whole batches per rank: the batch ranges (hence the digest) do not depend on N
    heads = list(synthdata.V2_HEADS)
    model = abcnet_b200.UNet(1, heads).to(dev).eval()
    model.load_state_dict(synthdata.make_state_dict(seed=0, variant="W1"))
    pool = torch.from_numpy(synthdata.binary_images(1000, 64, 512, 512, 0.05)).to(dev)
    W = pool.shape[-1]
    with torch.no_grad():                              # same calibration of the centre / omega biases as bench.py
        outs = model(pool[:8].contiguous())
        for k in (0, 4, 7):
            model.out_modules[k].conv2.bias += -1.0 - torch.quantile(outs[k].flatten()[:4_000_000].float(), 0.997)
    cols = torch.arange(W, device=dev)

    def make_batch(a, b):
        g = torch.arange(a, b, device=dev)
        shift = (g // 64) % W
        idx = (cols[None, :] - shift[:, None]) % W                                   # [n, W] source column of every output column
        x = pool[g % 64]                                                             # [n, 1, H, W]
        x = torch.take_along_dim(x, idx[:, None, None, :].expand(-1, 1, x.shape[2], -1), dim=3)
        if b - a < B:                                                                # pad the last batch of the shard
            x = torch.cat([x, x.new_zeros((B - (b - a),) + tuple(x.shape[1:]))])
        return x.contiguous()

    if args.sparse:
        sinks = [abcnet_b200.SparseHeadsPipeline(model, B, peak_cap=128, bond_cap=4096, device=dev) for _ in range(2)]
    else:
        sinks = [abcnet_b200.PeakDecoder(B, atom_cap=1024, bond_cap=4096, device=dev) for _ in range(2)]
    obufs = [None, None]
    lo, hi = shard.shard_bounds(total, rank, world)
    ranges = list(shard.batches(lo, hi, B))
    digests, n_mol = [], 0

    def finish(i):
        nonlocal n_mol
        a, b = ranges[i]
        sinks[i % 2].wait(B)
        texts = sinks[i % 2].molblocks(B, max(1, (os.cpu_count() or 1) // world))[:b - a]
        h = hashlib.blake2b(digest_size=32)
        for t in texts:
            h.update(b"\0" if t is None else t.encode())
            h.update(b"\1")
        n_mol += sum(t is not None for t in texts)
        digests.append((a, h.hexdigest()))

    def enqueue(i):
        x = make_batch(*ranges[i])
        if args.sparse:
            sinks[i % 2].launch(x)
        else:
            obufs[i % 2] = model.infer(x, obufs[i % 2], layout="p8f")
            sinks[i % 2].launch(obufs[i % 2])
        sinks[i % 2].fetch_async(B)

    for i in range(min(2, len(ranges))):               # warm-up (buffers, weight packing)
        enqueue(i)
        sinks[i % 2].wait(B)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(len(ranges)):
        enqueue(i)
        if i > 0:
            finish(i - 1)
    if ranges:
        finish(len(ranges) - 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
        parts = [None] * world
        dist.all_gather_object(parts, (digests, n_mol))
    else:
        parts = [(digests, n_mol)]
    if rank == 0:
        allb = sorted(d for p, _ in parts for d in p)
        h = hashlib.blake2b(digest_size=16)
        # the combined digest is over per-IMAGE-range digests; to be independent of where the shard boundaries fall the ranges
        # must be the same for every N: true when total / world is a multiple of the batch size
        for a, d in allb:
            h.update(f"{a}:{d};".encode())
        print(json.dumps({"tool": "shard_infer", "images": total, "n_gpus": world, "batch": B, "sparse_heads": bool(args.sparse),
                          "seconds": dt, "images_per_s": total / dt, "molecules": int(sum(n for _, n in parts)),
                          "batches": len(allb), "digest": h.hexdigest(),
                          "what": "forward + decode + native MOL-block assembly per image, sharded by contiguous image ranges, "
                                  "no data-path collective; digest over all MOL-block texts in global image order"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
