#!/usr/bin/env python
"""Image-sharded inference over a large synthetic set (BASELINE.json configs[2]: the multi_proc_img2smiles.py replacement).

    python tools/shard_infer.py --images 102400                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tools/shard_infer.py --images 102400 [--sparse]                             # N GPUs, one process each

Every rank takes a contiguous range of the global image list (abcnet_b200.shard.shard_bounds), runs U-Net forward + decode
on it in batches, assembles each image's records into MOL-block text with the native host assembler while the GPU works on
the next batch, and hashes the texts. There is NO data-path collective: only one 32-byte digest per batch and a molecule
count are gathered at the end (shard.gather_results). Image g of the global set is deterministic -- image g % P of a seeded
pool, rolled by (g // P) % W pixels -- so the combined digest must not depend on the number of GPUs: run with N = 1 and
N = 2 and compare "digest". Rank 0 prints one JSON line.

--train-steps K (K > 0): instead of random-init weights, rank 0 first trains the network for K iterations on labelled
pseudo-molecule drawings (synthdata.molecules; labels -> abcnet_b200.parse_labels -> TargetRasteriser, i.e. the product's own
target path) and broadcasts the weights once (setup, not data path); the image pool is then 1024 pseudo-molecule drawings.
--dump-subset M: rank 0 writes gpurun_out/shard_subset.pt = {state_dict, MOL-block texts of global images [0, M)}; the CPU
oracle check of that subset (exact match vs the fp32 reference path, BASELINE configs[2]) is tests/shard_subset_check.py,
which needs no GPU and is run in the build container.
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


def train_on_pool(model, pool, labels, steps, dev, batch=16, n_train=256):
    """A short training run with the product path only: labels -> label strings -> parse_labels -> TargetRasteriser (GPU) ->
    TrainStep (forward + 8 losses + backward + fused Adam as one CUDA graph). Schedule as in tests/test_trained_gpu.py."""
    from synthdata import molecules
    model.train()
    lr = torch.tensor(6e-4, device=dev)
    opt = abcnet_b200.make_optimizer(model, lr=lr, capturable=True)
    step = abcnet_b200.TrainStep(model, opt, class_weights=True, use_graph=True)
    rast = abcnet_b200.TargetRasteriser(batch, 128, 128, device=dev)
    parsed = [abcnet_b200.parse_labels(*molecules.label_strings(labels[i])) for i in range(n_train)]
    nb = n_train // batch
    for it in range(steps):
        b = it % nb
        if it == (steps * 5) // 8:
            lr.fill_(2.5e-4)
        tg = rast(parsed[b * batch:(b + 1) * batch])
        step(pool[b * batch:(b + 1) * batch].contiguous(), tg)
    torch.cuda.synchronize()


def ref_check(model, make_batch, ranges, B, dev):
    """Every image of this rank's shard against the reference's graph executed by torch / cuDNN in fp32 with TF32 off (SURVEY D5:
    the arithmetic of the CPU oracle) on the same GPU and weights, at RECORD level: two logit sets decode to the same records iff
    gpu_comparator.record_level_maps agree (validated against the oracle decode on planted and dense random maps). Untimed second
    pass; per-rank counts are summed on rank 0."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_comparator as gc
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    c = dict(images=0, identical_records=0, identical_atom_peak_set=0, identical_bond_peak_set=0, identical_omega_set=0,
             identical_classes=0, ref_atom_peaks=0, ref_bond_records=0, differing_atom_peaks=0, differing_bond_records=0,
             tf32_identical_records=0, tf32_differing_atom_peaks=0, tf32_differing_bond_records=0)
    outs = None
    with torch.no_grad():
        for (a, b) in ranges:
            x = make_batch(a, b)
            n = b - a
            outs = model.infer(x, outs, layout="nchw")
            mo = gc.record_level_maps([o[:n] for o in outs])
            mr = gc.record_level_maps([o[:n] for o in gc.torch_forward(model, x)])
            eq = [(p == q).flatten(1).all(1) for p, q in zip(mo, mr)]            # per image: atom_pk, a_cls, bond_pk, emitted, b_type
            c["images"] += n
            c["identical_atom_peak_set"] += int(eq[0].sum())
            c["identical_bond_peak_set"] += int(eq[2].sum())
            c["identical_omega_set"] += int(eq[3].sum())
            c["identical_classes"] += int((eq[1] & eq[4]).sum())
            c["identical_records"] += int((eq[0] & eq[1] & eq[3] & eq[4]).sum())
            c["ref_atom_peaks"] += int(mr[0].sum())
            c["ref_bond_records"] += int(mr[3].sum())
            c["differing_atom_peaks"] += int((mo[0] != mr[0]).sum())
            c["differing_bond_records"] += int((mo[3] != mr[3]).sum())
            # yardstick: the reference's OWN default GPU arithmetic (cuDNN TF32 convolutions) against its fp32 arithmetic,
            # same weights, same images, same metric -- how stable are the decode decisions of this network at all?
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            mt = gc.record_level_maps([o[:n] for o in gc.torch_forward(model, x)])
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            eqt = [(p == q).flatten(1).all(1) for p, q in zip(mt, mr)]
            c["tf32_identical_records"] += int((eqt[0] & eqt[1] & eqt[3] & eqt[4]).sum())
            c["tf32_differing_atom_peaks"] += int((mt[0] != mr[0]).sum())
            c["tf32_differing_bond_records"] += int((mt[3] != mr[3]).sum())
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=102400)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sparse", action="store_true", help="SparseHeadsPipeline instead of dense heads + PeakDecoder")
    ap.add_argument("--train-steps", type=int, default=0, help="train on pseudo-molecules first (rank 0) and broadcast the weights")
    ap.add_argument("--dump-subset", type=int, default=0, help="write weights + MOL blocks of the first M images to gpurun_out/")
    ap.add_argument("--weights", default="", help="load this state_dict (.pt) instead of training / random init")
    ap.add_argument("--act-dtype", default="bf16", choices=["bf16", "fp16"], help="UNet(act_dtype=...): storage format of the eval activations")
    ap.add_argument("--ref-check", action="store_true",
                    help="second pass: compare every image with the fp32 torch / cuDNN forward (TF32 off) of the same weights at record level")
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
    model = abcnet_b200.UNet(1, heads, act_dtype=args.act_dtype).to(dev)
    trained = args.train_steps > 0 or bool(args.weights)
    if trained:
        from synthdata import molecules
        P = 1024
        imgs, labels = molecules.pseudo_molecules(7, P, 512, 512)
        pool = torch.from_numpy(imgs).to(dev)
        if args.weights:
            model.load_state_dict(torch.load(args.weights, map_location=dev))
        else:
            model.load_state_dict(synthdata.make_state_dict(seed=11, variant="W0"))
            if rank == 0:
                train_on_pool(model, pool, labels, args.train_steps, dev)
            if world > 1:                              # one broadcast of the weights (multi_gpu_train2.py:89 does the same at start-up)
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0)
        model.eval()
    else:
        P = 64
        model.eval()
        model.load_state_dict(synthdata.make_state_dict(seed=0, variant="W1"))
        pool = torch.from_numpy(synthdata.binary_images(1000, P, 512, 512, 0.05)).to(dev)
        with torch.no_grad():                          # same calibration of the centre / omega biases as bench.py
            outs = model(pool[:8].contiguous())
            for k in (0, 4, 7):
                model.out_modules[k].conv2.bias += -1.0 - torch.quantile(outs[k].flatten()[:4_000_000].float(), 0.997)
    W = pool.shape[-1]
    cols = torch.arange(W, device=dev)

    def make_batch(a, b):
        g = torch.arange(a, b, device=dev)
        shift = (g // P) % W
        idx = (cols[None, :] - shift[:, None]) % W                                   # [n, W] source column of every output column
        x = pool[g % P]                                                              # [n, 1, H, W]
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
    subset = {}

    def finish(i):
        nonlocal n_mol
        a, b = ranges[i]
        sinks[i % 2].wait(B)
        texts = sinks[i % 2].molblocks(B, max(1, (os.cpu_count() or 1) // world))[:b - a]
        if a < args.dump_subset:                       # compact records of the subset, for the record-level oracle comparison
            dec = sinks[i % 2].dec if args.sparse else sinks[i % 2]
            recs = dec._parse(B)
            for g_, t_, r_ in zip(range(a, b), texts, recs):
                if g_ < args.dump_subset:
                    subset[g_] = (t_, r_[0].copy(), r_[1].copy(), r_[2])
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
    ref = None
    if args.ref_check:
        ref = ref_check(model, make_batch, ranges, B, dev)
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
        parts = [None] * world
        dist.all_gather_object(parts, (digests, n_mol))
        if ref is not None:
            refs = [None] * world
            dist.all_gather_object(refs, ref)
            ref = {k: sum(r[k] for r in refs) for k in ref}
    else:
        parts = [(digests, n_mol)]
    if rank == 0 and args.dump_subset:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                    "molblocks": [subset[g_][0] for g_ in range(min(args.dump_subset, total // world))],
                    "atoms": [subset[g_][1] for g_ in range(min(args.dump_subset, total // world))],
                    "bonds": [subset[g_][2] for g_ in range(min(args.dump_subset, total // world))],
                    "pool": P, "seed": 7, "trained": trained, "sparse": bool(args.sparse)},
                   os.path.join(ROOT, "gpurun_out", "shard_subset.pt"))
    if rank == 0:
        allb = sorted(d for p, _ in parts for d in p)
        h = hashlib.blake2b(digest_size=16)
        # the combined digest is over per-IMAGE-range digests; to be independent of where the shard boundaries fall the ranges
        # must be the same for every N: true when total / world is a multiple of the batch size
        for a, d in allb:
            h.update(f"{a}:{d};".encode())
        print(json.dumps({"tool": "shard_infer", "images": total, "n_gpus": world, "batch": B, "sparse_heads": bool(args.sparse),
                          "seconds": dt, "images_per_s": total / dt, "molecules": int(sum(n for _, n in parts)),
                          "batches": len(allb), "digest": h.hexdigest(), "weights": "trained on pseudo-molecules" if trained else "random init, calibrated", "act_dtype": args.act_dtype,
                          "subset_dumped": int(args.dump_subset), "ref_check_fp32_torch": ref,
                          "what": "forward + decode + native MOL-block assembly per image, sharded by contiguous image ranges, "
                                  "no data-path collective; digest over all MOL-block texts in global image order"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
