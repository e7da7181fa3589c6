"""Host-side logic of the N>1 paths, exercised with world_size 2 on the gloo backend (CPU only)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker_shard(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from abcnet_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    total = 11
    lo, hi = shard.shard_bounds(total, rank, world)
    local = [f"img{i}:rank{rank}" for i in range(lo, hi)]
    out = shard.gather_results(local, total)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def _worker_buckets(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from abcnet_b200.ddp import GradBuckets
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((7, 3), (5,), (64, 64), (11,), (2, 2, 2))]
    gb = GradBuckets(params, bucket_bytes=4 * 100)
    nb = len(gb.buckets)
    gb.zero()
    for p in gb.params:                     # "backward": gradients appear in reverse order
        p.grad.copy_(torch.full_like(p, float(rank + 1)) * p.numel())
        gb.grad_ready(p)
    gb.finish()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5) * p.numel()) for p in params)
    views = all(p.grad.data_ptr() >= gb.buckets[gb.bucket_of[id(p)]].data_ptr() for p in params)
    if rank == 0:
        q.put((ok, views, nb))
    dist.destroy_process_group()


def _run(fn, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=fn, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_bounds_cover_everything():
    from abcnet_b200 import shard
    for total in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert list(shard.batches(3, 10, 4)) == [(3, 7), (7, 10)]
    with pytest.raises(ValueError):
        shard.shard_bounds(5, 2, 2)


def test_sharded_inference_gather_world2():
    out = _run(_worker_shard, 29611)
    assert out == [f"img{i}:rank{0 if i < 6 else 1}" for i in range(11)]


def test_bucketed_allreduce_world2():
    ok, views, nb = _run(_worker_buckets, 29612)
    assert ok and views and nb >= 2


def test_grad_buckets_tail_bucket_holds_the_last_gradients():
    """The parameters whose gradients appear last in the backward pass (first registered) get a small bucket of their own, so
    that the last large bucket can be reduced while they are still being computed (abcnet_b200/ddp.py, tail_bytes)."""
    from abcnet_b200.ddp import GradBuckets
    shapes = [(10,), (16, 1, 3, 3), (16,), (16, 16, 3, 3), (256, 256, 3, 3), (512, 256, 3, 3), (128, 128, 3, 3), (360, 128, 1, 1)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    gb = GradBuckets(params, bucket_bytes=4 << 20, tail_bytes=16 << 10)
    tail = gb.members[-1]
    assert [tuple(p.shape) for p in tail] == [(16, 16, 3, 3), (16,), (16, 1, 3, 3), (10,)]      # reverse registration order, <= 16 KB
    assert sum(p.numel() for p in tail) * 4 <= 16 << 10
    assert all(id(p) in gb.bucket_of for p in params) and sum(len(m) for m in gb.members) == len(params)
    # every .grad is a view into its bucket, contiguous and in order
    for b, ms in zip(gb.buckets, gb.members):
        off = 0
        for p in ms:
            assert p.grad.data_ptr() == b.data_ptr() + 4 * off
            off += p.numel()
    # a model without small trailing parameters: no empty / degenerate tail bucket
    gb2 = GradBuckets([torch.nn.Parameter(torch.zeros(1 << 20)) for _ in range(3)], bucket_bytes=4 << 20, tail_bytes=16 << 10)
    assert all(len(m) >= 1 for m in gb2.members) and sum(len(m) for m in gb2.members) == 3
