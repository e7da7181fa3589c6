"""Image sharding for multi-GPU inference: one process per GPU, contiguous image ranges, NO data-path collective.

Replaces the roles of ``nn.DataParallel`` in ``/root/reference/src/img2smiles.py:43`` (scatter images / gather 32.8 MB
of dense maps per image to GPU 0) and of the ``multiprocessing.Pool`` fan-out in
``/root/reference/src/multi_proc_img2smiles.py:268,299-309``: every rank runs forward + decode on its own shard and
only the compact per-image results (peak records or strings, tens of kB per image) are gathered on the host.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) range of rank ``rank``; the first ``total % world`` ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batches(lo: int, hi: int, batch: int):
    """Yield [a, b) batch ranges covering [lo, hi)."""
    for a in range(lo, hi, batch):
        yield a, min(a + batch, hi)


def gather_results(local: Sequence, total: int, group=None) -> List:
    """Host-side gather of per-image results to every rank, in global image order (the only cross-rank step of the
    inference path). ``local`` must hold exactly the items of this rank's ``shard_bounds`` range."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert len(local) == total
        return list(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(total, rank, world)
    if len(local) != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} results, got {len(local)}")
    parts = [None] * world
    dist.all_gather_object(parts, list(local), group=group)
    out = []
    for p in parts:
        out.extend(p)
    return out
