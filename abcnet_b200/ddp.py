"""Bucketed gradient all-reduce for data-parallel training (one process per GPU, NCCL over NVLink / NVSwitch).

Replaces ``DistributedDataParallel`` of ``/root/reference/src/multi_gpu_train2.py:89`` (and ``multi_gpu_train.py:52``):
all 10 698 575 gradients live in a few contiguous fp32 buckets laid out in REVERSE layer order (heads first), each
parameter's ``.grad`` is a view into its bucket so that the wgrad kernels write straight into it, and a bucket's
all-reduce (mean) is enqueued on a dedicated communication stream as soon as the last gradient of the bucket has
been produced -- overlapping NCCL with the rest of the backward pass. BatchNorm statistics stay per replica, exactly
as in the reference (no SyncBN).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 12 << 20, group=None, tail_bytes: int = 128 << 10):
        """``tail_bytes``: the gradients produced LAST in the backward pass (the full-resolution stem layers: a few thousand
        parameters, but ~6 ms of backward time at 512 x 512) get a bucket of their own of at most this size, so that the last
        large bucket is already on the wire while they are computed and only a latency-sized all-reduce remains exposed after
        the backward pass (measured at N = 8 before this split: 0.75 ms exposed per step, profiles/r02_bench_n8.json)."""
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        self.group = group
        self.comm_enabled = True                         # False: buckets only (measurement of the step without the exchange)
        self.params = list(reversed(params))             # backward produces gradients in reverse registration order
        dev = self.params[0].device
        self.buckets: List[torch.Tensor] = []
        self.members: List[List[torch.nn.Parameter]] = []
        cur, cur_n = [], 0
        limit = max(1, bucket_bytes // 4)
        tail_start, acc = len(self.params), 0
        while tail_start > 1 and (acc + self.params[tail_start - 1].numel()) * 4 <= tail_bytes:
            tail_start -= 1
            acc += self.params[tail_start].numel()
        if tail_start == len(self.params):
            tail_start = -1                              # no parameter fits: no separate tail bucket
        for i, p in enumerate(self.params):
            if cur and (cur_n + p.numel() > limit or i == tail_start):
                self._close(cur, cur_n, dev)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        self._close(cur, cur_n, dev)
        self.bucket_of = {id(p): b for b, ms in enumerate(self.members) for p in ms}
        self._pending = [len(ms) for ms in self.members]
        self._works = []
        self._launched = False
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def _close(self, ps, n, dev):
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)     # .grad is a view: kernels write into the bucket directly
            off += p.numel()
        self.buckets.append(flat)
        self.members.append(list(ps))

    def zero(self):
        for b in self.buckets:
            b.zero_()
        self._pending = [len(ms) for ms in self.members]
        self._works = []
        self._launched = False

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def grad_ready(self, p: torch.nn.Parameter):
        """Call when the gradient of ``p`` has been written (enqueued) -- launches the bucket's all-reduce when complete."""
        b = self.bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int):
        if self.world() == 1 or not self.comm_enabled:
            return
        flat = self.buckets[b]
        self._launched = True
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                flat.div_(self.world())
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:                                                   # CPU tensors (gloo) -- host-side logic tests
            flat.div_(self.world())
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Flush buckets that never completed (unused parameters) and make the compute stream wait for the reductions."""
        for b, left in enumerate(self._pending):
            if left > 0:
                self._pending[b] = 0
                self._launch(b)
        for w in self._works:
            w.wait()
        # join the communication stream only if something was forked to it since zero(): inside a CUDA-graph capture a wait on a
        # stream that is not part of the capture is an error
        if self.comm_stream is not None and self._launched:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._works = []
