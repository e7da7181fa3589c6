"""The reference's optimiser -- ``optim.Adam(model.parameters(), lr=2.5e-4, weight_decay=1e-8)`` of
``/root/reference/src/train.py:55`` (re-created with lr 2.5e-5 at ``:84-85``) -- as ONE kernel launch over all parameter
tensors (``abc_adam_step``), with the step counter on the device so that the launch is CUDA-graph replayable.

Same update rule as ``torch.optim.Adam`` (L2 decay added to the gradient, bias-corrected first / second moments,
``amsgrad=False``); ``state_dict()`` / ``param_groups`` keep torch's layout (``exp_avg``, ``exp_avg_sq``, ``step``).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=2.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, capturable=True))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdam: one parameter group (the reference uses one, train.py:55)")
        self._key = None
        self._hyper_host = None
        self._step = None

    def _hyper(self):
        g = self.param_groups[0]
        lr = g["lr"]
        # a CUDA tensor lr (as torch's capturable Adam accepts) is re-read on the device by every step / graph replay
        return (-1.0 if torch.is_tensor(lr) and lr.is_cuda else float(lr), float(g["betas"][0]), float(g["betas"][1]),
                float(g["eps"]), float(g["weight_decay"]))

    def _build(self, params):
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("abcnet_b200.FusedAdam needs CUDA parameters (no CPU fallback)")
        _lib.require_device()
        if self._step is None:
            self._step = torch.zeros((), dtype=torch.float32, device=dev)
            self._hyper_dev = torch.zeros(5, dtype=torch.float32, device=dev)
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                raise ValueError("FusedAdam: parameters and gradients must be contiguous fp32 tensors")
            st = self.state[p]
            old = st.get("step")
            if torch.is_tensor(old) and old is not self._step:        # state loaded through load_state_dict: adopt its counter
                self._step.copy_(old.to(self._step.device, torch.float32).reshape(()))
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] = self._step                                   # one shared device counter
        chunk = int(lib.abc_adam_chunk_elems())
        chunks = [(i, c) for i, p in enumerate(params) for c in range((p.numel() + chunk - 1) // chunk)]

        def table(ptrs):
            return torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self._tab = (table([p.data_ptr() for p in params]), table([p.grad.data_ptr() for p in params]),
                     table([self.state[p]["exp_avg"].data_ptr() for p in params]),
                     table([self.state[p]["exp_avg_sq"].data_ptr() for p in params]),
                     table([p.numel() for p in params]), torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous())
        self._n_chunks = len(chunks)

    @torch.no_grad()
    def reset_state(self, lr=None):
        """Zero the moments and the step counter IN PLACE (graph-safe: a captured TrainStep keeps its pointers) and
        optionally set a new learning rate. This is what the reference's LR drop does: ``train.py:84-85`` re-creates
        ``optim.Adam`` with lr 2.5e-5 at epoch ``epoch_nums / 3``, which discards the moments and restarts the bias correction."""
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k].zero_()
        if self._step is not None:
            self._step.zero_()
        if lr is not None:
            g = self.param_groups[0]
            if torch.is_tensor(g["lr"]):
                g["lr"].fill_(float(lr))
            else:
                g["lr"] = float(lr)
            self.sync_hyperparams()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """torch's loader replaces the state tensors; a captured graph (and the device pointer tables) still reference the
        old ones. The loaded values are therefore copied INTO the existing buffers, which stay the optimiser's state."""
        old = {p: dict(st) for p, st in self.state.items()}
        super().load_state_dict(state_dict)
        for p, st in self.state.items():
            prev = old.get(p)
            if not prev:
                continue
            for k in ("exp_avg", "exp_avg_sq"):
                if k in prev and k in st and st[k] is not prev[k]:
                    prev[k].copy_(st[k])
                    st[k] = prev[k]
            new_step = st.get("step")
            if self._step is not None and new_step is not None and new_step is not self._step:
                self._step.copy_(torch.as_tensor(new_step, dtype=torch.float32).reshape(()))
                st["step"] = self._step
        self._hyper_host = None                                       # lr / betas may have changed: re-push before the next step

    def sync_hyperparams(self):
        """Push lr / betas / eps / weight_decay of ``param_groups[0]`` to the device if they changed (call before a CUDA-graph
        replay after editing ``param_groups``; ``step()`` does it itself)."""
        h = self._hyper()
        if h != self._hyper_host and self._step is not None:
            self._hyper_dev.copy_(torch.tensor(h, dtype=torch.float32))
            self._hyper_host = h

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params = [p for p in self.param_groups[0]["params"] if p.grad is not None]
        if not params:
            return loss
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr() if "exp_avg" in self.state[p] else 0,
                     self.state[p].get("step") is self._step) for p in params)
        if key != self._key:
            self._build(params)
            self._key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), True) for p in params)
        self.sync_hyperparams()
        lr = self.param_groups[0]["lr"]
        if torch.is_tensor(lr) and lr.is_cuda:
            self._hyper_dev[0:1].copy_(lr.detach().reshape(1).to(torch.float32))      # device -> device: part of a captured graph
        t = self._tab
        check(lib.abc_adam_step(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(),
                                self._n_chunks, self._hyper_dev.data_ptr(), self._step.data_ptr(), _lib.current_stream_ptr()),
              "abc_adam_step")
        return loss
