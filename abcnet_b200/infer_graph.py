"""Inference + decode of one fixed batch shape as ONE replayable CUDA graph.

The eval forward is ~45 kernel launches and the decoder one more; at the reference's own inference batch sizes
(``/root/reference/src/img2smiles.py:39``: 32 images; BASELINE configs[0]: 8) the device work is a millisecond or two, of the
same order as the host time needed to enqueue those launches one by one. ``InferGraph`` captures

    outs = model.infer(x, layout="p8f"); decoder.launch(outs)

once per (weights generation, input shape) and replays it with a single ``cudaGraphLaunch``. Results are bit-identical to the
eager calls (same kernels, same launch parameters). No CPU fallback: a CUDA model and CUDA inputs are required.
"""
from __future__ import annotations

import torch

from . import _lib
from .decode import PeakDecoder


class InferGraph:
    """g = InferGraph(model, batch, H, W, atom_cap=1024, bond_cap=4096, dtype=torch.uint8)
    recs = g(x)                  # x [batch,1,H,W] on the GPU (or pinned host memory: copied asynchronously); list of records
    n = g.launch(x); ...; recs = g.fetch(n)        # split form: enqueue without synchronising, fetch later
    ``g.outs`` are the eight (planar-8 fp32) maps of the last replay, ``g.decoder`` the PeakDecoder that owns the records."""

    def __init__(self, model, batch, H, W, atom_cap=1024, bond_cap=4096, dtype=torch.float32, thr=-1.0, omega_mode="nms"):
        dev = model.s.device
        if dev.type != "cuda":
            raise RuntimeError("abcnet_b200.InferGraph needs the model on a CUDA (sm_100) device; there is no CPU path")
        _lib.require_device()
        if model.training:
            raise RuntimeError("InferGraph captures the eval forward: call model.eval() first")
        self.model, self.thr, self.omega_mode = model, float(thr), omega_mode
        self.x = torch.zeros((batch, model.n_channels, H, W), dtype=dtype, device=dev)
        self.decoder = PeakDecoder(batch, atom_cap=atom_cap, bond_cap=bond_cap, device=dev)
        self.outs = None
        self.graph = None
        self._gen = -1
        self.launches_per_replay = 0

    def _capture(self):
        m = self.model
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up outside capture: weight packing, buffer allocation
            for _ in range(2):
                self.outs = m.infer(self.x, self.outs, layout="p8f")
                self.decoder.launch(self.outs, self.thr, self.omega_mode)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            m.infer(self.x, self.outs, layout="p8f")
            self.decoder.launch(self.outs, self.thr, self.omega_mode)
        self.launches_per_replay = _lib.launch_count() - l0
        self._gen = m._pack_gen

    @torch.no_grad()
    def launch(self, x):
        m = self.model
        if m.training:
            raise RuntimeError("InferGraph: the model is in train() mode")
        if x.shape != self.x.shape:
            raise ValueError(f"InferGraph was captured for inputs of shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        # the graph holds pointers into the packed weights: re-capture when they were re-packed (or never were)
        if self.graph is None or m._packed is None or m._packed_key != m._param_key() or self._gen != m._pack_gen:
            if m._packed is None or m._packed_key != m._param_key():
                m.prepare()
            self._capture()
        self.graph.replay()
        return self.x.shape[0]

    def fetch(self, n):
        return self.decoder.fetch(n)

    def __call__(self, x):
        return self.fetch(self.launch(x))
