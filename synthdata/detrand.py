"""Deterministic, platform-exact pseudo-random arrays (test infrastructure).

A counter-based generator (splitmix64 finaliser over ``seed`` and the flat element
index) evaluated with numpy uint64 arithmetic, so the very same bits come out in the
build container and on the GPU box regardless of numpy/torch RNG versions. Floats are
built from the top 24 bits, hence exactly representable in fp32.
"""
from __future__ import annotations

import zlib

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def key(*parts) -> int:
    """Stable 63-bit key from strings / ints (crc32 based, no Python hash())."""
    h = 0x1234567
    for p in parts:
        b = p.encode() if isinstance(p, str) else int(p).to_bytes(8, "little", signed=True)
        h = (h * 1000003 + zlib.crc32(b)) & 0x7FFFFFFFFFFFFFFF
    return h


def bits(seed: int, shape, offset: int = 0) -> np.ndarray:
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * _G + np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * _M2 + _G
    return _mix(_mix(z)).reshape(shape)


def uniform(seed: int, shape, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """float32 uniform in [lo, hi) with 24-bit resolution."""
    u = (bits(seed, shape) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))
    return (u * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def normalish(seed: int, shape, std: float = 1.0) -> np.ndarray:
    """Sum of 4 uniforms, centred and scaled to unit variance (Irwin-Hall) -- float32."""
    acc = np.zeros(shape, np.float32)
    for k in range(4):
        acc += uniform(seed * 4 + k + 17, shape)
    return ((acc - np.float32(2.0)) * np.float32(std * 1.7320508)).astype(np.float32)


def integers(seed: int, shape, lo: int, hi: int) -> np.ndarray:
    """int64 in [lo, hi)."""
    return (bits(seed, shape) >> np.uint64(11)).astype(np.int64) % (hi - lo) + lo
