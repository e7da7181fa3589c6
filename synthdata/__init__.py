"""Deterministic synthetic inputs and weights (no reference arithmetic, no product code): shared by tests, bench.py and the
profiling tools so that none of them needs the oracle for anything but checking."""
from . import detrand, inputs, weights  # noqa: F401
from .inputs import binary_images, dense_targets  # noqa: F401
from .weights import V2_HEADS, make_state_dict, param_shapes  # noqa: F401
