"""Deterministic pseudo-random tensors: moved to ``synthdata.detrand`` (shared with bench.py / tools); re-exported here."""
from synthdata.detrand import *  # noqa: F401,F403
from synthdata.detrand import integers, key, normalish, uniform  # noqa: F401
