"""Seeded synthetic signals / filters shared by tests and bench.py.

uniform(-1, 1) matches the reference tests' ``rand_m1_1``
(``/root/reference/framework/modules/saf_utilities/saf_utility_misc.c:225-234``); the
exponentially decaying noise RIR is the "realistic" long filter used for configs C4/C5.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5AF0C0DE


def uniform(shape, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=shape).astype(np.float32)


def decaying_rir(shape, seed: int = SEED, db_at_end: float = -60.0) -> np.ndarray:
    """h[..., k] = u[k] * exp(-6.9 k / L)  (-60 dB at the last tap by default)."""
    L = shape[-1]
    u = uniform(shape, seed)
    rate = -np.log(10.0 ** (db_at_end / 20.0))
    env = np.exp(-rate * np.arange(L, dtype=np.float64) / L).astype(np.float32)
    return (u * env).astype(np.float32)


# BASELINE.json configs (SURVEY.md §8 notation)
CONFIGS = {
    "C1": dict(kind="matrix", nIn=4, nOut=2, hop=256, L=1024),
    "C2": dict(kind="matrix", nIn=25, nOut=2, hop=128, L=512),
    "C3": dict(kind="multi", nCH=256, hop=512, L=4096),
    "C4": dict(kind="matrix", nIn=64, nOut=64, hop=1024, L=96000),
    "UT": dict(kind="matrix", nIn=32, nOut=40, hop=2048, L=512),
}
