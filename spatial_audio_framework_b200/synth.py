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


# ---- inputs of the filter producers (include/safconv_b200.h, "Filter producers") -------------------------------------
def fibonacci_grid_deg(n_dirs: int) -> np.ndarray:
    """n_dirs nearly uniform directions on the sphere, [azimuth, elevation] in degrees, float32."""
    i = np.arange(n_dirs) + 0.5
    el = np.degrees(np.arcsin(1.0 - 2.0 * i / n_dirs))
    az = np.degrees((np.pi * (1.0 + 5.0 ** 0.5) * i) % (2.0 * np.pi)) - 180.0
    return np.stack([az, el], 1).astype(np.float32)


def synthetic_hrtfs(n_dirs: int, fft_size: int, fs: float = 48000.0, seed: int = SEED):
    """A head-like HRTF set for the decoder design: per ear an interaural delay and level difference that follow the
    direction, a gentle roll-off and a little direction-dependent ripple.  Returns (hrtfs [fft_size/2+1, 2, n_dirs]
    complex64, dirs_deg [n_dirs, 2] float32, itd_s [n_dirs] float32)."""
    rng = np.random.default_rng(seed)
    dirs = fibonacci_grid_deg(n_dirs)
    az, el = np.radians(dirs[:, 0].astype(np.float64)), np.radians(dirs[:, 1].astype(np.float64))
    uy = np.cos(el) * np.sin(az)                                   # +y = left
    itd = (0.0875 / 343.0) * uy
    nb = fft_size // 2 + 1
    f = np.arange(nb) * fs / fft_size
    H = np.zeros((nb, 2, n_dirs), np.complex64)
    ripple = 0.05 * rng.standard_normal((2, 4))
    for e, s in enumerate((+1.0, -1.0)):
        tau = 0.0003 - s * itd / 2.0
        g = 1.0 + 0.4 * s * uy
        rp = 1.0 + sum(ripple[e, k] * np.cos((k + 1) * az + 0.3 * k) * np.cos(el) for k in range(4))
        H[:, e, :] = (g * rp)[None, :] * np.exp(-2j * np.pi * f[:, None] * tau[None, :]) / (1.0 + (f[:, None] / 16000.0) ** 2)
    return H, dirs, itd.astype(np.float32)


# wall absorption per octave band (7 bands x 6 walls) of the reference's own unit test, test__reverb_module.c:33-39
IMS_TEST_ABS_WALL = np.array(
    [[0.180791250, 0.207307300, 0.134990800, 0.229002250, 0.212128400, 0.241055000],
     [0.225971250, 0.259113700, 0.168725200, 0.286230250, 0.265139600, 0.301295000],
     [0.258251250, 0.296128100, 0.192827600, 0.327118250, 0.303014800, 0.344335000],
     [0.301331250, 0.345526500, 0.224994001, 0.381686250, 0.353562000, 0.401775000],
     [0.361571250, 0.414601700, 0.269973200, 0.457990250, 0.424243600, 0.482095000],
     [0.451931250, 0.518214500, 0.337442000, 0.572446250, 0.530266000, 0.602575000],
     [0.602591250, 0.690971300, 0.449934800, 0.763282250, 0.707040400, 0.803455000]], np.float32)
IMS_TEST_ROOM = np.array([10.0, 7.0, 3.0], np.float32)             # test__reverb_module.c:52
