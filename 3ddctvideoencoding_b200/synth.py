"""Seeded synthetic raw-grayscale clips.

Raw video layout is the reference's: frames x height x width unsigned bytes,
frame-major, no header (J/Encoder.java:47-56, C/encoder.c:21-35).
"""
from __future__ import annotations

import numpy as np


def natural(width: int, height: int, frames: int, seed: int = 1) -> np.ndarray:
    """g(x,y,t) = 128 + 60 sin((x+3t)/37) + 50 cos((y-2t)/23) + N(0, 6), rounded, clipped to [0,255]."""
    rng = np.random.default_rng(seed)
    t = np.arange(frames, dtype=np.float64)[:, None, None]
    y = np.arange(height, dtype=np.float64)[None, :, None]
    x = np.arange(width, dtype=np.float64)[None, None, :]
    g = 128.0 + 60.0 * np.sin((x + 3.0 * t) / 37.0) + 50.0 * np.cos((y - 2.0 * t) / 23.0)
    g = g + rng.normal(0.0, 6.0, size=(frames, height, width))
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def noise(width: int, height: int, frames: int, seed: int = 2) -> np.ndarray:
    """i.i.d. uniform [0,255]: worst case for the entropy coder (~3.3 bit/sample)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(frames, height, width), dtype=np.uint8)


def constant(width: int, height: int, frames: int, value: int = 128) -> np.ndarray:
    """All-DC clip: one long code + (cube size - 1) one-bit zeros per cube."""
    return np.full((frames, height, width), value, np.uint8)


def natural_slab(width: int, height: int, cube: int, slab: int, seed: int = 1) -> np.ndarray:
    """Slab `slab` (frames slab*cube .. slab*cube+cube-1) of the natural clip, seeded PER SLAB: any slab of an
    arbitrarily long clip can be made on its own, so sampled slabs of the big BASELINE configurations can be checked
    against the CPU oracle without generating the whole clip."""
    rng = np.random.default_rng(seed * 1000003 + slab)
    t = np.arange(slab * cube, (slab + 1) * cube, dtype=np.float64)[:, None, None]
    y = np.arange(height, dtype=np.float64)[None, :, None]
    x = np.arange(width, dtype=np.float64)[None, None, :]
    g = 128.0 + 60.0 * np.sin((x + 3.0 * t) / 37.0) + 50.0 * np.cos((y - 2.0 * t) / 23.0)
    g = g + rng.normal(0.0, 6.0, size=(cube, height, width))
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def natural_slabs(width: int, height: int, cube: int, slab_lo: int, slab_hi: int, seed: int = 1) -> np.ndarray:
    return np.concatenate([natural_slab(width, height, cube, s, seed) for s in range(slab_lo, slab_hi)])
