"""Seeded synthetic PCM: identical inputs for the GPU path and its CPU checker."""
from __future__ import annotations

import numpy as np


def _rng(seed: int, *stream) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([0x5EED0000 + int(seed), *[int(s) for s in stream]]))


def uniform_pcm(seed: int, n_samples: int, over_range_frac: float = 0.01) -> np.ndarray:
    """uniform [-1, 1) with ~1 % of the samples pushed into +-[1, 4) to exercise clipping"""
    r = _rng(seed, 1)
    x = r.random(n_samples, dtype=np.float32) * np.float32(2.0) - np.float32(1.0)
    if over_range_frac > 0:
        m = r.random(n_samples, dtype=np.float32) < np.float32(over_range_frac)
        big = (r.random(n_samples, dtype=np.float32) * np.float32(3.0) + np.float32(1.0)) * np.sign(x).astype(np.float32)
        x = np.where(m, big, x).astype(np.float32)
    return x


def gains(seed: int, n: int, lo: float = 0.0, hi: float = 4.0) -> np.ndarray:
    return (_rng(seed, 2).random(n, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def tone_streams(seed: int, tick: int, n_streams: int, frames: int, channels: int, rate: int) -> np.ndarray:
    """sum of 3 sines (220 Hz, 1 kHz, 7 kHz) + -40 dB noise, per-stream random phase, phase-continuous
    across ticks. Returns float32 [n_streams, frames * channels] interleaved."""
    r0 = _rng(seed, 3)
    ph = r0.random((n_streams, 3, channels)) * 2 * np.pi
    t = (np.arange(frames, dtype=np.float64) + tick * frames) / rate
    out = np.zeros((n_streams, frames, channels), dtype=np.float64)
    for j, f in enumerate((220.0, 1000.0, 7000.0)):
        out += 0.3 * np.sin(2 * np.pi * f * t[None, :, None] + ph[:, j, None, :])
    noise = _rng(seed, 4, tick).standard_normal((n_streams, frames, channels)) * 0.01
    return (out + noise).astype(np.float32).reshape(n_streams, frames * channels)


def noise_streams(seed: int, tick: int, n_streams: int, frames: int, channels: int, amp: float = 0.5) -> np.ndarray:
    """cheap uniform noise for large benchmark batches (timing is data independent)"""
    r = _rng(seed, 5, tick)
    x = r.random((n_streams, frames * channels), dtype=np.float32)
    x -= np.float32(0.5)
    x *= np.float32(2.0 * amp)
    return x
