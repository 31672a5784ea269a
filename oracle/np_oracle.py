"""numpy/pure-Python restatement of the hot path -- an INDEPENDENT second implementation used only to
cross-check the C oracle (two implementations written separately agreeing bit-for-bit is the only pin
available for the resampler and s16 conversion, see sk_oracle.h). TEST INFRASTRUCTURE ONLY.

Reference citations as in sk_oracle.h.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def gain(x, g):
    """gain.rs:187-189"""
    return (np.asarray(x, dtype=F32) * F32(g)).astype(F32)


def f32_to_s16(x):
    """build-defined: sat_s16(rint_half_even(x * 32768)), NaN -> 0"""
    y = np.asarray(x, dtype=F32) * F32(32768.0)
    r = np.rint(y.astype(np.float64))  # f32 -> f64 exact; rint = half-to-even
    r = np.where(np.isnan(r), 0.0, r)
    return np.clip(r, -32768.0, 32767.0).astype(np.int16)


def s16_to_f32(s):
    return (np.asarray(s, dtype=np.int16).astype(F32) * F32(1.0 / 32768.0)).astype(F32)


def mix_order(frames, out_channels, out_size):
    """mixer.rs:960-980. frames: list of (samples, channels, unique). Returns (order, has_base)."""
    cand = [(bool(f[2]) if len(f) > 2 else True, i) for i, f in enumerate(frames)
            if f[1] == out_channels and len(f[0]) == out_size]
    if not cand:
        return list(range(len(frames))), False
    base = max(cand)[1]
    vec = list(range(len(frames)))
    last = vec.pop()          # swap_remove
    if base < len(vec):
        vec[base] = last
    return [base] + vec, True


def mix_into(out, src, sc, oc):
    """mixer.rs:1027-1078 (in place on float32 array `out`)."""
    src = np.asarray(src, dtype=F32)
    spc = len(src) // sc
    ospc = len(out) // oc
    m = min(spc, ospc)
    if sc == oc:
        n = m * oc
        out[:n] = out[:n] + src[:n]
    elif sc == 1 and oc == 2:
        out[0:2 * m:2] = out[0:2 * m:2] + src[:m]
        out[1:2 * m:2] = out[1:2 * m:2] + src[:m]
    elif sc == 2 and oc == 1:
        out[:m] = out[:m] + (src[0:2 * m:2] + src[1:2 * m:2]) * F32(0.5)
    else:
        for ch in range(oc):
            out[ch:m * oc:oc] = out[ch:m * oc:oc] + src[(ch % sc):m * sc:sc]
    return out


def mix(frames, out_channels, out_size):
    order, has_base = mix_order(frames, out_channels, out_size)
    if has_base:
        out = np.array(frames[order[0]][0], dtype=F32, copy=True)
        rest = order[1:]
    else:
        out = np.zeros(out_size, dtype=F32)
        rest = order
    for i in rest:
        mix_into(out, frames[i][0], frames[i][1], out_channels)
    return out


def mix_sync(frames, max_channels_seen=0):
    """mixer.rs:944-1013"""
    oc = max([max_channels_seen, 1] + [f[1] for f in frames])
    spc = max([len(f[0]) // f[1] for f in frames] + [0])
    return mix(frames, oc, spc * oc), oc


class FastFixedIn:
    """rubato 0.16.2 FastFixedIn<f32>, Linear. Phase chain in Python floats (IEEE f64), data in numpy f32."""

    def __init__(self, in_rate, out_rate, chunk, channels):
        self.ratio = float(out_rate) / float(in_rate)
        self.chunk = chunk
        self.channels = channels
        self.last_index = -4.0
        self.buf = np.zeros((chunk + 16, channels), dtype=F32)

    def process(self, chunk_interleaved):
        N, C = self.chunk, self.channels
        x = np.asarray(chunk_interleaved, dtype=F32).reshape(N, C)
        self.buf[:16] = self.buf[N:N + 16].copy()
        self.buf[16:] = x
        t = 1.0 / self.ratio
        end_idx = N - 9 - math.ceil(t)
        idx = self.last_index
        pos, frac = [], []
        while idx < float(end_idx):
            idx += t
            fl = math.floor(idx)
            pos.append(int(fl) + 16)
            frac.append(idx - fl)
        self.last_index = idx - float(N)
        if not pos:
            return np.zeros(0, dtype=F32)
        p = np.asarray(pos)
        f = np.asarray(frac, dtype=np.float64).astype(F32)[:, None]
        y0, y1 = self.buf[p], self.buf[p + 1]
        out = ((F32(1.0) - f) * y0).astype(F32) + (f * y1).astype(F32)
        return out.astype(F32).reshape(-1)


def sinc_taps(L, O, fc):
    """tap table of the sinc spec (oracle/sk_sinc.c header), numpy f64 -> f32"""
    p = np.arange(O + 1, dtype=np.float64)[:, None]
    n = np.arange(L, dtype=np.float64)[None, :]
    tau = L / 2.0 - 1.0 - n + p / O
    z = fc * tau
    s = np.where(z == 0.0, 1.0, np.sin(np.pi * z) / np.where(z == 0.0, 1.0, np.pi * z))
    u = (tau + L / 2.0) / L
    w = 0.35875 - 0.48829 * np.cos(2 * np.pi * u) + 0.14128 * np.cos(4 * np.pi * u) - 0.01168 * np.cos(6 * np.pi * u)
    w = np.where((u <= 0.0) | (u >= 1.0), 0.0, w * w)
    g = fc * s * w
    return (g / g.sum(axis=1, keepdims=True)).astype(F32)


class SincFixedIn:
    """numpy restatement of the sinc spec; dot products in f64 (not the sequential f32 fma of the C oracle): agrees to ~1e-6"""

    def __init__(self, in_rate, out_rate, chunk, channels, sinc_len=64, oversampling=256, f_cutoff=0.95):
        self.ratio = float(out_rate) / float(in_rate)
        self.N, self.C, self.L, self.O, self.H = chunk, channels, sinc_len, oversampling, sinc_len + 8
        self.last_index = -float(sinc_len // 2)
        self.taps = sinc_taps(sinc_len, oversampling, f_cutoff * min(1.0, self.ratio)).astype(np.float64)
        self.buf = np.zeros((self.H + chunk, channels), dtype=F32)

    def process(self, chunk_interleaved):
        N, C, L, H = self.N, self.C, self.L, self.H
        self.buf[:H] = self.buf[N:N + H].copy()
        self.buf[H:] = np.asarray(chunk_interleaved, dtype=F32).reshape(N, C)
        t = 1.0 / self.ratio
        end_idx = N - L // 2 - 1 - math.ceil(t)
        idx = self.last_index
        out = []
        b64 = self.buf.astype(np.float64)
        while idx < float(end_idx):
            idx += t
            fl = math.floor(idx)
            fo = (idx - fl) * self.O
            p = min(int(math.floor(fo)), self.O - 1)
            q = float(F32(fo - p))
            base = int(fl) - L // 2 + 1 + H
            win = b64[base:base + L]
            y0 = win.T @ self.taps[p]
            y1 = win.T @ self.taps[p + 1]
            out.append((1.0 - q) * y0 + q * y1)
        self.last_index = idx - float(N)
        return np.asarray(out, dtype=F32).reshape(-1) if out else np.zeros(0, dtype=F32)
