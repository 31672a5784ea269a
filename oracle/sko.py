"""ctypes wrapper of the C oracle (oracle/libsk_oracle.so). TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
Nothing under streamkit_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsk_oracle.so")
REF_GAIN_PLUGIN = os.path.join(_HERE, "_ref", "libgain_plugin_c.so")


class Frame(C.Structure):
    _fields_ = [("samples", C.POINTER(C.c_float)), ("n_samples", C.c_uint32), ("channels", C.c_uint16), ("unique", C.c_uint8),
                ("_pad", C.c_uint8)]


class PacketMeta(C.Structure):
    _fields_ = [("timestamp_us", C.c_uint64), ("has_timestamp", C.c_uint8), ("duration_us", C.c_uint64), ("sequence", C.c_uint64)]


EMIT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint16, C.POINTER(C.c_float), C.c_size_t, C.POINTER(PacketMeta))

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(_HERE, f) for f in ("sk_oracle.c", "sk_chain.c", "sk_sinc.c", "sk_oracle.h")]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs):
        build()
    lib = C.CDLL(LIB_PATH)
    f32p, i16p, vp = C.POINTER(C.c_float), C.POINTER(C.c_int16), C.c_void_p
    lib.sko_gain_validate.restype = C.c_int
    lib.sko_gain_validate.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
    lib.sko_gain_apply.argtypes = [vp, C.c_size_t, C.c_float]
    lib.sko_f32_to_s16_buf.argtypes = [vp, vp, C.c_size_t]
    lib.sko_s16_to_f32_buf.argtypes = [vp, vp, C.c_size_t]
    lib.sko_gain_f32_to_s16_buf.argtypes = [vp, vp, C.c_size_t, C.c_float]
    lib.sko_mix_plan.argtypes = [C.POINTER(Frame), C.c_size_t, C.c_uint16, C.c_size_t, vp, C.POINTER(C.c_int)]
    lib.sko_mix_sync.restype = C.c_int
    lib.sko_mix_sync.argtypes = [C.POINTER(Frame), C.c_size_t, C.c_uint16, vp, C.c_size_t, C.POINTER(C.c_uint16), C.POINTER(C.c_size_t)]
    lib.sko_mix_clocked.restype = C.c_int
    lib.sko_mix_clocked.argtypes = [C.POINTER(Frame), C.c_size_t, C.c_uint16, C.c_size_t, vp, C.c_size_t]
    lib.sko_ffi_new.restype = vp
    lib.sko_ffi_new.argtypes = [C.c_double, C.c_size_t, C.c_size_t]
    lib.sko_ffi_free.argtypes = [vp]
    lib.sko_ffi_out_max.restype = C.c_size_t
    lib.sko_ffi_out_max.argtypes = [vp]
    lib.sko_ffi_process_interleaved.restype = C.c_size_t
    lib.sko_ffi_process_interleaved.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sko_ffi_last_index.restype = C.c_double
    lib.sko_ffi_last_index.argtypes = [vp]
    lib.sko_ffi_history.argtypes = [vp, vp]
    lib.sko_rsnode_new.restype = vp
    lib.sko_rsnode_new.argtypes = [C.c_uint32, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.sko_rsnode_free.argtypes = [vp]
    lib.sko_rsnode_push.restype = C.c_int
    lib.sko_rsnode_push.argtypes = [vp, C.c_uint32, C.c_uint16, vp, C.c_size_t, C.c_int, C.c_uint64, EMIT_FN, vp, C.c_char_p, C.c_size_t]
    lib.sko_rsnode_finish.restype = C.c_int
    lib.sko_rsnode_finish.argtypes = [vp, EMIT_FN, vp]
    lib.sko_duration_us_for_frames.restype = C.c_uint64
    lib.sko_duration_us_for_frames.argtypes = [C.c_uint32, C.c_size_t]
    lib.sko_chain_bench.restype = C.c_double
    lib.sko_chain_bench.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint16, vp, C.c_uint32, vp, vp, C.c_int, vp,
                                    C.POINTER(C.c_uint64)]
    lib.sko_sinc_new.restype = vp
    lib.sko_sinc_new.argtypes = [C.c_double, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double]
    lib.sko_sinc_free.argtypes = [vp]
    lib.sko_sinc_out_max.restype = C.c_size_t
    lib.sko_sinc_out_max.argtypes = [vp]
    lib.sko_sinc_last_index.restype = C.c_double
    lib.sko_sinc_last_index.argtypes = [vp]
    lib.sko_sinc_process_interleaved.restype = C.c_size_t
    lib.sko_sinc_process_interleaved.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sko_sinc_taps.argtypes = [C.c_size_t, C.c_size_t, C.c_double, vp]
    lib.sko_node_bench.restype = C.c_double
    lib.sko_node_bench.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, C.c_int, vp, vp, vp,
                                   C.c_uint32]
    _lib = lib
    return lib


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ---------------------------------------------------------------- gain / s16

def gain_validate(gain: float):
    err = C.create_string_buffer(256)
    rc = load().sko_gain_validate(C.c_float(gain), err, 256)
    return rc, err.value.decode()


def gain(x: np.ndarray, g: float) -> np.ndarray:
    y = np.array(x, dtype=np.float32, copy=True).ravel()
    load().sko_gain_apply(_p(y), y.size, C.c_float(g))
    return y


def f32_to_s16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = np.empty(x.size, dtype=np.int16)
    load().sko_f32_to_s16_buf(_p(x), _p(out), x.size)
    return out


def s16_to_f32(s: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.int16).ravel()
    out = np.empty(s.size, dtype=np.float32)
    load().sko_s16_to_f32_buf(_p(s), _p(out), s.size)
    return out


def gain_f32_to_s16(x: np.ndarray, g: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = np.empty(x.size, dtype=np.int16)
    load().sko_gain_f32_to_s16_buf(_p(x), _p(out), x.size, C.c_float(g))
    return out


# ---------------------------------------------------------------- mixer

def _frames(frames):
    """frames: list of (samples f32 1-D, channels, unique)"""
    keep = [np.ascontiguousarray(f[0], dtype=np.float32).ravel() for f in frames]
    arr = (Frame * max(len(frames), 1))()
    for i, (f, k) in enumerate(zip(frames, keep)):
        arr[i].samples = k.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].n_samples = k.size
        arr[i].channels = f[1]
        arr[i].unique = 1 if (len(f) < 3 or f[2]) else 0
    return arr, keep


def mix_plan(frames, out_channels: int, out_size: int):
    arr, _keep = _frames(frames)
    order = np.zeros(max(len(frames), 1), dtype=np.uint32)
    hb = C.c_int()
    load().sko_mix_plan(arr, len(frames), out_channels, out_size, _p(order), C.byref(hb))
    return order[: len(frames)].tolist(), bool(hb.value)


def mix_sync(frames, max_channels_seen: int = 0):
    arr, _keep = _frames(frames)
    cap = max([k.size for k in _keep] + [1]) * 8 * max(max_channels_seen, 1)
    out = np.zeros(cap, dtype=np.float32)
    oc, ol = C.c_uint16(), C.c_size_t()
    rc = load().sko_mix_sync(arr, len(frames), max_channels_seen, _p(out), cap, C.byref(oc), C.byref(ol))
    assert rc == 0
    return out[: ol.value].copy(), oc.value


def mix_clocked(frames, out_channels: int, frame_samples_per_channel: int):
    arr, _keep = _frames(frames)
    out = np.zeros(out_channels * frame_samples_per_channel, dtype=np.float32)
    rc = load().sko_mix_clocked(arr, len(frames), out_channels, frame_samples_per_channel, _p(out), out.size)
    assert rc == 0
    return out


# ---------------------------------------------------------------- resampler

class FastFixedIn:
    """rubato::FastFixedIn<f32>, PolynomialDegree::Linear (restated)."""

    def __init__(self, in_rate: int, out_rate: int, chunk_frames: int, channels: int):
        self.lib = load()
        self.channels = channels
        self.chunk = chunk_frames
        self.h = self.lib.sko_ffi_new(float(out_rate) / float(in_rate), chunk_frames, channels)
        assert self.h
        self.out_max = self.lib.sko_ffi_out_max(self.h)

    def process(self, chunk: np.ndarray) -> np.ndarray:
        chunk = np.ascontiguousarray(chunk, dtype=np.float32).ravel()
        assert chunk.size == self.chunk * self.channels
        out = np.empty(self.out_max * self.channels, dtype=np.float32)
        n = self.lib.sko_ffi_process_interleaved(self.h, _p(chunk), _p(out), self.out_max)
        return out[: n * self.channels].copy()

    @property
    def last_index(self) -> float:
        return self.lib.sko_ffi_last_index(self.h)

    def history(self) -> np.ndarray:
        h = np.empty(16 * self.channels, dtype=np.float32)
        self.lib.sko_ffi_history(self.h, _p(h))
        return h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sko_ffi_free(self.h)
            self.h = None


class ResamplerNode:
    """AudioResamplerNode (resampler.rs:197-740) restated: push packets, collect emitted packets."""

    def __init__(self, target_sample_rate: int, chunk_frames: int = 960, output_frame_size: int = 960):
        self.lib = load()
        err = C.create_string_buffer(256)
        self.h = self.lib.sko_rsnode_new(target_sample_rate, chunk_frames, output_frame_size, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self.out = []
        self._cb = EMIT_FN(self._emit)

    def _emit(self, ud, rate, ch, samples, n, meta):
        arr = np.ctypeslib.as_array(samples, shape=(n,)).copy() if n else np.zeros(0, dtype=np.float32)
        m = meta.contents
        self.out.append(dict(sample_rate=rate, channels=ch, samples=arr,
                             timestamp_us=(m.timestamp_us if m.has_timestamp else None), duration_us=m.duration_us,
                             sequence=m.sequence))

    def push(self, sample_rate: int, channels: int, samples: np.ndarray, timestamp_us=None):
        s = np.ascontiguousarray(samples, dtype=np.float32).ravel()
        err = C.create_string_buffer(256)
        rc = self.lib.sko_rsnode_push(self.h, sample_rate, channels, _p(s), s.size, 0 if timestamp_us is None else 1,
                                      0 if timestamp_us is None else timestamp_us, self._cb, None, err, 256)
        if rc != 0:
            raise RuntimeError(err.value.decode())

    def finish(self):
        self.lib.sko_rsnode_finish(self.h, self._cb, None)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sko_rsnode_free(self.h)
            self.h = None


def duration_us_for_frames(rate: int, frames: int) -> int:
    return load().sko_duration_us_for_frames(rate, frames)


def chain_bench(n_sessions: int, k_inputs: int, ticks: int, in_rate: int, channels: int, input_pool: np.ndarray,
                in_gains: np.ndarray, master_gains: np.ndarray, threads: int, want_last: bool = False):
    """multi-threaded, reference-shaped CPU run of the full chain (sk_chain.c). Returns (seconds, checksum, last_out)."""
    pool = np.ascontiguousarray(input_pool, dtype=np.float32)
    ig = np.ascontiguousarray(in_gains, dtype=np.float32)
    mg = np.ascontiguousarray(master_gains, dtype=np.float32)
    assert ig.size >= n_sessions * k_inputs and mg.size >= n_sessions
    last = np.zeros((n_sessions, 960 * channels), dtype=np.int16) if want_last else None
    cs = C.c_uint64()
    sec = load().sko_chain_bench(n_sessions, k_inputs, ticks, in_rate, channels, _p(pool), pool.shape[0], _p(ig), _p(mg), threads,
                                 _p(last) if want_last else None, C.byref(cs))
    return sec, cs.value, last


def node_bench(kind: int, n_units: int, iters: int, pool: np.ndarray, gains: np.ndarray | None, threads: int, in_rate: int = 48000,
               out_rate: int = 48000, k_inputs: int = 64, want_last: bool = False, cap: int = 0):
    """multi-threaded, reference-shaped CPU run of a standalone node workload (sk_chain.c sko_node_bench):
    kind 0 gain->s16, 1 mixer(64)->gain->s16, 2 resampler. Returns (seconds, last_out, last_counts)."""
    pool = np.ascontiguousarray(pool, dtype=np.float32)
    g = np.ascontiguousarray(gains if gains is not None else np.ones(n_units, np.float32), dtype=np.float32)
    assert g.size >= n_units
    out_s16 = np.zeros((n_units, 1920), np.int16) if (want_last and kind in (0, 1)) else None
    out_f32 = np.zeros((n_units, cap * 2), np.float32) if (want_last and kind == 2) else None
    out_n = np.zeros(n_units, np.uint32) if (want_last and kind == 2) else None
    sec = load().sko_node_bench(kind, n_units, iters, in_rate, out_rate, k_inputs, _p(pool), pool.shape[0], _p(g), threads,
                                _p(out_s16) if out_s16 is not None else None, _p(out_f32) if out_f32 is not None else None,
                                _p(out_n) if out_n is not None else None, cap)
    return sec, (out_s16 if kind in (0, 1) else out_f32), out_n


class SincFixedIn:
    """windowed-sinc polyphase resampler of the build's own spec (oracle/sk_sinc.c header; no reference implementation exists)"""

    def __init__(self, in_rate: int, out_rate: int, chunk_frames: int, channels: int, sinc_len: int = 64, oversampling: int = 256,
                 f_cutoff: float = 0.95):
        self.lib = load()
        self.channels, self.chunk = channels, chunk_frames
        self.h = self.lib.sko_sinc_new(float(out_rate) / float(in_rate), chunk_frames, channels, sinc_len, oversampling, f_cutoff)
        assert self.h
        self.out_max = self.lib.sko_sinc_out_max(self.h)

    def process(self, chunk: np.ndarray) -> np.ndarray:
        chunk = np.ascontiguousarray(chunk, dtype=np.float32).ravel()
        assert chunk.size == self.chunk * self.channels
        out = np.empty(self.out_max * self.channels, dtype=np.float32)
        n = self.lib.sko_sinc_process_interleaved(self.h, _p(chunk), _p(out), self.out_max)
        return out[: n * self.channels].copy()

    @property
    def last_index(self) -> float:
        return self.lib.sko_sinc_last_index(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sko_sinc_free(self.h)
            self.h = None


def sinc_taps(sinc_len: int, oversampling: int, fc: float) -> np.ndarray:
    out = np.empty((oversampling + 1, sinc_len), dtype=np.float32)
    load().sko_sinc_taps(sinc_len, oversampling, fc, _p(out))
    return out
