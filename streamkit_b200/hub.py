"""ctypes binding of the frame-batching layer (include/skgpu_hub.h, csrc/host/hub.cpp -> libskgpu_hub.so).
Python is only the test / bench harness here; a StreamKit engine binds the same C ABI from Rust."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as L

HUB_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libskgpu_hub.so")
OUT_S16, IN_S16 = 1, 2

EXPORTS = [
    "skgpu_hub_last_error", "skgpu_hub_create", "skgpu_hub_destroy", "skgpu_hub_session_open", "skgpu_hub_session_close",
    "skgpu_hub_set_input_gain", "skgpu_hub_set_master_gain", "skgpu_hub_chunk_frames", "skgpu_hub_push", "skgpu_hub_push_batch",
    "skgpu_hub_acquire", "skgpu_hub_commit", "skgpu_hub_commit_all",
    "skgpu_hub_tick",
    "skgpu_hub_wait", "skgpu_hub_wait_tick", "skgpu_hub_session_output", "skgpu_hub_live_sessions", "skgpu_hub_live_streams", "skgpu_hub_ticks",
    "skgpu_hub_get_stats", "skgpu_hub_state", "skgpu_hub_bind_thread", "skgpu_hub_session_open_ex", "skgpu_hub_input_eof", "skgpu_hub_session_state",
]


SESSION_SYNC = 1
SESSION_RUNNING, SESSION_DEGRADED, SESSION_STOPPED = 1, 2, 4


class SessionState(C.Structure):
    _fields_ = [("state", C.c_uint32), ("mixed", C.c_uint32), ("slow_mask", C.c_uint64), ("eof_mask", C.c_uint64), ("newly_slow", C.c_uint64),
                ("recovered", C.c_uint64)]


class HubStats(C.Structure):
    _fields_ = [("received", C.c_uint64), ("sent", C.c_uint64), ("discarded", C.c_uint64), ("errored", C.c_uint64)]


class HubConfig(C.Structure):
    _fields_ = [("max_sessions", C.c_uint32), ("max_streams", C.c_uint32), ("max_inputs_per_session", C.c_uint32),
                ("out_rate", C.c_uint32), ("out_frames", C.c_uint32), ("channels", C.c_uint16), ("flags", C.c_uint16),
                ("in_rates", C.POINTER(C.c_uint32)), ("n_in_rates", C.c_uint32), ("jitter_frames", C.c_uint32), ("slices", C.c_uint32)]


class HubFrame(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("session", C.c_uint32), ("input", C.c_uint32), ("n_frames", C.c_uint32), ("reserved", C.c_uint32)]


FRAME_DT = np.dtype([("samples", "<u8"), ("session", "<u4"), ("input", "<u4"), ("n_frames", "<u4"), ("reserved", "<u4")], align=True)
assert FRAME_DT.itemsize == C.sizeof(HubFrame) == 24

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L.load()   # libskgpu.so first (the hub links against it; rpath $ORIGIN finds it too)
    if not os.path.exists(HUB_LIB_PATH):
        raise RuntimeError("libskgpu_hub.so is not built (python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(HUB_LIB_PATH)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
    lib.skgpu_hub_last_error.restype = C.c_char_p
    lib.skgpu_hub_create.argtypes = [i32, C.POINTER(HubConfig), C.POINTER(vp)]
    lib.skgpu_hub_destroy.argtypes = [vp]
    lib.skgpu_hub_destroy.restype = None
    lib.skgpu_hub_session_open.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.skgpu_hub_session_close.argtypes = [vp, u32]
    lib.skgpu_hub_set_input_gain.argtypes = [vp, u32, u32, C.c_float]
    lib.skgpu_hub_set_master_gain.argtypes = [vp, u32, C.c_float]
    lib.skgpu_hub_chunk_frames.argtypes = [vp, u32, u32, C.POINTER(u32)]
    lib.skgpu_hub_push.argtypes = [vp, u32, u32, vp, u32]
    lib.skgpu_hub_push_batch.argtypes = [vp, vp, u32, u32]
    lib.skgpu_hub_acquire.argtypes = [vp, u32, u32, C.POINTER(vp), C.POINTER(u32)]
    lib.skgpu_hub_commit.argtypes = [vp, u32, u32]
    lib.skgpu_hub_commit_all.argtypes = [vp]
    lib.skgpu_hub_tick.argtypes = [vp]
    lib.skgpu_hub_wait.argtypes = [vp, C.POINTER(L.TickTiming)]
    lib.skgpu_hub_wait_tick.argtypes = [vp, C.c_uint64]
    lib.skgpu_hub_session_output.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(u32)]
    for n in ("skgpu_hub_live_sessions", "skgpu_hub_live_streams"):
        getattr(lib, n).argtypes = [vp]
        getattr(lib, n).restype = u32
    lib.skgpu_hub_ticks.argtypes = [vp]
    lib.skgpu_hub_ticks.restype = C.c_uint64
    lib.skgpu_hub_get_stats.argtypes = [vp, C.POINTER(HubStats)]
    lib.skgpu_hub_bind_thread.argtypes = [vp]
    lib.skgpu_hub_session_open_ex.argtypes = [vp, u32, C.POINTER(u32), u32, u32, C.POINTER(u32)]
    lib.skgpu_hub_input_eof.argtypes = [vp, u32, u32]
    lib.skgpu_hub_session_state.argtypes = [vp, u32, C.POINTER(SessionState)]
    lib.skgpu_hub_bind_thread.restype = i32
    lib.skgpu_hub_state.argtypes = [vp, C.POINTER(C.c_char_p)]
    lib.skgpu_hub_state.restype = u32
    _lib = lib
    return lib


class HubError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"skgpu_hub error {rc}: {msg}")
        self.rc, self.msg = rc, msg


def _chk(rc: int) -> None:
    if rc != 0:
        raise HubError(rc, (load().skgpu_hub_last_error() or b"").decode(errors="replace"))


class Hub:
    """One GPU's frame-batching layer: sessions of n resampled inputs -> gain -> clocked mix -> gain -> s16."""

    def __init__(self, max_sessions: int, max_streams: int, in_rates, max_inputs_per_session: int = 8, out_rate: int = 48000,
                 out_frames: int = 960, channels: int = 2, s16: bool = True, device: int = 0, jitter_frames: int = 1, in_s16: bool = False,
                 slices: int = 0):
        self.lib = load()
        rates = (C.c_uint32 * len(in_rates))(*in_rates)
        self._rates = rates
        cfg = HubConfig(max_sessions, max_streams, max_inputs_per_session, out_rate, out_frames, channels,
                        (OUT_S16 if s16 else 0) | (IN_S16 if in_s16 else 0), C.cast(rates, C.POINTER(C.c_uint32)), len(in_rates), jitter_frames, slices)
        self.in_dtype = np.int16 if in_s16 else np.float32
        self.h = C.c_void_p()
        _chk(self.lib.skgpu_hub_create(device, C.byref(cfg), C.byref(self.h)))
        self.F, self.C, self.s16 = out_frames, channels, s16

    def close(self):
        if self.h:
            self.lib.skgpu_hub_destroy(self.h)
            self.h = C.c_void_p()

    def session_open(self, in_rates) -> int:
        arr = (C.c_uint32 * len(in_rates))(*in_rates)
        sid = C.c_uint32()
        _chk(self.lib.skgpu_hub_session_open(self.h, len(in_rates), arr, C.byref(sid)))
        return sid.value

    def session_open_sync(self, in_rates, sync_timeout_ms: int | None = 100) -> int:
        """audio::mixer sync mode (mixer.rs:554-918); sync_timeout_ms None = wait forever"""
        arr = (C.c_uint32 * len(in_rates))(*in_rates)
        sid = C.c_uint32()
        _chk(self.lib.skgpu_hub_session_open_ex(self.h, len(in_rates), arr, SESSION_SYNC, sync_timeout_ms or 0, C.byref(sid)))
        return sid.value

    def input_eof(self, session: int, inp: int):
        _chk(self.lib.skgpu_hub_input_eof(self.h, session, inp))

    def session_state(self, session: int) -> SessionState:
        st = SessionState()
        _chk(self.lib.skgpu_hub_session_state(self.h, session, C.byref(st)))
        return st

    def session_close(self, session: int):
        _chk(self.lib.skgpu_hub_session_close(self.h, session))

    def set_input_gain(self, session: int, inp: int, gain: float):
        _chk(self.lib.skgpu_hub_set_input_gain(self.h, session, inp, gain))

    def set_master_gain(self, session: int, gain: float):
        _chk(self.lib.skgpu_hub_set_master_gain(self.h, session, gain))

    def chunk_frames(self, session: int, inp: int) -> int:
        n = C.c_uint32()
        _chk(self.lib.skgpu_hub_chunk_frames(self.h, session, inp, C.byref(n)))
        return n.value

    def push(self, session: int, inp: int, samples: np.ndarray):
        x = np.ascontiguousarray(samples, dtype=self.in_dtype).reshape(-1)
        _chk(self.lib.skgpu_hub_push(self.h, session, inp, x.ctypes.data_as(C.c_void_p), x.size // self.C))

    def acquire(self, session: int, inp: int) -> np.ndarray:
        """zero-copy: a writable float32 (int16 for s16 hubs) view of the stream's pinned slot for the next tick; fill it, then commit()"""
        p, n = C.c_void_p(), C.c_uint32()
        _chk(self.lib.skgpu_hub_acquire(self.h, session, inp, C.byref(p), C.byref(n)))
        ct = C.c_int16 if self.in_dtype == np.int16 else C.c_float
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n.value * self.C,))

    def commit(self, session: int, inp: int):
        _chk(self.lib.skgpu_hub_commit(self.h, session, inp))

    def commit_all(self):
        _chk(self.lib.skgpu_hub_commit_all(self.h))

    def push_batch(self, frames: np.ndarray, n_threads: int = 1):
        """frames: FRAME_DT array (samples = host addresses of interleaved f32 chunks that stay alive during the call)"""
        assert frames.dtype == FRAME_DT
        _chk(self.lib.skgpu_hub_push_batch(self.h, frames.ctypes.data_as(C.c_void_p), frames.size, n_threads))

    def tick(self):
        _chk(self.lib.skgpu_hub_tick(self.h))

    def wait(self) -> L.TickTiming:
        t = L.TickTiming()
        _chk(self.lib.skgpu_hub_wait(self.h, C.byref(t)))
        return t

    def wait_tick(self, tick: int):
        _chk(self.lib.skgpu_hub_wait_tick(self.h, tick))

    @property
    def ticks(self) -> int:
        return self.lib.skgpu_hub_ticks(self.h)

    def output(self, session: int):
        """(samples copy or None, n_mixed, status) of the last waited tick"""
        p, n, st = C.c_void_p(), C.c_uint32(), C.c_uint32()
        _chk(self.lib.skgpu_hub_session_output(self.h, session, C.byref(p), C.byref(n), C.byref(st)))
        if not p.value:
            return None, n.value, st.value
        cnt = self.F * self.C
        if self.s16:
            buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16)), shape=(cnt,)).copy()
        else:
            buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(cnt,)).copy()
        return buf, n.value, st.value

    def stats(self) -> dict:
        st = HubStats()
        _chk(self.lib.skgpu_hub_get_stats(self.h, C.byref(st)))
        return {"received": st.received, "sent": st.sent, "discarded": st.discarded, "errored": st.errored}

    def state(self):
        reason = C.c_char_p()
        code = self.lib.skgpu_hub_state(self.h, C.byref(reason))
        return code, (reason.value.decode() if reason.value else None)

    @property
    def live_sessions(self) -> int:
        return self.lib.skgpu_hub_live_sessions(self.h)

    @property
    def live_streams(self) -> int:
        return self.lib.skgpu_hub_live_streams(self.h)
