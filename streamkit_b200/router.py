"""ctypes binding of the single-process multi-GPU session router (include/skgpu_router.h, csrc/host/router.cpp ->
libskgpu_router.so). Test / bench harness only; a StreamKit engine binds the same C ABI from Rust."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import hub as H

ROUTER_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libskgpu_router.so")
EXPORTS = [
    "skgpu_router_last_error", "skgpu_fnv1a64", "skgpu_router_gpu_for", "skgpu_router_create", "skgpu_router_destroy", "skgpu_router_gpus",
    "skgpu_router_hub", "skgpu_router_numa_node", "skgpu_router_session_open", "skgpu_router_session_close", "skgpu_router_push",
    "skgpu_router_set_input_gain", "skgpu_router_set_master_gain", "skgpu_router_tick", "skgpu_router_wait", "skgpu_router_run_ticks",
    "skgpu_router_session_output",
]
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    H.load()
    if not os.path.exists(ROUTER_LIB_PATH):
        raise RuntimeError("libskgpu_router.so is not built (python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(ROUTER_LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    lib.skgpu_router_last_error.restype = C.c_char_p
    lib.skgpu_fnv1a64.restype = u64
    lib.skgpu_fnv1a64.argtypes = [C.c_char_p, C.c_size_t]
    lib.skgpu_router_gpu_for.restype = u32
    lib.skgpu_router_gpu_for.argtypes = [C.c_char_p, C.c_size_t, u32]
    lib.skgpu_router_create.argtypes = [C.POINTER(i32), u32, C.POINTER(H.HubConfig), C.POINTER(vp)]
    lib.skgpu_router_destroy.argtypes = [vp]
    lib.skgpu_router_destroy.restype = None
    lib.skgpu_router_gpus.argtypes = [vp]
    lib.skgpu_router_gpus.restype = u32
    lib.skgpu_router_hub.argtypes = [vp, u32]
    lib.skgpu_router_hub.restype = vp
    lib.skgpu_router_numa_node.argtypes = [vp, u32]
    lib.skgpu_router_numa_node.restype = i32
    lib.skgpu_router_session_open.argtypes = [vp, C.c_char_p, C.c_size_t, u32, C.POINTER(u32), C.POINTER(u64)]
    lib.skgpu_router_session_close.argtypes = [vp, u64]
    lib.skgpu_router_push.argtypes = [vp, u64, u32, vp, u32]
    lib.skgpu_router_set_input_gain.argtypes = [vp, u64, u32, C.c_float]
    lib.skgpu_router_set_master_gain.argtypes = [vp, u64, C.c_float]
    lib.skgpu_router_tick.argtypes = [vp]
    lib.skgpu_router_wait.argtypes = [vp]
    lib.skgpu_router_run_ticks.argtypes = [vp, u32, C.POINTER(C.c_double)]
    lib.skgpu_router_session_output.argtypes = [vp, u64, C.POINTER(vp), C.POINTER(u32), C.POINTER(u32)]
    _lib = lib
    return lib


class RouterError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"skgpu_router error {rc}: {msg}")
        self.rc, self.msg = rc, msg


def _chk(rc: int) -> None:
    if rc != 0:
        raise RouterError(rc, (load().skgpu_router_last_error() or b"").decode(errors="replace"))


def fnv1a64(data: bytes) -> int:
    return load().skgpu_fnv1a64(data, len(data))


def gpu_for(session_id: str, n_gpus: int) -> int:
    b = session_id.encode("utf-8")
    return load().skgpu_router_gpu_for(b, len(b), n_gpus)


class Router:
    """n GPUs, one hub + one NUMA-pinned tick thread each; sessions routed by fnv1a64(session id) % n"""

    def __init__(self, devices, max_sessions: int, max_streams: int, in_rates, max_inputs_per_session: int = 8, out_rate: int = 48000,
                 out_frames: int = 960, channels: int = 2, s16: bool = True, in_s16: bool = False, jitter_frames: int = 1, slices: int = 0):
        self.lib = load()
        rates = (C.c_uint32 * len(in_rates))(*in_rates)
        self._rates = rates
        cfg = H.HubConfig(max_sessions, max_streams, max_inputs_per_session, out_rate, out_frames, channels,
                          (H.OUT_S16 if s16 else 0) | (H.IN_S16 if in_s16 else 0), C.cast(rates, C.POINTER(C.c_uint32)), len(in_rates), jitter_frames, slices)
        devs = (C.c_int32 * len(devices))(*devices)
        self.h = C.c_void_p()
        _chk(self.lib.skgpu_router_create(devs, len(devices), C.byref(cfg), C.byref(self.h)))
        self.n, self.F, self.C, self.s16 = len(devices), out_frames, channels, s16
        self.in_dtype = np.int16 if in_s16 else np.float32

    def close(self):
        if self.h:
            self.lib.skgpu_router_destroy(self.h)
            self.h = C.c_void_p()

    def numa_nodes(self):
        return [self.lib.skgpu_router_numa_node(self.h, g) for g in range(self.n)]

    def hub_handle(self, g: int):
        return self.lib.skgpu_router_hub(self.h, g)

    def session_open(self, session_id: str, in_rates) -> int:
        b = session_id.encode("utf-8")
        arr = (C.c_uint32 * len(in_rates))(*in_rates)
        h = C.c_uint64()
        _chk(self.lib.skgpu_router_session_open(self.h, b, len(b), len(in_rates), arr, C.byref(h)))
        return h.value

    def session_close(self, handle: int):
        _chk(self.lib.skgpu_router_session_close(self.h, handle))

    def push(self, handle: int, inp: int, samples: np.ndarray):
        x = np.ascontiguousarray(samples, dtype=self.in_dtype).reshape(-1)
        _chk(self.lib.skgpu_router_push(self.h, handle, inp, x.ctypes.data_as(C.c_void_p), x.size // self.C))

    def set_input_gain(self, handle: int, inp: int, gain: float):
        _chk(self.lib.skgpu_router_set_input_gain(self.h, handle, inp, gain))

    def set_master_gain(self, handle: int, gain: float):
        _chk(self.lib.skgpu_router_set_master_gain(self.h, handle, gain))

    def tick(self):
        _chk(self.lib.skgpu_router_tick(self.h))

    def wait(self):
        _chk(self.lib.skgpu_router_wait(self.h))

    def run_ticks(self, n: int) -> float:
        ms = C.c_double()
        _chk(self.lib.skgpu_router_run_ticks(self.h, n, C.byref(ms)))
        return ms.value

    def output(self, handle: int):
        p, n, st = C.c_void_p(), C.c_uint32(), C.c_uint32()
        _chk(self.lib.skgpu_router_session_output(self.h, handle, C.byref(p), C.byref(n), C.byref(st)))
        if not p.value:
            return None, n.value, st.value
        cnt = self.F * self.C
        ct = C.c_int16 if self.s16 else C.c_float
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(cnt,)).copy(), n.value, st.value
