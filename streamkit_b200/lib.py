"""ctypes binding of the batch C ABI (include/skgpu_batch.h) -- test / benchmark harness side.

The product is the shared library (streamkit_b200/csrc/libskgpu.so); this module only declares its
signatures so pytest and bench.py can call THROUGH the C ABI exactly like a Rust/C host would.
There is no fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libskgpu.so")

SKGPU_NO_GAIN = 0xFFFFFFFF
CVT_F32_TO_F32, CVT_F32_TO_S16, CVT_S16_TO_F32 = 0, 1, 2
RS_TO_FIFO = 1
MIX_IN_UNIQUE, MIX_IN_FIFO = 1, 2
MIX_OUT_S16 = 1
SUBMIT_NO_H2D, SUBMIT_NO_D2H, SUBMIT_GRAPH, SUBMIT_TIME_OPS, SUBMIT_OVERLAP_D2H, SUBMIT_SLICED = 1, 2, 4, 8, 16, 32
PIN_NUMA_LOCAL, PIN_WRITE_COMBINED = 1, 2
STREAM_S16, STREAM_SINC = 1, 2


class CtxConfig(C.Structure):
    _fields_ = [("max_streams", C.c_uint32), ("max_channels", C.c_uint32), ("fifo_frames", C.c_uint32), ("flags", C.c_uint32)]


class StreamCfg(C.Structure):
    _fields_ = [("in_rate", C.c_uint32), ("out_rate", C.c_uint32), ("chunk_frames", C.c_uint32), ("channels", C.c_uint16),
                ("flags", C.c_uint16)]


class SliceTiming(C.Structure):
    _fields_ = [("upload_done_ms", C.c_float), ("kernels_ms", C.c_float), ("latency_ms", C.c_float)]


class TickTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernels_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float)]


# numpy dtypes with the exact C layout of the descriptor structs
SEG_DT = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("n_samples", "<u4"), ("gain_idx", "<u4")], align=True)
RS_ITEM_DT = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("slot", "<u4"), ("out_cap_frames", "<u4"), ("flags", "<u4"),
                       ("reserved", "<u4")], align=True)
RS_RESULT_DT = np.dtype([("out_frames", "<u4"), ("status", "<u4")], align=True)
MIX_INPUT_DT = np.dtype([("in_off", "<u8"), ("n_frames", "<u4"), ("channels", "<u2"), ("flags", "<u2"), ("gain_idx", "<u4"),
                         ("slot", "<u4")], align=True)
MIX_GROUP_DT = np.dtype([("out_off", "<u8"), ("first_input", "<u4"), ("n_inputs", "<u4"), ("out_frames", "<u4"),
                         ("out_channels", "<u2"), ("flags", "<u2"), ("gain_idx", "<u4"), ("reserved", "<u4")], align=True)
CHAIN_INPUT_DT = np.dtype([("in_off", "<u8"), ("slot", "<u4"), ("gain_idx", "<u4"), ("flags", "<u4"), ("reserved", "<u4")], align=True)
CHAIN_GROUP_DT = np.dtype([("out_off", "<u8"), ("first_input", "<u4"), ("n_inputs", "<u4"), ("gain_idx", "<u4"),
                           ("out_channels", "<u2"), ("flags", "<u2")], align=True)
CHAIN_RESULT_DT = np.dtype([("emitted", "<u4"), ("status", "<u4")], align=True)
SLICE_DT = np.dtype([("group_end", "<u4"), ("input_end", "<u4"), ("h2d_end", "<u8"), ("d2h_off", "<u8", (2,)), ("d2h_bytes", "<u8", (2,))], align=True)
assert SLICE_DT.itemsize == 48
assert CHAIN_INPUT_DT.itemsize == 24 and CHAIN_GROUP_DT.itemsize == 24
assert SEG_DT.itemsize == 24 and RS_ITEM_DT.itemsize == 32 and MIX_INPUT_DT.itemsize == 24 and MIX_GROUP_DT.itemsize == 32

EXPORTS = [
    "skgpu_abi_version", "skgpu_last_error", "skgpu_ctx_create", "skgpu_ctx_destroy", "skgpu_ctx_device_info",
    "skgpu_pinned_alloc", "skgpu_pinned_free", "skgpu_stream_open", "skgpu_stream_open_many", "skgpu_stream_reset",
    "skgpu_stream_close", "skgpu_stream_get_state", "skgpu_stream_max_out_frames", "skgpu_plan_create",
    "skgpu_plan_destroy", "skgpu_plan_add_convert", "skgpu_plan_add_resample", "skgpu_plan_add_mix", "skgpu_plan_add_chain",
    "skgpu_plan_add_chain_cap",
    "skgpu_plan_update_convert", "skgpu_plan_update_resample", "skgpu_plan_update_mix", "skgpu_plan_update_chain",
    "skgpu_plan_set_io", "skgpu_plan_set_banks", "skgpu_plan_tick_count",
    "skgpu_plan_set_gains", "skgpu_plan_set_present", "skgpu_plan_finalize", "skgpu_tick_submit", "skgpu_tick_wait", "skgpu_tick_wait_for",
    "skgpu_plan_op_time", "skgpu_plan_reset_op_times", "skgpu_plan_launches_per_tick", "skgpu_arena_upload",
    "skgpu_arena_download", "skgpu_arena_fill", "skgpu_timer_start", "skgpu_timer_stop", "skgpu_timer_elapsed_ms",
    "skgpu_ctx_sync", "skgpu_ctx_flush_l2",
    "skgpu_pinned_alloc_ex", "skgpu_ctx_numa_node", "skgpu_ctx_bind_thread", "skgpu_plan_set_slices", "skgpu_plan_auto_slices",
    "skgpu_tick_slice_timing", "skgpu_ctx_set_sinc",
]

_lib = None


def load() -> C.CDLL:
    """dlopen libskgpu.so and declare signatures. Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(streamkit_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    sig = {
        "skgpu_abi_version": (u32, []),
        "skgpu_last_error": (C.c_char_p, []),
        "skgpu_ctx_create": (i32, [i32, C.POINTER(CtxConfig), C.POINTER(vp)]),
        "skgpu_ctx_destroy": (None, [vp]),
        "skgpu_ctx_device_info": (i32, [vp, C.c_char_p, C.c_size_t, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "skgpu_pinned_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "skgpu_pinned_free": (i32, [vp, vp]),
        "skgpu_stream_open": (i32, [vp, C.POINTER(StreamCfg), C.POINTER(u32)]),
        "skgpu_stream_open_many": (i32, [vp, C.POINTER(StreamCfg), u32, vp]),
        "skgpu_stream_reset": (i32, [vp, u32]),
        "skgpu_stream_close": (i32, [vp, u32]),
        "skgpu_stream_get_state": (i32, [vp, u32, C.POINTER(C.c_double), vp, C.POINTER(u64), C.POINTER(u64)]),
        "skgpu_stream_max_out_frames": (u32, [C.POINTER(StreamCfg)]),
        "skgpu_plan_create": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "skgpu_plan_destroy": (None, [vp]),
        "skgpu_plan_add_convert": (i32, [vp, C.c_int, vp, u32, C.POINTER(u32)]),
        "skgpu_plan_add_resample": (i32, [vp, vp, u32, u64, C.POINTER(u32)]),
        "skgpu_plan_add_mix": (i32, [vp, vp, u32, vp, u32, C.POINTER(u32)]),
        "skgpu_plan_add_chain": (i32, [vp, vp, u32, vp, u32, u32, u64, C.POINTER(u32)]),
        "skgpu_tick_wait_for": (i32, [vp, C.c_uint64]),
        "skgpu_plan_add_chain_cap": (i32, [vp, vp, u32, vp, u32, u32, u32, u32, u32, u64, C.POINTER(u32)]),
        "skgpu_plan_update_chain": (i32, [vp, u32, vp, u32, vp, u32]),
        "skgpu_plan_set_banks": (i32, [vp, u64]),
        "skgpu_plan_tick_count": (u64, [vp]),
        "skgpu_plan_update_convert": (i32, [vp, u32, vp, u32]),
        "skgpu_plan_update_resample": (i32, [vp, u32, vp, u32]),
        "skgpu_plan_update_mix": (i32, [vp, u32, vp, u32, vp, u32]),
        "skgpu_plan_set_io": (i32, [vp, u64, u64, u64, u64]),
        "skgpu_plan_set_gains": (i32, [vp, vp, u32]),
        "skgpu_plan_set_present": (i32, [vp, u32, vp, u32]),
        "skgpu_plan_finalize": (i32, [vp]),
        "skgpu_tick_submit": (i32, [vp, vp, vp, u32]),
        "skgpu_tick_wait": (i32, [vp, C.POINTER(TickTiming)]),
        "skgpu_plan_op_time": (i32, [vp, u32, u32, C.POINTER(C.c_float), C.POINTER(u32)]),
        "skgpu_plan_reset_op_times": (i32, [vp]),
        "skgpu_plan_launches_per_tick": (u32, [vp]),
        "skgpu_arena_upload": (i32, [vp, u64, vp, C.c_size_t]),
        "skgpu_arena_download": (i32, [vp, u64, vp, C.c_size_t]),
        "skgpu_arena_fill": (i32, [vp, u64, C.c_int, C.c_size_t]),
        "skgpu_timer_start": (i32, [vp]),
        "skgpu_timer_stop": (i32, [vp]),
        "skgpu_timer_elapsed_ms": (i32, [vp, C.POINTER(C.c_float)]),
        "skgpu_ctx_sync": (i32, [vp]),
        "skgpu_ctx_flush_l2": (i32, [vp]),
        "skgpu_pinned_alloc_ex": (i32, [vp, C.c_size_t, u32, C.POINTER(vp), C.POINTER(i32)]),
        "skgpu_ctx_numa_node": (i32, [vp]),
        "skgpu_ctx_bind_thread": (i32, [vp]),
        "skgpu_plan_set_slices": (i32, [vp, u32, vp, u32]),
        "skgpu_plan_auto_slices": (i32, [vp, u32, u32]),
        "skgpu_tick_slice_timing": (i32, [vp, u64, C.POINTER(SliceTiming), u32, C.POINTER(u32)]),
        "skgpu_ctx_set_sinc": (i32, [vp, u32, u32, C.c_double]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SkgpuError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"skgpu error {rc}: {msg}")
        self.rc = rc
        self.msg = msg


def _chk(rc: int) -> None:
    if rc != 0:
        raise SkgpuError(rc, load().skgpu_last_error().decode("utf-8", "replace"))


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Context:
    """One GPU: a CUDA stream + per-stream resampler state in HBM."""

    def __init__(self, device: int = 0, max_streams: int = 1024, max_channels: int = 2, fifo_frames: int = 0):
        self.lib = load()
        cfg = CtxConfig(max_streams, max_channels, fifo_frames, 0)
        h = C.c_void_p()
        _chk(self.lib.skgpu_ctx_create(device, C.byref(cfg), C.byref(h)))
        self.h = h
        self.max_channels = max_channels
        self._pinned = []

    def close(self):
        if self.h:
            self.lib.skgpu_ctx_destroy(self.h)
            self.h = None

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
        _chk(self.lib.skgpu_ctx_device_info(self.h, name, 128, C.byref(sm), C.byref(ma), C.byref(mi)))
        return name.value.decode(), sm.value, ma.value, mi.value

    def set_sinc(self, sinc_len: int = 64, oversampling: int = 256, f_cutoff: float = 0.95):
        """enable the windowed-sinc polyphase mode for STREAM_SINC streams (spec: include/skgpu_batch.h skgpu_ctx_set_sinc)"""
        _chk(self.lib.skgpu_ctx_set_sinc(self.h, sinc_len, oversampling, f_cutoff))

    def numa_node(self) -> int:
        return self.lib.skgpu_ctx_numa_node(self.h)

    def bind_thread(self):
        _chk(self.lib.skgpu_ctx_bind_thread(self.h))

    def pinned(self, nbytes: int, dtype=np.uint8, flags: int = PIN_NUMA_LOCAL) -> np.ndarray:
        p = C.c_void_p()
        node = C.c_int32()
        _chk(self.lib.skgpu_pinned_alloc_ex(self.h, nbytes, flags, C.byref(p), C.byref(node)))
        self.last_pinned_node = node.value
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
        self._pinned.append((p, arr))
        return arr.view(dtype)

    def stream_open(self, in_rate: int, out_rate: int, chunk_frames: int, channels: int, flags: int = 0) -> int:
        cfg = StreamCfg(in_rate, out_rate, chunk_frames, channels, flags)
        slot = C.c_uint32()
        _chk(self.lib.skgpu_stream_open(self.h, C.byref(cfg), C.byref(slot)))
        return slot.value

    def stream_open_many(self, in_rate: int, out_rate: int, chunk_frames: int, channels: int, n: int, flags: int = 0) -> np.ndarray:
        cfg = StreamCfg(in_rate, out_rate, chunk_frames, channels, flags)
        slots = np.zeros(n, dtype=np.uint32)
        _chk(self.lib.skgpu_stream_open_many(self.h, C.byref(cfg), n, _ptr(slots)))
        return slots

    def stream_reset(self, slot: int):
        _chk(self.lib.skgpu_stream_reset(self.h, slot))

    def stream_close(self, slot: int):
        _chk(self.lib.skgpu_stream_close(self.h, slot))

    def stream_state(self, slot: int, channels: int):
        li = C.c_double()
        hist = np.zeros(16 * channels, dtype=np.float32)
        w, r = C.c_uint64(), C.c_uint64()
        _chk(self.lib.skgpu_stream_get_state(self.h, slot, C.byref(li), _ptr(hist), C.byref(w), C.byref(r)))
        return li.value, hist, w.value, r.value

    @staticmethod
    def max_out_frames(in_rate, out_rate, chunk_frames, channels=2) -> int:
        cfg = StreamCfg(in_rate, out_rate, chunk_frames, channels, 0)
        return load().skgpu_stream_max_out_frames(C.byref(cfg))

    def timer_start(self):
        _chk(self.lib.skgpu_timer_start(self.h))

    def timer_stop(self):
        _chk(self.lib.skgpu_timer_stop(self.h))

    def timer_ms(self) -> float:
        ms = C.c_float()
        _chk(self.lib.skgpu_timer_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def sync(self):
        _chk(self.lib.skgpu_ctx_sync(self.h))

    def flush_l2(self):
        _chk(self.lib.skgpu_ctx_flush_l2(self.h))


class Plan:
    """One compiled tick: device arena + ordered ops with device-resident descriptor tables."""

    def __init__(self, ctx: Context, arena_bytes: int):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        _chk(self.lib.skgpu_plan_create(ctx.h, arena_bytes, C.byref(h)))
        self.h = h
        self.arena_bytes = arena_bytes

    def destroy(self):
        if self.h:
            self.lib.skgpu_plan_destroy(self.h)
            self.h = None

    def set_gains(self, gains):
        g = np.ascontiguousarray(gains, dtype=np.float32)
        _chk(self.lib.skgpu_plan_set_gains(self.h, _ptr(g), g.size))

    def add_convert(self, mode: int, segs: np.ndarray) -> int:
        segs = np.ascontiguousarray(segs, dtype=SEG_DT)
        op = C.c_uint32()
        _chk(self.lib.skgpu_plan_add_convert(self.h, mode, _ptr(segs), segs.size, C.byref(op)))
        return op.value

    def add_resample(self, items: np.ndarray, results_off: int) -> int:
        items = np.ascontiguousarray(items, dtype=RS_ITEM_DT)
        op = C.c_uint32()
        _chk(self.lib.skgpu_plan_add_resample(self.h, _ptr(items), items.size, results_off, C.byref(op)))
        return op.value

    def add_mix(self, groups: np.ndarray, inputs: np.ndarray) -> int:
        groups = np.ascontiguousarray(groups, dtype=MIX_GROUP_DT)
        inputs = np.ascontiguousarray(inputs, dtype=MIX_INPUT_DT)
        op = C.c_uint32()
        _chk(self.lib.skgpu_plan_add_mix(self.h, _ptr(groups), groups.size, _ptr(inputs), inputs.size, C.byref(op)))
        return op.value

    def add_chain(self, groups: np.ndarray, inputs: np.ndarray, output_frame_size: int, results_off: int) -> int:
        groups = np.ascontiguousarray(groups, dtype=CHAIN_GROUP_DT)
        inputs = np.ascontiguousarray(inputs, dtype=CHAIN_INPUT_DT)
        op = C.c_uint32()
        _chk(self.lib.skgpu_plan_add_chain(self.h, _ptr(groups), groups.size, _ptr(inputs), inputs.size, output_frame_size,
                                           results_off, C.byref(op)))
        return op.value

    def update_chain(self, op: int, groups: np.ndarray, inputs: np.ndarray):
        groups = np.ascontiguousarray(groups, dtype=CHAIN_GROUP_DT)
        inputs = np.ascontiguousarray(inputs, dtype=CHAIN_INPUT_DT)
        _chk(self.lib.skgpu_plan_update_chain(self.h, op, _ptr(groups), groups.size, _ptr(inputs), inputs.size))

    def set_slices(self, op: int, slices: np.ndarray):
        sl = np.ascontiguousarray(slices, dtype=SLICE_DT)
        _chk(self.lib.skgpu_plan_set_slices(self.h, op, _ptr(sl), sl.size))

    def auto_slices(self, op: int, n: int):
        _chk(self.lib.skgpu_plan_auto_slices(self.h, op, n))

    def slice_timing(self, tick: int):
        """[(upload_done_ms, kernels_ms, latency_ms)] of a finished sliced tick (one of the two most recent)"""
        buf = (SliceTiming * 4096)()
        n = C.c_uint32()
        _chk(self.lib.skgpu_tick_slice_timing(self.h, tick, buf, 4096, C.byref(n)))
        return [(buf[i].upload_done_ms, buf[i].kernels_ms, buf[i].latency_ms) for i in range(n.value)]

    def wait_for(self, tick: int):
        _chk(self.lib.skgpu_tick_wait_for(self.h, tick))

    def set_banks(self, bank_stride: int):
        _chk(self.lib.skgpu_plan_set_banks(self.h, bank_stride))

    def tick_count(self) -> int:
        return self.lib.skgpu_plan_tick_count(self.h)

    def update_convert(self, op: int, segs: np.ndarray):
        segs = np.ascontiguousarray(segs, dtype=SEG_DT)
        _chk(self.lib.skgpu_plan_update_convert(self.h, op, _ptr(segs), segs.size))

    def update_resample(self, op: int, items: np.ndarray):
        items = np.ascontiguousarray(items, dtype=RS_ITEM_DT)
        _chk(self.lib.skgpu_plan_update_resample(self.h, op, _ptr(items), items.size))

    def update_mix(self, op: int, groups: np.ndarray, inputs: np.ndarray):
        groups = np.ascontiguousarray(groups, dtype=MIX_GROUP_DT)
        inputs = np.ascontiguousarray(inputs, dtype=MIX_INPUT_DT)
        _chk(self.lib.skgpu_plan_update_mix(self.h, op, _ptr(groups), groups.size, _ptr(inputs), inputs.size))

    def set_present(self, mix_op: int, present):
        p = np.ascontiguousarray(present, dtype=np.uint8)
        _chk(self.lib.skgpu_plan_set_present(self.h, mix_op, _ptr(p), p.size))

    def set_io(self, h2d_off: int, h2d_bytes: int, d2h_off: int, d2h_bytes: int):
        _chk(self.lib.skgpu_plan_set_io(self.h, h2d_off, h2d_bytes, d2h_off, d2h_bytes))

    def finalize(self):
        _chk(self.lib.skgpu_plan_finalize(self.h))

    def submit(self, host_in: np.ndarray | None = None, host_out: np.ndarray | None = None, flags: int = 0):
        pi = _ptr(host_in) if host_in is not None else None
        po = _ptr(host_out) if host_out is not None else None
        _chk(self.lib.skgpu_tick_submit(self.h, pi, po, flags))

    def wait(self) -> TickTiming:
        t = TickTiming()
        _chk(self.lib.skgpu_tick_wait(self.h, C.byref(t)))
        return t

    def op_time(self, op: int, sub: int = 0):
        ms, n = C.c_float(), C.c_uint32()
        _chk(self.lib.skgpu_plan_op_time(self.h, op, sub, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def reset_op_times(self):
        _chk(self.lib.skgpu_plan_reset_op_times(self.h))

    def launches_per_tick(self) -> int:
        return self.lib.skgpu_plan_launches_per_tick(self.h)

    def upload(self, off: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr)
        _chk(self.lib.skgpu_arena_upload(self.h, off, _ptr(a), a.nbytes))

    def download(self, off: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        _chk(self.lib.skgpu_arena_download(self.h, off, _ptr(out), nbytes))
        return out.view(dtype)

    def fill(self, off: int, value: int, nbytes: int):
        _chk(self.lib.skgpu_arena_fill(self.h, off, value, nbytes))
