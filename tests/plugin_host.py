"""Minimal host for StreamKit's native plugin C ABI v2, in ctypes (test infrastructure).

Mirrors the call discipline of crates/plugin-native/src/wrapper.rs:398-457: one packet per
process_packet call on pin "in", outputs COPIED inside the output callback
(sdks/plugin-sdk/native/src/conversions.rs:340-346), error strings borrowed (types.rs:42-48).
Struct layouts follow sdks/plugin-sdk/native/src/types.rs:137-261.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

API_VERSION = 2
PACKET_RAW_AUDIO, PACKET_BINARY = 0, 5


class CResult(C.Structure):
    _fields_ = [("success", C.c_bool), ("error_message", C.c_char_p)]


class CAudioFrame(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint16), ("samples", C.POINTER(C.c_float)), ("sample_count", C.c_size_t)]


class CPacket(C.Structure):
    _fields_ = [("packet_type", C.c_int), ("data", C.c_void_p), ("len", C.c_size_t)]


class CAudioFormat(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint16), ("sample_format", C.c_int)]


class CPacketTypeInfo(C.Structure):
    _fields_ = [("type_discriminant", C.c_int), ("audio_format", C.POINTER(CAudioFormat)), ("custom_type_id", C.c_char_p)]


class CInputPin(C.Structure):
    _fields_ = [("name", C.c_char_p), ("accepts_types", C.POINTER(CPacketTypeInfo)), ("accepts_types_count", C.c_size_t)]


class COutputPin(C.Structure):
    _fields_ = [("name", C.c_char_p), ("produces_type", CPacketTypeInfo)]


class CNodeMetadata(C.Structure):
    _fields_ = [("kind", C.c_char_p), ("description", C.c_char_p), ("inputs", C.POINTER(CInputPin)), ("inputs_count", C.c_size_t),
                ("outputs", C.POINTER(COutputPin)), ("outputs_count", C.c_size_t), ("param_schema", C.c_char_p),
                ("categories", C.POINTER(C.c_char_p)), ("categories_count", C.c_size_t)]


LOG_CB = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p)
# struct-returning callbacks cannot be written in ctypes: they live in tests/host/host_shim.c
OUT_CB = C.c_void_p
TEL_CB = C.c_void_p


class SkhOut(C.Structure):
    _fields_ = [("packet_type", C.c_int), ("sample_rate", C.c_uint32), ("channels", C.c_uint16), ("n", C.c_size_t),
                ("data", C.c_void_p), ("pin", C.c_char * 32)]


_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM_SRC = os.path.join(_HERE, "host", "host_shim.c")
_SHIM_SO = os.path.join(_HERE, "host", "libskh.so")
_shim = None


def shim() -> C.CDLL:
    global _shim
    if _shim is None:
        if not os.path.exists(_SHIM_SO) or os.path.getmtime(_SHIM_SO) < os.path.getmtime(_SHIM_SRC):
            subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-o", _SHIM_SO, _SHIM_SRC])
        s = C.CDLL(_SHIM_SO)
        s.skh_collector_new.restype = C.c_void_p
        s.skh_collector_clear.argtypes = [C.c_void_p]
        s.skh_collector_free.argtypes = [C.c_void_p]
        s.skh_collector_count.restype = C.c_size_t
        s.skh_collector_count.argtypes = [C.c_void_p]
        s.skh_collector_get.restype = C.POINTER(SkhOut)
        s.skh_collector_get.argtypes = [C.c_void_p, C.c_size_t]
        _shim = s
    return _shim


class CNativePluginAPI(C.Structure):
    _fields_ = [
        ("version", C.c_uint32),
        ("get_metadata", C.CFUNCTYPE(C.POINTER(CNodeMetadata))),
        ("create_instance", C.CFUNCTYPE(C.c_void_p, C.c_char_p, LOG_CB, C.c_void_p)),
        # ABI v2: 7 arguments (types.rs:229-237)
        ("process_packet", C.CFUNCTYPE(CResult, C.c_void_p, C.c_char_p, C.POINTER(CPacket), OUT_CB, C.c_void_p, TEL_CB, C.c_void_p)),
        ("update_params", C.CFUNCTYPE(CResult, C.c_void_p, C.c_char_p)),
        # ABI v2: 5 arguments (types.rs:250-256)
        ("flush", C.CFUNCTYPE(CResult, C.c_void_p, OUT_CB, C.c_void_p, TEL_CB, C.c_void_p)),
        ("destroy_instance", C.CFUNCTYPE(None, C.c_void_p)),
    ]


class PluginError(RuntimeError):
    pass


class NativePlugin:
    """LoadedNativePlugin (crates/plugin-native/src/lib.rs:50-103)"""

    def __init__(self, path: str):
        self.lib = C.CDLL(path)
        entry = self.lib.streamkit_native_plugin_api
        entry.restype = C.POINTER(CNativePluginAPI)
        self.api = entry().contents
        if self.api.version != API_VERSION:  # lib.rs:90-95
            raise PluginError(f"API version mismatch: {self.api.version} != {API_VERSION}")
        md = self.api.get_metadata().contents
        self.kind = md.kind.decode()
        self.description = md.description.decode() if md.description else None
        self.param_schema = md.param_schema.decode() if md.param_schema else None
        self.inputs = [md.inputs[i].name.decode() for i in range(md.inputs_count)]
        self.outputs = [md.outputs[i].name.decode() for i in range(md.outputs_count)]
        self.input_formats = []
        for i in range(md.inputs_count):
            pin = md.inputs[i]
            for j in range(pin.accepts_types_count):
                ti = pin.accepts_types[j]
                fmt = ti.audio_format.contents if ti.audio_format else None
                self.input_formats.append((ti.type_discriminant, (fmt.sample_rate, fmt.channels, fmt.sample_format) if fmt else None))
        self.categories = [md.categories[i].decode() for i in range(md.categories_count)]

    def create(self, params_json: str | None):
        return PluginInstance(self, params_json)


class PluginInstance:
    """NativeNodeWrapper (wrapper.rs): sequential calls per instance."""

    def __init__(self, plugin: NativePlugin, params_json: str | None):
        self.plugin = plugin
        self.logs = []
        self._log_cb = LOG_CB(lambda lvl, tgt, msg, ud: self.logs.append((lvl, (tgt or b"").decode(), (msg or b"").decode())))
        self.h = plugin.api.create_instance(params_json.encode() if params_json is not None else None, self._log_cb, None)
        if not self.h:
            raise PluginError("Plugin failed to create instance")  # wrapper.rs:184-188
        sh = shim()
        self._coll = C.c_void_p(sh.skh_collector_new())
        self._out_cb = C.cast(sh.skh_output_cb, C.c_void_p)
        self._tel_cb = C.cast(sh.skh_telemetry_cb, C.c_void_p)

    def _drain(self):
        sh = shim()
        outs = []
        for i in range(sh.skh_collector_count(self._coll)):
            o = sh.skh_collector_get(self._coll, i).contents
            pin = o.pin.decode()
            if o.packet_type == PACKET_RAW_AUDIO:
                samples = np.frombuffer(C.string_at(o.data, o.n * 4), dtype=np.float32).copy()
                outs.append((pin, dict(kind="audio", sample_rate=o.sample_rate, channels=o.channels, samples=samples)))
            elif o.packet_type == PACKET_BINARY:
                outs.append((pin, dict(kind="binary", data=C.string_at(o.data, o.n))))
            else:
                outs.append((pin, dict(kind="other", packet_type=o.packet_type)))
        sh.skh_collector_clear(self._coll)
        return outs

    def process_audio(self, sample_rate: int, channels: int, samples: np.ndarray):
        s = np.ascontiguousarray(samples, dtype=np.float32).ravel()
        fr = CAudioFrame(sample_rate, channels, s.ctypes.data_as(C.POINTER(C.c_float)), s.size)
        pkt = CPacket(PACKET_RAW_AUDIO, C.cast(C.pointer(fr), C.c_void_p), C.sizeof(CAudioFrame))
        res = self.plugin.api.process_packet(self.h, b"in", C.byref(pkt), self._out_cb, self._coll, self._tel_cb, self._coll)
        if not res.success:
            raise PluginError((res.error_message or b"Unknown plugin error").decode())
        return self._drain()

    def process_binary(self, data: bytes):
        buf = C.create_string_buffer(data, len(data))
        pkt = CPacket(PACKET_BINARY, C.cast(buf, C.c_void_p), len(data))
        res = self.plugin.api.process_packet(self.h, b"in", C.byref(pkt), self._out_cb, self._coll, self._tel_cb, self._coll)
        if not res.success:
            raise PluginError((res.error_message or b"Unknown plugin error").decode())
        return self._drain()

    def update_params(self, params_json: str | None):
        res = self.plugin.api.update_params(self.h, params_json.encode() if params_json is not None else None)
        return res.success, (res.error_message.decode() if (not res.success and res.error_message) else None)

    def flush(self):
        res = self.plugin.api.flush(self.h, self._out_cb, self._coll, self._tel_cb, self._coll)
        if not res.success:
            raise PluginError((res.error_message or b"Unknown plugin error").decode())
        return self._drain()

    def destroy(self):
        if self.h:
            self.plugin.api.destroy_instance(self.h)
            self.h = None
            shim().skh_collector_free(self._coll)
