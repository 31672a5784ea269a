"""CPU tests of the C-ABI boundary: the shared library loads without a GPU, exports every symbol that
include/skgpu_batch.h declares, struct layouts match the ctypes mirrors, and without a CUDA device the
product path FAILS LOUDLY (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from streamkit_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared("skgpu_batch.h")
    assert len(names) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/skgpu_batch.h but not exported: {missing}"
    assert sorted(L.EXPORTS) == names, "streamkit_b200/lib.py binds a different symbol set than the header declares"


def test_library_loads_and_reports_version():
    lib = L.load()
    assert lib.skgpu_abi_version() == 1


def test_struct_layouts_match_header():
    """compile a tiny C program against the header and compare sizeof/offsetof with the numpy/ctypes mirrors"""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "skgpu_batch.h"
int main(void){
 printf("%zu %zu %zu %zu\n", sizeof(skgpu_seg), offsetof(skgpu_seg,out_off), offsetof(skgpu_seg,n_samples), offsetof(skgpu_seg,gain_idx));
 printf("%zu %zu %zu %zu %zu\n", sizeof(skgpu_rs_item), offsetof(skgpu_rs_item,out_off), offsetof(skgpu_rs_item,slot), offsetof(skgpu_rs_item,out_cap_frames), offsetof(skgpu_rs_item,flags));
 printf("%zu %zu %zu %zu %zu %zu\n", sizeof(skgpu_mix_input), offsetof(skgpu_mix_input,n_frames), offsetof(skgpu_mix_input,channels), offsetof(skgpu_mix_input,flags), offsetof(skgpu_mix_input,gain_idx), offsetof(skgpu_mix_input,slot));
 printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(skgpu_mix_group), offsetof(skgpu_mix_group,first_input), offsetof(skgpu_mix_group,n_inputs), offsetof(skgpu_mix_group,out_frames), offsetof(skgpu_mix_group,out_channels), offsetof(skgpu_mix_group,flags), offsetof(skgpu_mix_group,gain_idx));
 printf("%zu %zu %zu %zu\n", sizeof(skgpu_rs_result), sizeof(skgpu_ctx_config), sizeof(skgpu_stream_cfg), sizeof(skgpu_tick_timing));
 return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        rows = [[int(v) for v in ln.split()] for ln in subprocess.check_output([os.path.join(d, "t")], text=True).splitlines()]
    dt = L.SEG_DT
    assert rows[0] == [dt.itemsize, dt.fields["out_off"][1], dt.fields["n_samples"][1], dt.fields["gain_idx"][1]]
    dt = L.RS_ITEM_DT
    assert rows[1] == [dt.itemsize] + [dt.fields[k][1] for k in ("out_off", "slot", "out_cap_frames", "flags")]
    dt = L.MIX_INPUT_DT
    assert rows[2] == [dt.itemsize] + [dt.fields[k][1] for k in ("n_frames", "channels", "flags", "gain_idx", "slot")]
    dt = L.MIX_GROUP_DT
    assert rows[3] == [dt.itemsize] + [dt.fields[k][1] for k in ("first_input", "n_inputs", "out_frames", "out_channels", "flags", "gain_idx")]
    assert rows[4] == [L.RS_RESULT_DT.itemsize, C.sizeof(L.CtxConfig), C.sizeof(L.StreamCfg), C.sizeof(L.TickTiming)]


def _gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_gpu_present(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    with pytest.raises(L.SkgpuError) as e:
        L.Context(device=0, max_streams=4)
    assert e.value.rc == -5 and "no CPU fallback" in e.value.msg


def test_product_package_never_imports_oracle():
    """parity claims are void if the product routes through the oracle: no file under streamkit_b200/ may mention it"""
    bad = []
    for dp, _dn, fn in os.walk(os.path.join(ROOT, "streamkit_b200")):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"\boracle\b|sk_oracle|sko_", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_native_abi_header_matches_rust_layout():
    """include/streamkit_native_abi.h: sizes the Rust #[repr(C)] structs have on x86-64 (types.rs:38-261)"""
    prog = r'''
#include <stdio.h>
#include "streamkit_native_abi.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(sk_result), sizeof(sk_audio_frame), sizeof(sk_packet),
  sizeof(sk_audio_format), sizeof(sk_packet_type_info), sizeof(sk_node_metadata), sizeof(sk_native_plugin_api), sizeof(sk_packet_metadata)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(v) for v in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    assert got == [16, 24, 24, 12, 24, 72, 56, 48]


def test_plugin_libraries_export_the_entry_point_and_metadata_without_a_gpu():
    """the host reads plugin metadata before any instance exists (plugin-native/src/lib.rs:107): that must work on a box
    without a GPU; creating an instance there must fail (NULL handle), never fall back to the CPU"""
    from tests import plugin_host as ph
    csrc = os.path.join(ROOT, "streamkit_b200", "csrc")
    for so, kind in (("libskgpu_plugin_gain.so", "gpu_gain"), ("libskgpu_plugin_resampler.so", "gpu_resampler"),
                     ("libskgpu_plugin_pcm16.so", "gpu_pcm16")):
        p = ph.NativePlugin(os.path.join(csrc, so))           # checks version == 2 like lib.rs:90-95
        assert p.kind == kind and p.inputs == ["in"] and p.outputs == ["out"]
        assert "::" not in p.kind and not p.kind.startswith("core")   # plugin kind rules (lib.rs:307-333)
        import json
        json.loads(p.param_schema)
        if not _gpu_present():
            with pytest.raises(ph.PluginError, match="failed to create instance"):
                p.create('{"gain": 1.0, "target_sample_rate": 48000}')


def test_abi_v3_header_compiles_as_c_and_v3_plugin_exports_version_3():
    import tempfile
    root = ROOT
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write('#include "streamkit_native_abi_v3.h"\nint main(void){ return sizeof(sk_native_plugin_api_v3) == 80 && sizeof(sk_packet_v3) == 32 ? 0 : 1; }\n')
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        assert subprocess.call([os.path.join(d, "t")]) == 0
    lib = C.CDLL(os.path.join(ROOT, "streamkit_b200", "csrc", "libskgpu_plugin_mixer_v3.so"))
    lib.streamkit_native_plugin_api.restype = C.POINTER(C.c_uint32)
    assert lib.streamkit_native_plugin_api()[0] == 3


