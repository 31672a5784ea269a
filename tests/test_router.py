"""Single-process multi-GPU session router (include/skgpu_router.h): the routing function is the reference's own session
hash (apps/skit/src/session.rs:35-45); sessions never leave their GPU (no collective)."""
import os
import re
import subprocess
import uuid

import numpy as np
import pytest

from streamkit_b200 import router as R, shard
from tests.test_hub import _OracleSession, _chunk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_router_library_exports_every_declared_symbol():
    src = open(os.path.join(ROOT, "include", "skgpu_router.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(skgpu_(?:router_[a-z0-9_]+|fnv1a64))\s*\(", src)))
    out = subprocess.check_output(["nm", "-D", "--defined-only", R.ROUTER_LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert not [n for n in names if n not in exported]
    assert sorted(R.EXPORTS) == names


def test_routing_is_the_reference_session_hash():
    # FNV-1a 64 known answers, and agreement with the pure-Python statement of session.rs:35-45
    assert R.fnv1a64(b"") == 0xCBF29CE484222325
    assert R.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    ids = [str(uuid.UUID(int=i * 104729 + 7)) for i in range(5000)]
    for n in (1, 2, 4, 8):
        got = [R.gpu_for(s, n) for s in ids]
        assert got == [shard.gpu_for_session(s, n) for s in ids]
        counts = np.bincount(got, minlength=n)
        assert counts.max() - counts.min() < 0.15 * len(ids) / n + 40      # UUIDs spread evenly


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_router_sessions_match_oracle_on_every_gpu():
    """sessions opened by id land on fnv1a64(id) % n_gpus, are ticked by that GPU's own thread and deliver the oracle's bytes"""
    n = max(1, _n_gpus())
    r = R.Router(list(range(n)), max_sessions=32, max_streams=64, in_rates=[44100, 48000], max_inputs_per_session=2)
    try:
        assert len(r.numa_nodes()) == n
        ids = [str(uuid.UUID(int=i * 31337 + 5)) for i in range(12)]
        shapes = [[44100, 48000] if i % 3 else [44100, 44100] for i in range(len(ids))]
        handles = [r.session_open(s, sh) for s, sh in zip(ids, shapes)]
        for s, h in zip(ids, handles):
            assert h >> 32 == shard.gpu_for_session(s, n)
        osess = [_OracleSession(sh, 2, 960) for sh in shapes]
        r.set_master_gain(handles[3], 0.5)
        osess[3].master = 0.5
        for t in range(6):
            want = []
            for a, (h, o, sh) in enumerate(zip(handles, osess, shapes)):
                for i, rate in enumerate(sh):
                    x = _chunk(900 + a * 4 + i, t, rate, rate * 960 // 48000, 2)
                    r.push(h, i, x)
                    o.push(i, x)
                want.append(o.tick())
            r.tick()
            r.wait()
            for h, (w, nm) in zip(handles, want):
                got, n_mixed, status = r.output(h)
                assert status == 0 and n_mixed == nm and np.array_equal(got, w), (t, hex(h))
        ms = r.run_ticks(5)                       # the zero-copy steady-state loop used by bench.py
        assert ms > 0
        r.session_close(handles[0])
        with pytest.raises(R.RouterError):
            r.push(handles[0], 0, np.zeros(882 * 2, np.float32))
    finally:
        r.close()
