"""bench.py contract checks that need no GPU: the reference arm (the CPU port of the reference's nodes, the one place besides
the tests where bench.py may execute oracle/) prints ONE JSON line with the keys the driver reads, for the headline workload and
for the node configs; a failing CUDA arm must not be silently replaced by it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    return p, lines


@pytest.mark.parametrize("extra", [[], ["--config", "2"], ["--config", "3"], ["--config", "4"]])
def test_reference_arm_prints_one_json_line_with_the_contract_keys(extra):
    p, lines = _run("--impl", "reference", "--steps", "2", "--warmup", "1", *extra)
    assert p.returncode == 0, p.stderr[-500:]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    for k in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "timed_sample" in d["config"]
    assert d["gpu_launches"] == 0


def test_cuda_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: on a box without CUDA the product arm exits non-zero and prints no result line"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p, lines = _run("--steps", "1", "--warmup", "1", "--no-hub", "--no-router")
    assert p.returncode != 0
    assert not any(ln.lstrip().startswith("{") and '"value"' in ln for ln in lines)


def test_reference_arm_under_torchrun_prints_from_rank_0_only():
    """N > 1: the driver launches the reference arm under torchrun like the CUDA arm; rank 0 alone runs and prints, the others exit 0"""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-800:]
    lines = [ln for ln in p.stdout.splitlines() if ln.lstrip().startswith("{")]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2
