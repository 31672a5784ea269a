"""Session -> GPU sharding: the unit is the mix group (= session); all of a session's streams live on one
GPU, so there is no exchange step and no collective on the data path. torch.distributed is used by bench.py only
for the barrier and the max-over-ranks of the device time."""
from __future__ import annotations


def fnv1a64(data: bytes) -> int:
    """FNV-1a 64-bit, the hash the reference already applies to session ids (apps/skit/src/session.rs:35-45)."""
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def gpu_for_session(session_id: str, n_gpus: int) -> int:
    return fnv1a64(session_id.encode("utf-8")) % n_gpus


def partition(session_ids, n_gpus: int):
    """rank -> list of session ids; every session appears exactly once"""
    parts = [[] for _ in range(n_gpus)]
    for s in session_ids:
        parts[gpu_for_session(s, n_gpus)].append(s)
    return parts


def aggregate(per_rank_sessions: int, world: int, max_ms_per_step: float, tick_ms: float = 20.0) -> float:
    """whole-job metric: real-time sessions = (sessions processed by all ranks per tick) * tick / slowest rank's step"""
    return per_rank_sessions * world * tick_ms / max_ms_per_step
