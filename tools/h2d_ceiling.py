#!/usr/bin/env python
"""tools/h2d_ceiling.py -- what the BOX can do: plain cudaMemcpyAsync from pinned host memory, N = 1, 2, 4, 8 GPUs
concurrently (one thread + one stream per GPU, no kernels, nothing of streamkit_b200 involved), H2D alone, D2H alone
and both directions at once, with the pinned buffers (a) allocated wherever the calling thread happens to run
("unbound") and (b) allocated by a thread pinned to the CPUs of the GPU's own NUMA node ("numa"). The end-to-end
sessions/GPU of bench.py cannot exceed  h2d_GBps / (K * 7056 B * 50/s).

Usage: python tools/h2d_ceiling.py [--mb 1024] [--reps 8] > profiles/r2_h2d_ceiling.json
"""
from __future__ import annotations

import argparse
import json
import os
import threading
import time

import torch


def gpu_numa_node(i: int) -> int:
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        dev = torch.cuda.get_device_properties(i).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        return int(open(path).read().strip())
    except Exception:
        return -1


def node_cpus(node: int):
    try:
        s = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    except Exception:
        return None
    cpus = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    allowed = os.sched_getaffinity(0)
    cpus &= allowed
    return cpus or None


def run(n_gpus: int, mb: int, reps: int, numa: bool, direction: str) -> dict:
    nbytes = mb << 20
    res = [None] * n_gpus
    barrier = threading.Barrier(n_gpus)
    all_cpus = os.sched_getaffinity(0)

    def worker(i: int):
        cpus = node_cpus(gpu_numa_node(i)) if numa else None
        if cpus:
            os.sched_setaffinity(0, cpus)          # thread-local on Linux: allocation + first touch happen on the GPU's node
        torch.cuda.set_device(i)
        h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(nbytes // 4, dtype=torch.uint8, pin_memory=True)
        h_in.fill_(1)
        h_out.fill_(0)
        d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(nbytes // 4, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        barrier.wait()
        if direction in ("h2d", "both"):
            with torch.cuda.stream(s1):
                e0.record()
                for _ in range(reps):
                    d_in.copy_(h_in, non_blocking=True)
                e1.record()
        if direction in ("d2h", "both"):
            with torch.cuda.stream(s2):
                f0.record()
                for _ in range(reps):
                    h_out.copy_(d_out, non_blocking=True)
                f1.record()
        torch.cuda.synchronize()
        out = {"gpu": i, "numa_node": gpu_numa_node(i), "cpus_bound": len(cpus) if cpus else 0}
        if direction in ("h2d", "both"):
            out["h2d_gbs"] = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        if direction in ("d2h", "both"):
            out["d2h_gbs"] = (nbytes // 4) * reps / (f0.elapsed_time(f1) * 1e-3) / 1e9
        res[i] = out
        if cpus:
            os.sched_setaffinity(0, all_cpus)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_gpus)]
    t0 = time.time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    agg = {"n_gpus": n_gpus, "numa_bound": numa, "direction": direction, "per_gpu": res, "wall_s": time.time() - t0}
    if direction in ("h2d", "both"):
        agg["h2d_gbs_total"] = sum(r["h2d_gbs"] for r in res)
        agg["h2d_gbs_min"] = min(r["h2d_gbs"] for r in res)
    if direction in ("d2h", "both"):
        agg["d2h_gbs_total"] = sum(r["d2h_gbs"] for r in res)
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    out = {"host": {"cpus": len(os.sched_getaffinity(0)), "numa_nodes": sorted(
        int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()) if os.path.isdir("/sys/devices/system/node") else []},
        "gpus": n, "gpu_numa": [gpu_numa_node(i) for i in range(n)], "runs": []}
    for k in (1, 2, 4, 8):
        if k > n:
            break
        for numa in (False, True):
            for direction in ("h2d", "both"):
                out["runs"].append(run(k, args.mb, args.reps, numa, direction))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
