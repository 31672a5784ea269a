#!/usr/bin/env python
"""bench.py -- benchmark of streamkit_b200 (driver contract in the task statement).

Default workload (--config 5 = BASELINE.json configs[4], the configuration the metric is quoted on; fits one GPU):
    full chain resample 44.1k->48k -> per-input gain -> ordered mix -> master gain -> clip -> s16,
    K (--k, default 2) stereo f32 inputs per session, --sessions per GPU, one 20 ms tick per step.
Metric: concurrent real-time 48 kHz stereo sessions per GPU whose 20 ms tick costs <= 2 ms of device time.

  value        device-resident (inputs already in HBM when the timed region starts), CUDA events, max over ranks:
               sessions x 2 ms / ms_per_step. `capacity_check` then RUNS that many sessions and reports the measured p99.
  e2e          the same tick through the C ABI with HOST (pinned) buffers, SLICED (skgpu_plan_auto_slices): upload, kernels and
               read-back of consecutive slices overlap on three streams, two ticks in flight. H2D of every input frame and D2H
               of every s16 result are inside the timed region. `e2e.latency` = per slice upload-done -> read-back-done
               (SURVEY 8d's added device latency), worst slice of each tick, p50 / p99 over --latency-ticks ticks.
  roofline     dominant kernel, algorithmic bytes / CUDA-event time vs the measured HBM peak; `traffic` is read from the
               committed ncu capture of the same workload (profiles/r2_ncu_full_summary.csv), not a constant in this file
  parity       sampled sessions of the last end-to-end tick compared with the CPU chain (oracle/sk_chain.c) after the same
               number of ticks -- the oracle is the checker here, never the thing measured
  cpu_baseline the oracle's reference-shaped CPU run of the same workload on the box's host cores

--config 2 | 3 | 4 run the standalone node workloads (BASELINE configs[1..3]) with the same JSON contract.
`--impl reference` times the CPU arm (the reference is Rust and cannot be built in this image: DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IN_RATE, OUT_RATE, CHANNELS = 44100, 48000, 2
TICK_MS, BUDGET_MS = 20.0, 2.0
METRIC = "concurrent real-time 48 kHz stereo sessions per GPU within the 2 ms device budget per 20 ms tick (resample->mix->gain->s16)"
UNIT = "sessions"


S16_IN = False


def chain_config(sessions_per_gpu: int, n_gpus: int, k: int) -> dict:
    if IN_RATE == OUT_RATE:
        wl = ("Opus-shaped variant of BASELINE configs[4] (samples/pipelines/dynamic/moq_mixing.yml): %d x 48 kHz %s decoder outputs -> gain -> "
              "%d-input ordered mix -> gain -> clip -> s16; the resampler node bypasses rate-equal inputs (resampler.rs:299-373)") % (
            k, "mono" if CHANNELS == 1 else "stereo", k)
    else:
        wl = "BASELINE configs[4]: full chain resample %.1fk->48k -> gain -> %d-input ordered mix -> gain -> clip -> s16" % (IN_RATE / 1e3, k)
    mb = sessions_per_gpu * k * (IN_RATE // 50) * CHANNELS * (2 if S16_IN else 4) / 1e6
    return {
        "workload": wl + (" (inputs arrive as s16 on PCIe)" if S16_IN else ""),
        "sessions_per_gpu": sessions_per_gpu, "inputs_per_session": k,
        "in_rate": IN_RATE, "out_rate": OUT_RATE, "channels": CHANNELS, "tick_ms": TICK_MS, "device_budget_ms": BUDGET_MS,
        "chunk_frames": IN_RATE // 50, "output_frame_size": 960, "input_format": "s16" if S16_IN else "f32",
        "sharding": "sessions split evenly across %d GPU(s) by session id, no collective" % n_gpus,
        "l2": ("inputs are %.0f MB per tick per GPU (> 126 MB L2): no flush needed" % mb) if mb > 126 else
              ("inputs are %.0f MB per tick per GPU, double-banked (two ticks' worth alternate) -- may partly stay in the 126 MB L2" % mb),
    }


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("SK_BENCH_CLOCK_LMS", "100")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mx = float(f[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:  # timed region shorter than the sampling period: take every sample we have
            for ts, ln in self.rows:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic(kernel_substr: str, grid_hint: str | None = None):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed `ncu --set full` capture of THIS bench
    command (profiles/r2_ncu_full_summary.csv, written by tools/export_profiles.sh). Returns (bytes per launch, source) or
    (None, why)."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_full_summary.csv")
    if not os.path.exists(path):
        return None, "no committed capture (profiles/r2_ncu_full_summary.csv)"
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel_substr in d.get("Kernel Name", ""):
            u = dict(zip(hdr, units))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            try:
                tot = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            except (KeyError, ValueError):
                return None, "capture lacks dram__bytes"
            return int(tot), "profiles/r2_ncu_full_summary.csv (ncu --set full, per launch)"
    return None, "kernel not in the committed capture"


# ---------------------------------------------------------------------------------------------- CPU arm

def cpu_chain(n_sessions: int, k: int, ticks: int, threads: int, seed: int = 0, pool=None, ig=None, mg=None, want_last=False):
    """the oracle's reference-shaped chain on host cores; returns (seconds for `ticks` ticks of n_sessions, last outputs)"""
    from oracle import sko
    from streamkit_b200 import synth

    if pool is None:
        pool = synth.noise_streams(seed, 0, 256, IN_RATE // 50, CHANNELS)
        ig = synth.gains(seed, n_sessions * k, 0.25, 1.5)
        mg = synth.gains(seed + 1, n_sessions, 0.5, 2.0)
    sec, _cs, last = sko.chain_bench(n_sessions, k, ticks, IN_RATE, CHANNELS, pool, ig, mg, threads, want_last=want_last)
    return sec, last


def cpu_node(config: int, units: int, iters: int, threads: int):
    """reference-shaped CPU run of a standalone node workload; returns seconds"""
    from oracle import sko

    rng = np.random.default_rng(3)
    if config == 2:
        pool = (rng.random((512, 1920), dtype=np.float32) * 2 - 1).astype(np.float32)
        return sko.node_bench(0, units, iters, pool, rng.random(units, dtype=np.float32) * 4, threads)[0]
    if config == 3:
        pool = (rng.random((2048, 1920), dtype=np.float32) * 2 - 1).astype(np.float32)
        return sko.node_bench(1, units, iters, pool, rng.random(units, dtype=np.float32), threads, k_inputs=64)[0]
    pool = (rng.random((256, 882 * 2), dtype=np.float32) - 0.5).astype(np.float32)
    return sko.node_bench(2, units, iters, pool, None, threads, in_rate=IN_RATE, out_rate=OUT_RATE, cap=1000)[0]


NODE_META = {
    2: ("concurrent real-time 48 kHz stereo sessions per GPU within the 2 ms device budget (gain + f32->s16)", "sessions", 512 * 16),
    3: ("64-input mix groups per 20 ms tick per GPU within the 2 ms device budget (ordered sum + gain + clip + s16)", "mix groups", 256),
    4: ("concurrent real-time stereo resampler streams 44.1k->48k per GPU within the 2 ms device budget", "streams", 8192),
}


def node_config(config: int, n_gpus: int) -> dict:
    w = {2: "BASELINE configs[1]: gain + f32<->s16 conversion over 4,096 concurrent 48 kHz stereo sessions, 20 ms frames",
         3: "BASELINE configs[2]: N-input mixer (64 inputs per mix, with clip) x 1,024 mix groups per tick",
         4: "BASELINE configs[3]: batched resampling 44.1k->48k over 16,384 stereo streams (48k->16k: --rs-down)"}[config]
    return {"workload": w, "tick_ms": TICK_MS, "device_budget_ms": BUDGET_MS,
            "sharding": "units split evenly across %d GPU(s), no collective" % n_gpus,
            "l2": "device working set > 126 MB L2 (8 rotated tick buffers for configs[1]): no flush needed"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    if args.config == 5:
        n = args.ref_sessions
        cpu_chain(n, args.k, max(args.warmup, 1), cores)
        sec, _ = cpu_chain(n, args.k, args.steps, cores)
        ms_per_step = sec * 1e3 / args.steps
        value = n * TICK_MS / ms_per_step       # sessions the host cores sustain in real time (the CPU has no 2 ms device budget)
        metric, unit, cfg = METRIC, UNIT, chain_config(args.sessions, args.gpus, args.k)
        sample = ("bounded sample: %d sessions x %d ticks on %d host threads (oracle/sk_chain.c, reference-shaped per-packet nodes); value = "
                  "sessions sustained in real time (20 ms per tick)") % (n, args.steps, cores)
    else:
        metric, unit, n = NODE_META[args.config]
        cpu_node(args.config, n, max(args.warmup, 1), cores)
        sec = cpu_node(args.config, n, args.steps, cores)
        ms_per_step = sec * 1e3 / args.steps
        value = n * TICK_MS / ms_per_step
        cfg = node_config(args.config, args.gpus)
        sample = "bounded sample: %d %s x %d passes on %d host threads (oracle/sk_chain.c sko_node_bench); value = units sustained in real time" % (n, unit, args.steps, cores)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": dict(cfg, timed_sample=sample),   # the workload is the GPU arm's; the CPU times a bounded sample of it
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust (no toolchain in this image): timed arm is the C restatement of its nodes; tokio scheduling"
                " and channel hops of the real engine are not included (optimistic for the reference)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- frame-batching layer

def run_hub_e2e(S: int, k: int, steps: int, warmup: int, device: int, threads: int) -> dict:
    """The same workload driven through the frame-batching layer (include/skgpu_hub.h): per tick the chunks of all
    S x K streams are gathered from DISTINCT host buffers (one private chunk per stream: the whole ~0.9 GB tick is read
    from host memory every tick, nothing is cache resident) into the hub's pinned arena by `threads` worker threads
    (skgpu_hub_push_batch), then one asynchronous sliced tick; the gather of tick n + 1 overlaps tick n on the GPU.
    Wall-clock time (it includes host work), results read back every tick."""
    from streamkit_b200 import hub as H, synth

    chunk = IN_RATE // 50
    hub = H.Hub(max_sessions=S, max_streams=S * k, in_rates=[IN_RATE], max_inputs_per_session=k, channels=CHANNELS, device=device, in_s16=S16_IN)
    try:
        n = S * k
        src = np.empty((n, chunk * CHANNELS), dtype=np.int16 if S16_IN else np.float32)   # every stream's own frame buffer (pageable, like decoder output)
        blk = synth.noise_streams(4242, 0, 4096, chunk, CHANNELS)
        if S16_IN:
            blk = np.rint(blk * 32767.0).astype(np.int16)
        for b in range(0, n, 4096):
            m = min(4096, n - b)
            src[b:b + m] = blk[:m]
            src[b:b + m, 0] += (np.arange(b, b + m) % 97).astype(src.dtype)
        frames = np.zeros(n, dtype=H.FRAME_DT)
        for s in range(S):
            sid = hub.session_open([IN_RATE] * k)
            for i in range(k):
                f = frames[s * k + i]
                f["session"], f["input"], f["n_frames"] = sid, i, chunk
        frames["samples"] = src.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(src.strides[0])
        for _ in range(max(4, warmup)):   # also fills every arena of the input ring
            hub.push_batch(frames, threads)
            hub.tick()
            hub.wait()
        t_push = t_tick = t_wait = 0.0
        t0 = time.perf_counter()
        hub.push_batch(frames, threads)
        hub.tick()
        for _ in range(steps - 1):
            a = time.perf_counter()
            hub.push_batch(frames, threads)   # gather of the next tick while the GPU works on the current one
            b = time.perf_counter()
            hub.wait()
            c = time.perf_counter()
            hub.tick()
            d = time.perf_counter()
            t_push += b - a; t_wait += c - b; t_tick += d - c
        hub.wait()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        out, n_mixed, status = hub.output(0)
        # zero-copy variant: the producers wrote their chunks straight into the pinned slots (skgpu_hub_acquire), so a tick
        # is commit + upload + kernels + read-back (both arenas hold a chunk for every stream after the loop above)
        for _ in range(2):
            hub.commit_all(); hub.tick(); hub.wait()
        t1 = time.perf_counter()
        hub.commit_all(); hub.tick()
        for _ in range(steps - 1):
            prev = hub.ticks
            hub.commit_all()
            hub.tick()             # tick n + 1 is submitted first ...
            hub.wait_tick(prev)    # ... then tick n is collected while n + 1 uploads (its read-back overlaps that upload)
        hub.wait()
        ms_zc = (time.perf_counter() - t1) * 1e3 / steps
        return {"value": S * TICK_MS / ms, "zero_copy": {"value": S * TICK_MS / ms_zc, "ms_per_step": ms_zc,
                                                        "what": "producers write in place (skgpu_hub_acquire / commit), pipelined collection (skgpu_hub_wait_tick)"}, "unit": UNIT, "ms_per_step": ms, "gather_threads": threads, "sessions": S,
                "host_ms": {"gather": t_push * 1e3 / max(steps - 1, 1), "wait": t_wait * 1e3 / max(steps - 1, 1), "tick_call": t_tick * 1e3 / max(steps - 1, 1)},
                "what": "skgpu_hub: multi-threaded gather of %.0f MB of distinct per-stream frames into the pinned arena + H2D + kernels + D2H per tick, wall clock" % (src.nbytes / 1e6),
                "stats": hub.stats(),
                "check": {"n_mixed": int(n_mixed), "status": int(status), "nonzero": bool(out is not None and np.any(out != 0))}}
    finally:
        hub.close()


def run_router_e2e(S: int, k: int, n_gpus: int, steps: int) -> dict:
    """ONE process driving all n_gpus GPUs through the session router (include/skgpu_router.h): S x n_gpus sessions opened by
    UUID, routed by fnv1a64(id) % n_gpus, one NUMA-pinned tick thread + NUMA-bound pinned arenas per GPU, zero-copy producers
    (skgpu_hub_acquire / commit), sliced ticks, pipelined collection. Wall clock of the slowest GPU's tick loop."""
    import uuid

    from streamkit_b200 import hub as H, router as R, synth

    chunk = IN_RATE // 50
    r = R.Router(list(range(n_gpus)), max_sessions=int(S * 1.08) + 64, max_streams=(int(S * 1.08) + 64) * k, in_rates=[IN_RATE],
                 max_inputs_per_session=k, channels=CHANNELS, in_s16=S16_IN)
    try:
        per_gpu = [[] for _ in range(n_gpus)]
        for i in range(S * n_gpus):
            h = r.session_open(str(uuid.UUID(int=i * 2654435761 + 12345)), [IN_RATE] * k)
            per_gpu[h >> 32].append(h & 0xFFFFFFFF)
        blk = synth.noise_streams(31337, 0, 4096, chunk, CHANNELS)
        if S16_IN:
            blk = np.rint(blk * 32767.0).astype(np.int16)
        hl = H.load()
        for g in range(n_gpus):            # fill every arena of each hub's input ring once (the producers own the slots afterwards)
            ns = len(per_gpu[g])
            frames = np.zeros(ns * k, dtype=H.FRAME_DT)
            frames["session"] = np.repeat(np.asarray(per_gpu[g], dtype=np.uint32), k)
            frames["input"] = np.tile(np.arange(k, dtype=np.uint32), ns)
            frames["n_frames"] = chunk
            frames["samples"] = blk.ctypes.data + (np.arange(frames.size, dtype=np.uint64) % 4096) * np.uint64(blk.strides[0])
            for _ in range(3):
                rc = hl.skgpu_hub_push_batch(r.hub_handle(g), frames.ctypes.data_as(C.c_void_p), frames.size, 8)
                assert rc == 0
                r.tick()
                r.wait()
        r.run_ticks(3)
        ms = r.run_ticks(steps)
        out, n_mixed, status = r.output((0 << 32) | per_gpu[0][0])
        return {"value": S * n_gpus * TICK_MS / ms, "unit": UNIT, "ms_per_step": ms, "n_gpus": n_gpus, "sessions": S * n_gpus,
                "sessions_per_gpu": [len(x) for x in per_gpu], "numa_nodes": r.numa_nodes(),
                "what": "single process, skgpu_router: fnv1a64 session routing, one NUMA-pinned tick thread and NUMA-bound pinned arenas per GPU, "
                        "zero-copy producers, sliced ticks, pipelined collection; wall clock per tick of the slowest GPU",
                "check": {"n_mixed": int(n_mixed), "status": int(status), "nonzero": bool(out is not None and np.any(out != 0))}}
    finally:
        r.close()


# ---------------------------------------------------------------------------------------------- GPU arm

class Dist:
    def __init__(self):
        # rank 0 prints exactly ONE line on stdout. NCCL (and anything else written from C) goes to fd 1 whenever it likes, so the
        # real stdout is parked on a spare descriptor and fd 1 points at stderr until the JSON line is ready (emit()).
        sys.stdout.flush()
        self._real_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; streamkit_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist  # plumbing only: barrier + max-reduce of the timings (no data-path collective)

            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def emit(self, line: dict):
        sys.stdout.flush()
        os.write(self._real_stdout, (json.dumps(line) + "\n").encode())

    def pcie_ceiling(self, mb: int = 512, reps: int = 6) -> dict:
        """what the BOX allows, measured in the same run: plain cudaMemcpyAsync from / to pinned memory on every rank at the same time
        (upload and read-back concurrently, 4 : 1 like the workload, no kernels, nothing of streamkit_b200 involved)"""
        torch = self.torch
        h_in = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty((mb << 20) // 4, dtype=torch.uint8, pin_memory=True)
        d_in = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
        d_out = torch.empty((mb << 20) // 4, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(2):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        self.barrier()
        with torch.cuda.stream(s1):
            ev[0].record()
            for _ in range(reps):
                d_in.copy_(h_in, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s2):
            ev[2].record()
            for _ in range(reps):
                h_out.copy_(d_out, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
        h2d = (mb << 20) * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        d2h = ((mb << 20) // 4) * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9
        h2d_min = -self.max(-h2d)
        self.barrier()
        return {"h2d_gbs_per_gpu_min": h2d_min, "h2d_gbs_this_gpu": h2d, "d2h_gbs_this_gpu": d2h, "ranks_copying_at_once": self.world,
                "what": "plain cudaMemcpyAsync pinned->device (and device->pinned, a quarter of the bytes) on all ranks simultaneously"}

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(len(v) * q))]


def capacity_check(S2: int, k: int, device: int, kslices: int, ticks: int = 60) -> dict:
    """run S2 sessions (the claimed `value`) device-resident, tick by tick, and report the measured per-tick device time"""
    from streamkit_b200 import chain, lib as L, synth

    ct = chain.ChainTick(S2, k, in_rate=IN_RATE, channels=CHANNELS, device=device, seed=5, alloc_host=False, s16=S16_IN)
    try:
        blk = synth.noise_streams(77, 0, 8192, ct.chunk, CHANNELS)
        tile = np.zeros((8192, ct.in_stride), np.uint8)
        src = np.rint(blk * 32767.0).astype(np.int16) if S16_IN else blk
        tile[:, : src.shape[1] * src.itemsize] = src.view(np.uint8).reshape(8192, -1)
        for bank in (0, ct.bank_stride):
            for b in range(0, ct.n_streams, 8192):
                m = min(8192, ct.n_streams - b)
                ct.plan.upload(bank + b * ct.in_stride, tile[:m])
        flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_GRAPH
        if kslices > 1:
            ct.plan.auto_slices(ct.op_chain, kslices)
            flags |= L.SUBMIT_SLICED
        for _ in range(3):
            ct.plan.submit(None, None, flags)
        ct.plan.wait()
        lat = []
        for _ in range(ticks):
            ct.ctx.timer_start()
            ct.plan.submit(None, None, flags)
            ct.ctx.timer_stop()
            lat.append(ct.ctx.timer_ms())
        ct.plan.wait()
        res = ct.results()
        return {"sessions": S2, "ticks": ticks, "p50_ms": pct(lat, 0.5), "p99_ms": pct(lat, 0.99), "max_ms": max(lat),
                "within_budget": pct(lat, 0.99) <= BUDGET_MS, "all_emitted": bool(np.all(res["emitted"] == 1) and np.all(res["status"] == 0)),
                "what": "the claimed session count run for real: device time of one whole tick's kernels, tick by tick"}
    finally:
        ct.close()


def auto_slices(ct, args, d2h_gbs) -> int:
    """slices per end-to-end tick: --slices N, or (default) as many as keep ONE slice's read-back near 0.5 ms at the device->host
    rate this box sustains with every rank copying (a slice's results cannot reach the host faster than that), at least 16"""
    if args.slices > 0:
        return max(1, min(args.slices, ct.S))
    n = 16
    if d2h_gbs and d2h_gbs > 0:
        n = int(np.ceil(ct.out_bytes / (d2h_gbs * 1e9 * 0.5e-3)))
    return max(16, min(n, 128, ct.S))


def e2e_sliced(ct, args, D, latency_ticks=None, d2h_gbs=None):
    """end-to-end through the C ABI with host buffers: SLICED ticks, two in flight; every step uploads its inputs from pinned host
    memory and reads every s16 result back. Returns (dict, host buffer holding the last tick's outputs)."""
    from streamkit_b200 import lib as L

    plan, ctx, S = ct.plan, ct.ctx, ct.S
    lt = args.latency_ticks if latency_ticks is None else latency_ticks
    n_sl = auto_slices(ct, args, d2h_gbs)
    plan.auto_slices(ct.op_chain, n_sl)
    ins = [ct.host_in, ctx.pinned(ct.in_bytes, np.float32)]
    outs = [ct.host_out, ctx.pinned(ct.out_bytes, np.int16)]
    ins[1][:] = ins[0]
    fl = L.SUBMIT_SLICED

    def pipelined(n_ticks, collect):
        first = plan.tick_count()
        for t in range(n_ticks):
            plan.submit(ins[t & 1], outs[t & 1], fl)
            if t >= 1:
                plan.wait_for(first + t)              # the previous tick: collected while this one uploads
                if collect is not None:
                    collect(plan.slice_timing(first + t))
        plan.wait()
        if collect is not None:
            collect(plan.slice_timing(first + n_ticks))

    pipelined(max(3, min(args.warmup, 4)), None)
    D.barrier()
    ctx.timer_start()
    pipelined(args.steps, None)
    ctx.timer_stop()
    e2e_ms = ctx.timer_ms()
    D.barrier()
    e2e_ms = D.max(e2e_ms)
    worst, kern, updone = [], [], []

    dbg = []

    def collect(tm):
        if os.environ.get("SK_BENCH_LAT_DEBUG"):
            i = max(range(len(tm)), key=lambda j: tm[j][2])
            dbg.append((tm[i][2], len(dbg), i, tm[i][0], tm[i][1]))
        worst.append(max(t[2] for t in tm))
        kern.append(max(t[1] for t in tm))
        updone.append(tm[-1][0])

    if lt > 0:
        pipelined(lt, collect)
    if dbg:
        print("[lat debug] worst slices (latency ms, tick, slice, upload_done ms, kernels ms):", sorted(dbg, reverse=True)[:12], file=sys.stderr)
    last_out = outs[(lt - 1) & 1] if lt > 0 else outs[(args.steps - 1) & 1]
    e2e_ms_per_step = e2e_ms / args.steps
    e2e = {"ms_per_step": e2e_ms_per_step, "slices": n_sl, "ticks_in_flight": 2,
           "h2d_gbs_per_gpu": ct.in_bytes / (e2e_ms_per_step * 1e-3) / 1e9, "d2h_gbs_per_gpu": ct.out_bytes / (e2e_ms_per_step * 1e-3) / 1e9,
           "latency": {"ticks": len(worst), "p50_ms": pct(worst, 0.5), "p99_ms": pct(worst, 0.99), "max_ms": max(worst),
                       "kernels_p99_ms": pct(kern, 0.99), "upload_of_whole_tick_ms_p50": pct(updone, 0.5),
                       "ticks_over_1ms": int(sum(1 for w in worst if w > 1.0)), "p90_ms": pct(worst, 0.9),
                       "budget_ms": BUDGET_MS, "within_budget": pct(worst, 0.99) <= BUDGET_MS,
                       "what": "per slice: its upload done -> its results in host memory (kernels + read-back, SURVEY 8d added device "
                               "latency); worst slice of each tick; CUDA events"} if worst else None,
           "host_in_to_host_out_ms": {"first_slice": (pct(updone, 0.5) / n_sl + pct(worst, 0.5)) if worst else None,
                                      "last_slice": (pct(updone, 0.5) + pct(worst, 0.5)) if worst else None,
                                      "what": "a frame handed over at the start of the tick's upload is back in host memory after this long "
                                              "(its slice's share of the upload + the slice latency)"}}
    return e2e, last_out


def run_chain(args, D: Dist) -> None:
    from streamkit_b200 import chain, lib as L, synth

    world, rank, local_rank = D.world, D.rank, D.local_rank
    S, K = args.sessions, args.k
    ct = chain.ChainTick(S, K, in_rate=IN_RATE, channels=CHANNELS, device=local_rank, seed=rank, fused=not args.unfused, s16=S16_IN)
    plan, ctx = ct.plan, ct.ctx
    ctx.bind_thread()
    # synthetic input: a tick of noise for every stream of every session on this rank (all streams distinct)
    x = synth.noise_streams(1000 + rank, 0, ct.n_streams, ct.chunk, CHANNELS)
    if S16_IN:
        from oracle import sko as _sko   # only to derive the CPU checker's input from the s16 data (x = s / 32768)
        xi = np.rint(x * 32767.0).astype(np.int16)
        ct.fill_rows(ct.host_in, xi)
        x = _sko.s16_to_f32(xi.reshape(-1)).reshape(xi.shape)
    else:
        ct.fill_rows(ct.host_in, x)
    plan.upload(0, ct.host_in)  # resident in HBM for the device-timed region
    if ct.fused:
        plan.upload(ct.bank_stride, ct.host_in)  # both input banks (the fused kernel reads the previous tick's bank)

    # ---- device-resident region: W warm-up + K timed ticks, CUDA events on the library's stream
    dev_flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_TIME_OPS
    for _ in range(args.warmup):
        plan.submit(None, None, dev_flags)
    plan.wait()
    plan.reset_op_times()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    D.barrier()
    t_wall0 = time.time()
    ctx.timer_start()
    for _ in range(args.steps):
        plan.submit(None, None, dev_flags)
    ctx.timer_stop()
    dev_ms = ctx.timer_ms()
    D.barrier()
    dev_ms = D.max(dev_ms)
    ms_per_step = dev_ms / args.steps
    if ct.fused:
        phase_ms, _ = plan.op_time(ct.op_chain, 0)
        main_ms, n_main = plan.op_time(ct.op_chain, 1)
        kernels_ms = {"k_phase_chain": phase_ms, "k_chain": main_ms}
    else:
        phase_ms, _ = plan.op_time(ct.op_rs, 0)
        main_ms, n_main = plan.op_time(ct.op_rs, 1)
        mix_ms, _ = plan.op_time(ct.op_mix, 0)
        kernels_ms = {"k_phase_prog": phase_ms, "k_resample_prog": main_ms, "k_mix+k_fifo_commit": mix_ms}
    unsliced_ms_per_step = ms_per_step
    # ---- the same tick as a SLICED kernel network (one CUDA graph): k_phase_chain of slice i + 1 runs next to k_chain of
    # slice i on a second stream instead of in front of it. This is the whole-tick figure `value` is computed from.
    if ct.fused and args.kslices > 1:
        plan.auto_slices(ct.op_chain, min(args.kslices, S))
        sl_flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_SLICED | L.SUBMIT_GRAPH
        for _ in range(args.warmup):
            plan.submit(None, None, sl_flags)
        plan.wait()
        D.barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            plan.submit(None, None, sl_flags)
        ctx.timer_stop()
        sl_ms = ctx.timer_ms()
        D.barrier()
        ms_per_step = min(ms_per_step, D.max(sl_ms) / args.steps)
        tick_flags = sl_flags
    else:
        tick_flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_GRAPH
    # the same ticks submitted and awaited one by one (what a 20 ms tick loop sees)
    lat = []
    for _ in range(min(args.latency_ticks, 200)):
        t0 = time.perf_counter()
        plan.submit(None, None, tick_flags)
        plan.wait()
        lat.append((time.perf_counter() - t0) * 1e3)
    tick_lat = {"ticks": len(lat), "p50_ms": pct(lat, 0.5), "p99_ms": pct(lat, 0.99), "max_ms": max(lat),
                "what": "one whole tick's kernels (CUDA graph), submitted and awaited tick by tick: host wall clock around submit + wait"} if lat else None

    # ---- end-to-end region: SLICED ticks, two in flight; every step uploads its inputs from pinned host memory and reads
    # every s16 result back
    e2e = {}
    ceiling = None
    if ct.fused:
        try:   # what the box allows (plain copies on every rank at once): also sizes the slices
            ceiling = D.pcie_ceiling()
        except Exception as e:
            ceiling = {"error": str(e)[:200]}
        e2e, last_out = e2e_sliced(ct, args, D, d2h_gbs=ceiling.get("d2h_gbs_this_gpu"))
    else:
        for _ in range(3):
            plan.submit(ct.host_in, ct.host_out, 0)
        plan.wait()
        D.barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            plan.submit(ct.host_in, ct.host_out, L.SUBMIT_GRAPH | L.SUBMIT_OVERLAP_D2H)
        ctx.timer_stop()
        e2e_ms = D.max(ctx.timer_ms())
        plan.wait()
        e2e_ms_per_step = e2e_ms / args.steps
        e2e = {"ms_per_step": e2e_ms_per_step}
        last_out = ct.host_out
    D.barrier()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1)

    # ---- parity of the bytes the timed run produced: sampled sessions vs the CPU chain after the same number of ticks
    parity = None
    if rank == 0 and ct.fused and args.parity_sessions > 0:
        n_ticks = int(plan.tick_count())
        ids = np.unique(np.concatenate([np.arange(min(64, S)), np.arange(max(0, S - 64), S),
                                        np.random.default_rng(1).integers(0, S, args.parity_sessions)]))
        streams = (ids[:, None] * K + np.arange(K)[None, :]).reshape(-1)
        pool = np.ascontiguousarray(x[streams])
        _sec, want = cpu_chain(len(ids), K, n_ticks, len(os.sched_getaffinity(0)), pool=pool, ig=ct.in_gains[streams], mg=ct.master_gains[ids], want_last=True)
        got = np.asarray(last_out).reshape(S, -1)[ids]
        bad = int(np.count_nonzero(np.any(got != want, axis=1)))
        parity = {"sessions_checked": int(len(ids)), "ticks": n_ticks, "sessions_differing": bad, "bit_exact": bad == 0,
                  "nonzero_samples_frac": float(np.count_nonzero(got)) / got.size,
                  "what": "s16 bytes of the last end-to-end tick vs oracle/sk_chain.c run for the same number of ticks on the same inputs"}
        if bad:
            print("bench.py: PARITY FAILURE: %d of %d sampled sessions differ from the CPU chain" % (bad, len(ids)), file=sys.stderr)

    launches_per_tick = plan.launches_per_tick()
    dev_name, dev_sms, dev_cc_ma, dev_cc_mi = ctx.device_info()
    numa = {"gpu_node": ctx.numa_node(), "pinned_node": getattr(ctx, "last_pinned_node", -1)}
    chain_bytes = ct.algorithmic_bytes_per_tick()
    in_bytes, out_bytes, n_streams, in_stride = ct.in_bytes, ct.out_bytes, ct.n_streams, ct.in_stride
    fused = ct.fused
    ct.close()   # frees the arenas before the next legs allocate their own
    del x

    total_sessions = S * world
    value = total_sessions * BUDGET_MS / ms_per_step
    # ---- the same workload with the inputs as s16 on PCIe (natively 16-bit sources: x = s / 32768): half the upload
    e2e_s16 = None
    if args.s16_extra and fused and not S16_IN and (IN_RATE // 50 + 32) * CHANNELS <= 4096:
        try:
            ct2 = chain.ChainTick(S, K, in_rate=IN_RATE, channels=CHANNELS, device=local_rank, seed=rank, s16=True)
            ct2.ctx.bind_thread()
            blk = np.rint(synth.noise_streams(2000 + rank, 0, 8192, ct2.chunk, CHANNELS) * 32767.0).astype(np.int16)
            ct2.fill_rows(ct2.host_in, np.tile(blk, ((ct2.n_streams + 8191) // 8192, 1))[: ct2.n_streams])
            d, _ = e2e_sliced(ct2, args, D, latency_ticks=min(args.latency_ticks, 100), d2h_gbs=(ceiling or {}).get("d2h_gbs_this_gpu"))
            d["value"] = total_sessions * TICK_MS / d["ms_per_step"]
            d["unit"] = UNIT
            d["h2d_bytes_per_step"], d["d2h_bytes_per_step"] = ct2.in_bytes * world, ct2.out_bytes * world
            d["what"] = "the same sessions with s16 input frames (skgpu_stream_cfg.flags = SKGPU_STREAM_S16): expanded to f32 in shared memory"
            e2e_s16 = d
            ct2.close()
        except Exception as e:
            e2e_s16 = {"error": str(e)[:200]}
    cap = None
    if args.capacity_check and fused:
        S2 = int(S * BUDGET_MS / ms_per_step * 0.97) // 1024 * 1024
        try:
            cap = capacity_check(S2, K, local_rank, args.kslices)
            cap["p99_ms"] = D.max(cap["p99_ms"])
        except Exception as e:  # an extra: never lose the main line
            cap = {"error": str(e)[:200], "sessions": S2}
    hub_e2e = None
    if args.hub and fused:
        try:
            hub_e2e = run_hub_e2e(S, K, max(5, min(args.steps, 20)), 2, local_rank, max(1, len(os.sched_getaffinity(0)) // max(world, 1)))
            if D.dist is not None:
                hub_e2e["ms_per_step"] = D.max(hub_e2e["ms_per_step"])
                hub_e2e["value"] = S * world * TICK_MS / hub_e2e["ms_per_step"]
                zc = hub_e2e.get("zero_copy")
                if zc:   # whole job, like every other figure of the line
                    zc["ms_per_step"] = D.max(zc["ms_per_step"])
                    zc["value"] = S * world * TICK_MS / zc["ms_per_step"]
        except Exception as e:
            hub_e2e = {"error": str(e)[:200]}

    e2e_value = total_sessions * TICK_MS / e2e["ms_per_step"]
    if ceiling is not None and "error" not in ceiling:
        per_session = in_bytes / S
        ceiling["sessions_per_gpu_at_that_upload_rate"] = ceiling["h2d_gbs_per_gpu_min"] * 1e9 / (per_session * (1e3 / TICK_MS))
        ceiling["e2e_fraction_of_ceiling"] = (e2e_value / world) / ceiling["sessions_per_gpu_at_that_upload_rate"]
    router_e2e = None
    if args.router and fused:
        # every rank is done with its own GPU: rank 0 alone now drives ALL the GPUs of the box from one process
        D.barrier()
        if rank == 0:
            try:
                router_e2e = run_router_e2e(S, K, world, max(5, min(args.steps, 20)))
            except Exception as e:
                router_e2e = {"error": str(e)[:300]}
    if rank == 0:
        peak, peak_src = hbm_peak()
        if fused:
            # dominant kernel = k_chain: the fully fused algorithmic bytes of SURVEY 8(d) / BASELINE.md:
            # per session-tick K x (7056 in + 272 state r/w) + 3840 s16 out = 18,496 B (K = 2)
            dom_name, dom_bytes = "k_chain<%d,1,%s>" % (CHANNELS, ("bypass" if IN_RATE == OUT_RATE else "plain") + ("_s16" if S16_IN else "")), chain_bytes
            std = S == 65536 and K == 2 and IN_RATE == 44100 and CHANNELS == 2 and not S16_IN
            traffic, traffic_src = ncu_traffic("k_chain<") if std else (None, "the committed capture is of 65,536 sessions x 2 stereo f32 44.1 kHz inputs")
        else:
            dom_name, dom_bytes = "k_resample_prog<2>", n_streams * (in_stride + 960 * CHANNELS * 4 + 2 * (8 + 16 * CHANNELS * 4))
            traffic, traffic_src = None, "no capture of the unfused path"
        achieved = dom_bytes / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
        cores = len(os.sched_getaffinity(0))
        cpu_sessions, cpu_ticks = args.ref_sessions, 50
        cpu_chain(cpu_sessions, K, 2, cores)
        cpu_sec, _ = cpu_chain(cpu_sessions, K, cpu_ticks, cores)
        cpu_value = cpu_sessions * TICK_MS / (cpu_sec * 1e3 / cpu_ticks)
        e2e_line = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes * world,
                    "what": "sessions sustained in real time (20 ms per tick) with every input uploaded from and every result read back to host memory"}
        e2e_line.update(e2e)
        e2e_line["pcie_ceiling"] = ceiling
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": chain_config(S, world, K),
            "value_definition": "sessions_per_gpu x n_gpus x 2 ms / ms_per_step: sessions whose whole tick (all kernels, inputs resident in HBM) "
                                "fits the 2 ms device budget; capacity_check runs that many for real. Real-time capacity at 20 ms per tick is 10x that.",
            "path": "fused k_phase_chain + k_chain" if fused else "unfused k_resample_prog + k_mix",
            "device_realtime_sessions_20ms": total_sessions * TICK_MS / ms_per_step,
            "samples_per_s": {"device": total_sessions * 960 * CHANNELS / (ms_per_step * 1e-3),
                              "e2e": (e2e_line or {}).get("value", 0.0) * 960 * CHANNELS / (TICK_MS * 1e-3),
                              "what": "output samples (48 kHz frames x channels) per second: device-resident kernels / end to end in real time"},
            "capacity_check": cap,
            "e2e": e2e_line,
            "e2e_s16_ingest": e2e_s16,
            "e2e_hub": hub_e2e,
            "e2e_router": router_e2e,
            "parity": parity,
            "gpu_launches": launches_per_tick * args.steps + (2 * e2e.get("slices", 0) + 1) * args.steps,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000_gbs": achieved / 8000.0,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "avg_launch_ms": main_ms, "launches_timed": n_main, "peak_source": peak_src},
            "kernels_ms": kernels_ms,
            "chain": {"algorithmic_bytes_per_tick": chain_bytes, "achieved_gbs": chain_bytes / (ms_per_step * 1e-3) / 1e9,
                      "frac_of_peak": chain_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                      "device_ms_per_tick": ms_per_step, "device_ms_per_tick_unsliced": unsliced_ms_per_step, "kernel_slices": args.kslices,
                      "what": "whole tick (every kernel of the tick), inputs resident; sliced = k_phase_chain(i + 1) overlaps k_chain(i)",
                      "tick_latency": tick_lat},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d sessions x %d ticks on %d host threads (oracle/sk_chain.c); sessions sustained in real time" % (cpu_sessions, cpu_ticks, cores)},
            "host": {"cpus": cores, "numa": numa},
            "device": {"name": dev_name, "sms": dev_sms, "cc": "%d.%d" % (dev_cc_ma, dev_cc_mi)},
        }
        D.emit(line)


def sms_of(ctx) -> int:
    return ctx.device_info()[1]


def run_node(args, D: Dist) -> None:
    """BASELINE configs[1..3]: the standalone node kernels, same JSON contract"""
    from oracle import sko
    from streamkit_b200 import lib as L, workloads as W

    world, rank, local_rank = D.world, D.rank, D.local_rank
    ctx = None
    if args.config == 2:
        ctx = L.Context(device=local_rank, max_streams=16, max_channels=2)
        w = W.GainS16(ctx, L.CVT_F32_TO_S16)
    elif args.config == 3:
        ctx = L.Context(device=local_rank, max_streams=16, max_channels=2)
        w = W.Mix64(ctx, s16=True)
    else:
        sinc = (64, 256, 0.95) if args.sinc else None
        w = W.Resample(48000, 16000, 16384, local_rank, sinc=sinc) if args.rs_down else W.Resample(44100, OUT_RATE, 16384, local_rank, sinc=sinc)
    plan, wctx = w.plan, w.ctx
    metric, unit, cpu_units = NODE_META[args.config]
    host_in = wctx.pinned(w.in_bytes, np.uint8)
    host_out = wctx.pinned(w.out_bytes, np.uint8)
    w.fill_host(host_in)
    plan.upload(0, host_in)
    dev_flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_TIME_OPS
    for _ in range(args.warmup):
        plan.submit(None, None, dev_flags)
    plan.wait()
    plan.reset_op_times()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    D.barrier()
    t_wall0 = time.time()
    wctx.timer_start()
    for _ in range(args.steps):
        plan.submit(None, None, dev_flags)
    wctx.timer_stop()
    ms_per_step = D.max(wctx.timer_ms()) / args.steps
    kernels_ms = {name: plan.op_time(op, sub)[0] for op, sub, name in w.ops}
    main_ms, n_main = plan.op_time(w.ops[0][0], w.ops[0][1])
    # end to end: host buffers, upload + kernels + read-back every step
    for _ in range(3):
        plan.submit(host_in, host_out, 0)
    plan.wait()
    D.barrier()
    wctx.timer_start()
    for _ in range(args.steps):
        plan.submit(host_in, host_out, L.SUBMIT_GRAPH | L.SUBMIT_OVERLAP_D2H)
    wctx.timer_stop()
    e2e_ms_per_step = D.max(wctx.timer_ms()) / args.steps
    timing = plan.wait()
    D.barrier()
    clk = clocks.stop(t_wall0, time.time())
    # parity of what the timed run produced (rank 0): every unit of a sample against the oracle
    parity = None
    if rank == 0:
        cores = len(os.sched_getaffinity(0))
        if args.config == 2:
            n = w.units_per_launch
            _s, want, _ = sko.node_bench(0, n, 1, w.pool, w.gains, cores, want_last=True)
            got = host_out.view(np.int16).reshape(n, 1920)
        elif args.config == 3:
            n = w.G
            _s, want, _ = sko.node_bench(1, n, 1, w.pool, w.master, cores, k_inputs=64, want_last=True)
            got = host_out.view(np.int16)[: n * 1920].reshape(n, 1920)
        else:
            n = w.S
            ticks = int(plan.tick_count())
            if args.sinc:   # the sinc oracle, stream class by stream class (256 distinct inputs tiled over the streams)
                D_ = w.base.shape[0]
                refs = [sko.SincFixedIn(w.in_rate, w.out_rate, w.chunk, 2, 64, 256, 0.95) for _ in range(D_)]
                last = None
                for _t in range(ticks):
                    last = [r.process(w.base[i]) for i, r in enumerate(refs)]
                cnt = np.array([last[i % D_].size // 2 for i in range(n)], dtype=np.uint32)
                want_f = np.zeros((n, w.cap * 2), np.float32)
                for i in range(n):
                    want_f[i, : last[i % D_].size] = last[i % D_]
            else:
                _s, want_f, cnt = sko.node_bench(2, n, ticks, w.base, None, cores, in_rate=w.in_rate, out_rate=w.out_rate, want_last=True, cap=w.cap)
            res = host_out[: 8 * n].view(L.RS_RESULT_DT)
            o0 = w.out_off - w.res_off
            outs = host_out[o0: o0 + n * w.out_stride].reshape(n, w.out_stride)[:, : w.cap * 8].copy().view(np.float32)
            m = int(cnt.min())
            got = np.concatenate([res["out_frames"].astype(np.uint32)[:, None], outs[:, : m * 2].view(np.uint32)], axis=1)
            want = np.concatenate([cnt[:, None], want_f[:, : m * 2].view(np.uint32)], axis=1)
        bad = int(np.count_nonzero(np.any(got != want, axis=1)))
        parity = {"units_checked": int(n), "units_differing": bad, "bit_exact": bad == 0,
                  "what": "every unit of the last end-to-end step vs the oracle (oracle/sk_chain.c sko_node_bench) on the same inputs"}
    units = w.units_per_launch if args.config != 2 else w.S * w.rot
    total_units = units * world
    value = total_units * BUDGET_MS / ms_per_step
    if rank == 0:
        peak, peak_src = hbm_peak()
        achieved = w.algorithmic_bytes / (main_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic({2: "k_convert<1", 3: "k_mix", 4: "k_resample_sinc_tiled<2" if args.sinc else "k_resample_prog<2"}[args.config]) if not args.rs_down else (None, "no capture of 48k->16k")
        cpu_node(args.config, cpu_units, 2, cores)
        cpu_iters = 20
        cpu_sec = cpu_node(args.config, cpu_units, cpu_iters, cores)
        cpu_value = cpu_units * TICK_MS / (cpu_sec * 1e3 / cpu_iters)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": node_config(args.config, world),
            "value_definition": "%s per launch (%d) x n_gpus x 2 ms / ms_per_step; the BASELINE size (%s) takes %.1f us" % (
                unit, units, {2: "4,096 sessions", 3: "1,024 groups", 4: "16,384 streams"}[args.config],
                ms_per_step * 1e3 * {2: 4096, 3: 1024, 4: 16384}[args.config] / units),
            "workload_detail": w.name,
            "e2e": {"value": total_units * TICK_MS / e2e_ms_per_step, "unit": unit, "h2d_bytes_per_step": w.in_bytes * world,
                    "d2h_bytes_per_step": w.out_bytes * world, "ms_per_step": e2e_ms_per_step,
                    "last_step_ms": {"h2d": timing.h2d_ms, "kernels": timing.kernels_ms, "d2h": timing.d2h_ms},
                    "what": "units sustained in real time (20 ms per tick) with host buffers: upload + kernels + read-back every step"},
            "parity": parity,
            "gpu_launches": plan.launches_per_tick() * args.steps * 2,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": w.ops[0][2], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": w.algorithmic_bytes, "avg_launch_ms": main_ms, "launches_timed": n_main,
                         "peak_source": peak_src},
            "kernels_ms": kernels_ms,
            "cpu_baseline": {"value": cpu_value, "unit": unit, "cores": cores, "kind": "port",
                             "sample": "%d %s x %d passes on %d host threads (oracle/sk_chain.c sko_node_bench); units sustained in real time" % (cpu_units, unit, cpu_iters, cores)},
        }
        if getattr(w, "fma_per_launch", None):
            # the sinc mode is bound by the FMA / LSU pipes (16 fma per byte moved), not by HBM: say how far from the FP32 peak it runs
            peak_fma = sms_of(wctx) * 128 * (clk.get("sm_mhz") or 1965.0) * 1e6
            line["roofline"]["fma"] = {"fma_per_launch": w.fma_per_launch, "achieved_tfma_s": w.fma_per_launch / (main_ms * 1e-3) / 1e12,
                                       "peak_tfma_s": peak_fma / 1e12, "frac": w.fma_per_launch / (main_ms * 1e-3) / peak_fma,
                                       "what": "f32 fused multiply-adds of the two tap-row dot products vs SMs x 128 lanes x SM clock"}
            line["cpu_baseline"]["note"] = "CPU arm of this line is the LINEAR resampler (the reference has no sinc mode)"
        D.emit(line)
    w.close()
    if ctx is not None:
        ctx.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5], help="BASELINE.json config number (1-based): 5 = full chain (default)")
    ap.add_argument("--sessions", type=int, default=65536, help="sessions per GPU (weak scaling)")
    ap.add_argument("--k", type=int, default=2, choices=[1, 2, 3, 4], help="inputs per session (SURVEY 8d: K = 1 and K = 2)")
    ap.add_argument("--in-rate", type=int, default=44100, help="input sample rate; 48000 = bypass inputs (Opus-decoder shaped, SURVEY 8f #3)")
    ap.add_argument("--channels", type=int, default=2, choices=[1, 2])
    ap.add_argument("--s16-in", action="store_true", help="inputs arrive as s16 on PCIe (x = s / 32768), half the upload")
    ap.add_argument("--slices", type=int, default=0, help="slices per end-to-end tick; 0 = automatic (>= 16, more when the box's device->host rate is low)")
    ap.add_argument("--kslices", type=int, default=8, help="slices of the device-resident tick (phase / chain kernel overlap); 1 = whole-tick launches")
    ap.add_argument("--unfused", action="store_true", help="use the general unfused ops (k_resample_prog -> ring -> k_mix)")
    ap.add_argument("--rs-down", action="store_true", help="--config 4: 48k->16k instead of 44.1k->48k")
    ap.add_argument("--sinc", action="store_true", help="--config 4: the windowed-sinc polyphase mode (sinc_len 64, oversampling 256) instead of rubato's Linear")
    ap.add_argument("--ref-sessions", type=int, default=8192, help="bounded session sample of the CPU arm")
    ap.add_argument("--parity-sessions", type=int, default=256, help="sessions of the timed run compared with the CPU chain (0 = skip)")
    ap.add_argument("--no-hub", dest="hub", action="store_false", help="skip the frame-batching-layer end-to-end measurement")
    ap.add_argument("--no-router", dest="router", action="store_false", help="skip the single-process multi-GPU router measurement")
    ap.add_argument("--no-s16-extra", dest="s16_extra", action="store_false", help="skip the s16-ingest end-to-end variant")
    ap.add_argument("--no-capacity-check", dest="capacity_check", action="store_false")
    ap.add_argument("--latency-ticks", type=int, default=500, help="ticks of the per-slice latency measurement (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    global IN_RATE, CHANNELS, S16_IN
    IN_RATE, CHANNELS, S16_IN = args.in_rate, args.channels, args.s16_in
    if args.impl == "reference":
        run_reference(args)
        return
    D = Dist()
    try:
        if args.config == 5:
            run_chain(args, D)
        else:
            run_node(args, D)
    finally:
        D.close()


if __name__ == "__main__":
    main()
