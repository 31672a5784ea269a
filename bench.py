#!/usr/bin/env python
"""bench.py -- headline benchmark of streamkit_b200 (driver contract in the task statement).

Workload (BASELINE.json configs[4], the configuration the metric is quoted on; fits one GPU):
    full chain resample 44.1k->48k -> per-input gain -> ordered mix -> master gain -> clip -> s16,
    K = 2 stereo f32 inputs per session, SESSIONS_PER_GPU sessions per GPU, one 20 ms tick per step.
Metric: concurrent real-time 48 kHz stereo sessions = session-ticks processed per second / 50.

  value : device-resident (inputs already in HBM when the timed region starts), CUDA events, max over ranks
  e2e   : the same tick submitted through the C ABI with HOST (pinned) buffers: H2D of every input frame and
          D2H of every s16 result inside the timed region
  roofline     : dominant kernel (k_chain; k_resample with --unfused), algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline : the oracle's reference-shaped CPU chain (oracle/sk_chain.c) on the box's host cores

`--impl reference` times that CPU chain as the reference arm (the reference is Rust and cannot be built in this
image: DESIGN.md "Oracle"; the C restatement is the port).
"""
from __future__ import annotations

import argparse
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

IN_RATE, OUT_RATE, CHANNELS, K_INPUTS = 44100, 48000, 2, 2
TICK_MS = 20.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
# (profiles/r1_ncu_full_summary.csv: 1.100923 GB read + 252.77 MB written at 65,536 sessions x 2 inputs), or None
TRAFFIC_NCU: dict = {"k_chain<2>": 1100923000 + 252765952}
METRIC = "concurrent real-time 48 kHz stereo sessions (resample->mix->gain->s16, 20 ms ticks)"
UNIT = "sessions"


def workload_config(sessions_per_gpu: int, n_gpus: int) -> dict:
    return {
        "workload": "BASELINE configs[4]: full chain resample 44.1k->48k -> gain -> %d-input ordered mix -> gain -> clip -> s16"
                    % K_INPUTS,
        "sessions_per_gpu": sessions_per_gpu,
        "inputs_per_session": K_INPUTS,
        "in_rate": IN_RATE, "out_rate": OUT_RATE, "channels": CHANNELS, "tick_ms": TICK_MS,
        "chunk_frames": IN_RATE // 50, "output_frame_size": 960,
        "sharding": "sessions split evenly across %d GPU(s) by session id, no collective" % n_gpus,
        "l2": "inputs are %.0f MB per tick per GPU (> 126 MB L2): no flush needed" %
              (sessions_per_gpu * K_INPUTS * (IN_RATE // 50) * CHANNELS * 4 / 1e6),
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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ---------------------------------------------------------------------------------------------- CPU arm

def cpu_chain(n_sessions: int, ticks: int, threads: int, seed: int = 0):
    """the oracle's reference-shaped chain on host cores; returns seconds for `ticks` ticks of n_sessions"""
    from oracle import sko
    from streamkit_b200 import synth

    pool = synth.noise_streams(seed, 0, 256, IN_RATE // 50, CHANNELS)
    ig = synth.gains(seed, n_sessions * K_INPUTS, 0.25, 1.5)
    mg = synth.gains(seed + 1, n_sessions, 0.5, 2.0)
    sec, _cs, _ = sko.chain_bench(n_sessions, K_INPUTS, ticks, IN_RATE, CHANNELS, pool, ig, mg, threads)
    return sec


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_sessions = args.ref_sessions
    # one "step" = one 20 ms tick of the bounded session sample; state carries across steps inside one call
    cpu_chain(n_sessions, max(args.warmup, 1), cores)
    sec = cpu_chain(n_sessions, args.steps, cores)
    ms_per_step = sec * 1e3 / args.steps
    value = n_sessions * TICK_MS / ms_per_step
    sample = "%d sessions x %d ticks on %d host threads (oracle/sk_chain.c, reference-shaped per-packet nodes)" % (
        n_sessions, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.sessions, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust (no toolchain in this image): timed arm is the C restatement of its nodes; tokio scheduling"
                " and channel hops of the real engine are not included (optimistic for the reference)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- frame-batching layer

def run_hub_e2e(S: int, steps: int, warmup: int, device: int, threads: int) -> dict:
    """The same workload driven through the frame-batching layer (include/skgpu_hub.h): per tick the chunks of all
    S x K streams are gathered from separate host buffers into the hub's pinned arena by `threads` worker threads
    (skgpu_hub_push_batch), then one asynchronous tick; the gather of tick n + 1 overlaps tick n on the GPU.
    Wall-clock time (it includes host work), results read back every tick."""
    from streamkit_b200 import hub as H, synth

    chunk = IN_RATE // 50
    hub = H.Hub(max_sessions=S, max_streams=S * K_INPUTS, in_rates=[IN_RATE], max_inputs_per_session=K_INPUTS, channels=CHANNELS, device=device)
    try:
        pool = np.ascontiguousarray(synth.noise_streams(4242, 0, 256, chunk, CHANNELS))   # 256 distinct chunks, reused round-robin
        frames = np.zeros(S * K_INPUTS, dtype=H.FRAME_DT)
        for s in range(S):
            sid = hub.session_open([IN_RATE] * K_INPUTS)
            for i in range(K_INPUTS):
                f = frames[s * K_INPUTS + i]
                f["session"], f["input"], f["n_frames"] = sid, i, chunk
        frames["samples"] = pool.ctypes.data + (np.arange(frames.size, dtype=np.uint64) % 256) * np.uint64(pool.strides[0])
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
                "what": "skgpu_hub: multi-threaded gather into pinned arena + H2D + kernels + D2H per tick, wall clock",
                "check": {"n_mixed": int(n_mixed), "status": int(status), "nonzero": bool(out is not None and np.any(out != 0))}}
    finally:
        hub.close()


# ---------------------------------------------------------------------------------------------- GPU arm

def run_gpu(args) -> None:
    import torch

    from streamkit_b200 import chain, lib as L, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; streamkit_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of the timings (no data-path collective)

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S = args.sessions
    ct = chain.ChainTick(S, K_INPUTS, in_rate=IN_RATE, channels=CHANNELS, device=local_rank, seed=rank, fused=not args.unfused)
    plan, ctx = ct.plan, ct.ctx
    # synthetic input: a tick of noise for every stream of every session on this rank
    x = synth.noise_streams(1000 + rank, 0, ct.n_streams, ct.chunk, CHANNELS)
    ct.host_in[:] = x.reshape(-1)
    del x
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
    barrier()
    t_wall0 = time.time()
    ctx.timer_start()
    for _ in range(args.steps):
        plan.submit(None, None, dev_flags)
    ctx.timer_stop()
    dev_ms = ctx.timer_ms()
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / args.steps
    if ct.fused:
        phase_ms, _ = plan.op_time(ct.op_chain, 0)
        main_ms, n_main = plan.op_time(ct.op_chain, 1)
        kernels_ms = {"k_phase": phase_ms, "k_chain": main_ms}
    else:
        phase_ms, _ = plan.op_time(ct.op_rs, 0)
        main_ms, n_main = plan.op_time(ct.op_rs, 1)
        mix_ms, _ = plan.op_time(ct.op_mix, 0)
        kernels_ms = {"k_phase": phase_ms, "k_resample": main_ms, "k_mix+k_fifo_commit": mix_ms}

    # ---- per-tick device latency (added latency per 20 ms tick, p99 over >= 500 ticks): one tick at a time,
    # kernels only (upload-done -> results ready for the read-back), CUDA events inside the library
    lat = []
    for _ in range(args.latency_ticks):
        plan.submit(None, None, L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H)
        lat.append(plan.wait().kernels_ms)
    lat.sort()
    latency = {"ticks": len(lat), "p50_ms": lat[len(lat) // 2], "p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))], "max_ms": lat[-1],
               "what": "device time of one tick's kernels at sessions_per_gpu sessions, submitted and awaited tick by tick"} if lat else None

    # ---- end-to-end region: every step copies its inputs from pinned host memory and reads the s16 result back
    for _ in range(max(1, min(args.warmup, 3))):
        plan.submit(ct.host_in, ct.host_out, 0)
    plan.wait()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        plan.submit(ct.host_in, ct.host_out, L.SUBMIT_GRAPH | L.SUBMIT_OVERLAP_D2H)
    ctx.timer_stop()
    e2e_ms = ctx.timer_ms()
    timing = plan.wait()
    barrier()
    t_wall1 = time.time()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_ms_per_step = e2e_ms / args.steps
    clk = clocks.stop(t_wall0, t_wall1)

    launches_per_tick = plan.launches_per_tick()
    dev_name, dev_sms, dev_cc_ma, dev_cc_mi = ctx.device_info()
    ct.close()   # frees the arenas before the frame-batching layer allocates its own
    hub_e2e = None
    if args.hub:
        try:
            hub_e2e = run_hub_e2e(S, max(5, min(args.steps, 20)), 2, local_rank, max(1, len(os.sched_getaffinity(0)) // max(world, 1)))
            if dist is not None:
                hub_e2e["ms_per_step"] = max_over_ranks(hub_e2e["ms_per_step"])
                hub_e2e["value"] = S * world * TICK_MS / hub_e2e["ms_per_step"]
        except Exception as e:  # the layer is an extra: never lose the main line
            hub_e2e = {"error": str(e)[:200]}

    total_sessions = S * world
    value = total_sessions * TICK_MS / ms_per_step
    e2e_value = total_sessions * TICK_MS / e2e_ms_per_step

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        chain_bytes = ct.algorithmic_bytes_per_tick()
        if ct.fused:
            # dominant kernel = k_chain: the fully fused algorithmic bytes of SURVEY 8(d) / BASELINE.md:
            # per session-tick K x (7056 in + 272 state r/w) + 3840 s16 out = 18,496 B (K = 2)
            dom_name, dom_bytes = "k_chain<2>", chain_bytes
        else:
            # dominant kernel = k_resample: in + out + state r/w per stream-tick (15,008 B for 44.1k->48k stereo)
            dom_name, dom_bytes = "k_resample<2>", ct.n_streams * (ct.in_stride + 960 * CHANNELS * 4 + 2 * (8 + 16 * CHANNELS * 4))
        achieved = dom_bytes / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
        cores = len(os.sched_getaffinity(0))
        cpu_sessions, cpu_ticks = args.ref_sessions, 50
        cpu_chain(cpu_sessions, 2, cores)
        cpu_sec = cpu_chain(cpu_sessions, cpu_ticks, cores)
        cpu_value = cpu_sessions * TICK_MS / (cpu_sec * 1e3 / cpu_ticks)
        name, sms, cc_ma, cc_mi = dev_name, dev_sms, dev_cc_ma, dev_cc_mi
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(workload_config(S, world), path="fused k_chain" if ct.fused else "unfused k_resample + k_mix"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ct.in_bytes * world, "d2h_bytes_per_step": ct.out_bytes * world,
                    "ms_per_step": e2e_ms_per_step, "last_tick_ms": {"h2d": timing.h2d_ms, "kernels": timing.kernels_ms, "d2h": timing.d2h_ms},
                    "h2d_gbs_per_gpu": ct.in_bytes / (timing.h2d_ms * 1e-3) / 1e9 if timing.h2d_ms > 0 else None,
                    "d2h_gbs_per_gpu": ct.out_bytes / (timing.d2h_ms * 1e-3) / 1e9 if timing.d2h_ms > 0 else None},
            "e2e_hub": hub_e2e,
            "gpu_launches": launches_per_tick * args.steps * 2,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000_gbs": achieved / 8000.0,
                         "traffic": TRAFFIC_NCU.get(dom_name) if (S == 65536 and K_INPUTS == 2) else None,
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "avg_launch_ms": main_ms, "launches_timed": n_main, "peak_source": peak_src},
            "kernels_ms": kernels_ms,
            "chain": {"algorithmic_bytes_per_tick": chain_bytes, "achieved_gbs": chain_bytes / (ms_per_step * 1e-3) / 1e9,
                      "frac_of_peak": chain_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                      "device_ms_per_tick": ms_per_step, "latency_budget_ms": 2.0, "latency": latency},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d sessions x %d ticks on %d host threads (oracle/sk_chain.c)" % (cpu_sessions, cpu_ticks, cores)},
            "device": {"name": name, "sms": sms, "cc": "%d.%d" % (cc_ma, cc_mi)},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sessions", type=int, default=65536, help="sessions per GPU (weak scaling)")
    ap.add_argument("--unfused", action="store_true", help="use the general unfused ops (k_resample -> ring -> k_mix)")
    ap.add_argument("--ref-sessions", type=int, default=8192, help="bounded session sample of the CPU arm")
    ap.add_argument("--no-hub", dest="hub", action="store_false", help="skip the frame-batching-layer end-to-end measurement")
    ap.add_argument("--latency-ticks", type=int, default=500, help="ticks of the per-tick latency measurement (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
