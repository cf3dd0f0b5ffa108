#!/usr/bin/env python
"""bench.py — radar scans registered per second (Oxford shape, 400 x 3768 polar) on N B200s vs the reference CPU path.

Workload (BASELINE.json configs[1]): CFEAR scan-to-4-keyframes P2L registration on a synthetic Oxford-shape stream —
k-strongest filter (k=40, z_min=60) -> motion compensation -> oriented surface points (r=3, intensity weights) ->
registration of the scan against up to 4 keyframes (P2L, Huber 0.1, combined weights, <=8 association x <=20 LM
iterations) -> keyframe policy.  One "step" advances `--seqs` independent sequences PER GPU by one frame each (frame t of
a sequence needs the pose of frame t-1, so the batch is across sequences — the reference's own scaling model is one
worker process per sequence).  Multi-GPU: sequences are sharded over ranks, no data-path collective (weak scaling).

  value : scans/s with every step's scans already resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e   : scans/s through the reference-facing call with HOST buffers: pinned host scans -> H2D -> pipeline -> D2H of
          one pose record per sequence, double-buffered (tbv_odom_submit / tbv_odom_collect), wall clock between syncs
  --impl reference : the CPU oracle (the reference cannot be built here) on all host threads, same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AZ, N_RANGE, K_STRONGEST = 400, 3768, 40
SCAN_BYTES = N_AZ * N_RANGE
METRIC = "radar scans registered/sec (400x3768 polar)"
UNIT = "scans/s"
WORKLOAD = "CFEAR scan-to-4-keyframes P2L registration, synthetic Oxford-shape stream (configs[1])"
POOL_EXTRA = 16  # distinct starting offsets into the synthetic stream


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs", type=int, default=592, help="independent sequences per GPU advanced in lock-step (592 = 148 SMs x 4)")
    ap.add_argument("--cpu-seqs", type=int, default=0, help="sequences in the cpu_baseline sample (0: sized for ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


N_PLACES = 8  # distinct stretches of the figure-8 the sequences start from


def make_pool(n_frames: int, rank: int):
    """[N_PLACES * n_frames] scans: N_PLACES stretches of the synthetic world, n_frames consecutive frames each.  Every rank renders
    the same stretches; what differs per rank is which sequence drives which stretch (first_offsets), so every GPU gets the same mix
    of scene densities — weak scaling measures the machine, not the luck of one rank's neighbourhood."""
    from tbv_slam_public_b200 import synth
    return np.concatenate([synth.make_stream(n_frames, s0=137.0 * p).scans for p in range(N_PLACES)])


def first_offsets(n_seq: int, n_frames: int, rank: int = 0):
    """Index of every sequence's first scan in the pool: stretch (j + rank) mod N_PLACES, starting frame (5 j) mod POOL_EXTRA."""
    j = np.arange(n_seq)
    return (((j + rank) % N_PLACES) * n_frames + (j * 5) % POOL_EXTRA).astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (the profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.1] or [ln for (_, ln) in self.lines[-3:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def pick_hbm_peak(peaks):
    """(GB/s, key path) of the HBM figure in MEASURED_PEAKS.json, whatever its exact layout: the file is written by the driver and is not in
    the build container, so keys are matched by name — an HBM / copy-bandwidth number in GB/s (or TB/s), the SUSTAINED one when both a burst
    and a sustained figure are given (K1 is timed inside a long step).  (6650.0, None) = the profiling recipe's fallback."""
    flat = []

    def walk(o, path):
        if isinstance(o, dict):
            for k, v in o.items():
                walk(v, path + [str(k)])
        elif isinstance(o, (int, float)) and not isinstance(o, bool):
            flat.append((".".join(path), float(o)))

    walk(peaks, [])
    cand = []
    for key, v in flat:
        k = key.lower()
        if not any(t in k for t in ("hbm", "copy", "bandwidth", "dram", "mem_bw", "membw")) or any(t in k for t in ("tflop", "tf_s", "tfs", "bf16", "fp8", "l2")):
            continue
        gbs = v * 1000.0 if 1.0 <= v <= 20.0 else v              # TB/s -> GB/s
        if 1000.0 <= gbs <= 10000.0:
            cand.append((0 if "sustain" in k else (2 if "burst" in k else 1), key, gbs))
    if not cand:
        return 6650.0, None
    cand.sort()
    return cand[0][2], cand[0][1]


def algorithmic_bytes(kernel: str, n_seq: int, st: dict) -> float | None:
    """ALGORITHMIC bytes one launch of `kernel` must move for n_seq scans (DESIGN.md 'Kernels'); None if not HBM-shaped."""
    npts, nsmp, ncell, nkf = st["n_points"], st["n_samples"], st["n_cells"], st["n_keyframes"]
    per = {
        # fused filter kernel: scan bytes in, the two clouds out (x,y f32, I u8, az,rg u16 = 13 B/pt; peaks ~ a third of the points)
        "k1_filter_fused": SCAN_BYTES + 13 * npts * 1.33,
        "k_compensate": 2 * 8 * npts,
        # fused cells kernel: points (x,y f32 + I u8) in, one 16-double record per valid cell out (all tables in shared memory)
        "cells_fused": 9 * npts + 128 * ncell,
        # legacy multi-kernel cells path (fine voxel grids only)
        "c5_cells": 9 * npts + 128 * nsmp,
        "c4_centroids": 12 * npts + 8 * nsmp,
        # (moving + keyframes) cells x 6 doubles used (u, n, N, planarity) in, one result record out
        "k_register": (1 + nkf) * ncell * 48 + 136,
    }
    return per[kernel] * n_seq if kernel in per else None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tbv_slam_public_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K, S = args.warmup, args.steps, args.seqs
    T = W + K
    pool = make_pool(T + POOL_EXTRA, rank)
    first = first_offsets(S, T + POOL_EXTRA, rank)

    ctx = api.Context(local)
    par = api.default_odom_params()
    fuser = api.OdometryKeyframeFuser(ctx, S, N_AZ, N_RANGE, par)
    stream = torch.cuda.ExternalStream(ctx.stream)

    # host side: the scans of the steps the e2e leg replays, in pinned memory (what a sensor-facing producer would hand over).
    # All ranks of the node pin memory at once: keep the node total under 40 % of what is available (T_pin steps per rank, the
    # e2e leg then times T_pin - W steps; on the single-GPU box that is every step).
    step_bytes = S * SCAN_BYTES
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    T_pin = int(max(W + 3, min(T, 0.4 * avail / local_world // step_bytes)))
    T_pin = min(T_pin, T)
    if world > 1:  # the same number of e2e steps on every rank
        tp = torch.tensor([T_pin], dtype=torch.int64, device="cuda")
        dist.all_reduce(tp, op=dist.ReduceOp.MIN)
        T_pin = int(tp[0])
    pinned = api.PinnedBuffer(T_pin * step_bytes)
    host = pinned.array.reshape(T_pin, S, N_AZ, N_RANGE)
    # device side: every step's scans resident in HBM (each step reads S x 1.5 MB, far larger than the 126 MB L2)
    dev = torch.empty((T, step_bytes), dtype=torch.uint8, device="cuda")
    for t in range(T):   # stage through the pinned ring
        slot = t % T_pin
        np.take(pool, first + t, axis=0, out=host[slot])
        dev[t].copy_(torch.from_numpy(pinned.array.reshape(T_pin, step_bytes)[slot]), non_blocking=False)
    for t in range(T_pin):  # the pinned buffer ends up holding steps 0 .. T_pin-1
        if T_pin < T:
            np.take(pool, first + t, axis=0, out=host[t])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- value: inputs resident in HBM ----------------------------------------------------------
    for t in range(W):
        fuser.step_dev(dev[t].data_ptr())
    ctx.synchronize(); torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for t in range(W, T):
        fuser.step_dev(dev[t].data_ptr())
    e1.record(stream)
    e1.synchronize(); ctx.synchronize(); torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    launches = ctx.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    outs = fuser.fetch()
    final_dev = api.poses(outs).copy()
    stats = {"n_points": float(np.mean([o.n_points for o in outs])), "n_cells": float(np.mean([o.n_cells for o in outs])),
             "n_samples": float(np.mean([o.n_samples for o in outs])),
             "n_keyframes": float(np.mean([o.n_keyframes for o in outs])), "itrs": float(np.mean([o.itrs for o in outs])),
             "lm_iterations": float(np.mean([o.lm_iterations for o in outs])), "num_residuals": float(np.mean([o.num_residuals for o in outs])),
             "reg_ok": float(np.mean([o.reg_ok for o in outs])), "status_max": int(max(abs(o.status) for o in outs))}
    if stats["status_max"] != 0:
        raise RuntimeError("a per-scan capacity was exceeded inside the timed region: results invalid")

    # ---------------- per-kernel device time (events after every launch; separate pass, not part of `value`) ----
    fuser.reset()
    for t in range(W):
        fuser.step_dev(dev[t].data_ptr())
    ctx.synchronize()
    n_prof = min(3, K)
    ctx.profile_begin()
    for t in range(W, W + n_prof):
        fuser.step_dev(dev[t].data_ptr())
    kern = {}
    for name, ms in ctx.profile_end():
        kern[name] = kern.get(name, 0.0) + ms / n_prof

    # ---------------- e2e: host buffers, H2D + D2H inside the timed region, double-buffered --------------------
    fuser.reset()
    for t in range(W):
        fuser.pointcloudCallback(host[t])
    ctx.synchronize(); torch.cuda.synchronize()
    barrier()
    K_e2e = T_pin - W
    t0 = time.perf_counter()
    for t in range(W, T_pin):
        fuser.submit(pinned.ptr + t * step_bytes)
        if t > W:
            fuser.collect()
    outs_e2e = fuser.collect()
    ctx.synchronize()
    t1 = time.perf_counter()
    barrier()
    e2e_s = t1 - t0
    final_e2e = api.poses(outs_e2e).copy()
    if T_pin == T:
        same = np.array_equal(final_dev, final_e2e)
    else:  # fewer e2e steps than device-resident ones: replay those steps from HBM and compare
        fuser.reset()
        for t in range(T_pin):
            fuser.step_dev(dev[t].data_ptr())
        same = np.array_equal(api.poses(fuser.fetch()), final_e2e)
    if not same:
        raise RuntimeError("device-resident and host-buffer runs disagree")

    # ---------------- max over ranks ---------------------------------------------------------------------------
    if world > 1:
        tt = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(tt[0]), float(tt[1])
    total_scans = S * K * world
    value = total_scans / (ms_total * 1e-3)
    e2e_value = S * K_e2e * world / e2e_s

    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---------------- cpu baseline (rank 0, N=1 only): the oracle, single thread, bounded sample ----------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_py
        n_cpu = args.cpu_seqs or max(8, min(S, int(10.0 / (0.0035 * T))))
        sec, ref_poses = oracle_py.odom_run_timed(oracle_py.default_odom_params(), pool, first[:n_cpu], W, T, 1)
        cpu_baseline = {"value": n_cpu * K / sec, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"{n_cpu} of the {S} sequences x {K} timed frames (after {W} warm-up frames), oracle/ C++ port, 1 thread"}
        d = final_dev[:n_cpu] - ref_poses[:, T - 1, :]
        parity = {"sequences": n_cpu, "frames": T, "max_abs_xy_m": float(np.abs(d[:, :2]).max()),
                  "max_abs_yaw_rad": float(np.abs(np.arctan2(np.sin(d[:, 2]), np.cos(d[:, 2]))).max())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm, peak_key = pick_hbm_peak(peaks)
    kernels = {}
    for name, ms in sorted(kern.items(), key=lambda kv: -kv[1]):
        b = algorithmic_bytes(name, S, stats)
        kernels[name] = {"ms_per_step": round(ms, 4), "share": round(ms / sum(kern.values()), 4),
                         "algorithmic_GBps": round(b / ms / 1e6, 1) if b else None}
    # The roofline object describes the kernel that dominates the step's HBM traffic: K1 streams every scan byte (96 % of the
    # step's algorithmic bytes).  The longest kernels (k_register, cells_fused) move ~100x fewer bytes and are bound by fp64
    # issue / dependent-load latency / barriers, not by HBM or the tensor pipe: their lines are in `kernels`, the whole step's in `step`.
    dominant = "k1_filter_fused"
    longest = max(kern, key=kern.get)
    b_dom = algorithmic_bytes(dominant, S, stats)
    achieved = b_dom / kern[dominant] / 1e6
    step_bytes_alg = sum(algorithmic_bytes(k, S, stats) or 0.0 for k in kern)
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": round(achieved, 2), "peak": peak_hbm, "unit": "GB/s",
                "frac": round(achieved / peak_hbm, 5),
                # dram__bytes_read.sum + dram__bytes_write.sum of one k1_kstrongest launch over 592 scans (ncu --set full,
                # profiles/r1h_full_k1_kstrongest.txt): 898.1 MB + 17.6 MB -> 1.5467 MB per scan; the algorithmic figure (1.572 MB)
                # is slightly larger because part of the 64.8 kB of row keys per scan is still in L2 when K2 consumes it
                "traffic": round(1.5467e6 * S, 0), "traffic_source": "profiles/r1h_full_k1_kstrongest.txt (592 scans per launch), scaled by sequences_per_gpu / 592",
                "algorithmic_bytes_per_launch": b_dom, "launch_ms": round(kern[dominant], 4),
                "peak_source": f"MEASURED_PEAKS.json {peak_key} (of measured)" if peak_key else "fallback 6650 GB/s (of fallback)",
                "share_of_step": kernels[dominant]["share"], "longest_kernel": longest, "longest_kernel_share": kernels[longest]["share"],
                "step": {"algorithmic_bytes": step_bytes_alg, "achieved": round(step_bytes_alg / (ms_total / K) / 1e6, 2),
                         "frac": round(step_bytes_alg / (ms_total / K) / 1e6 / peak_hbm, 5)},
                "note": "launch duration from CUDA events recorded after every launch on the library's stream (tbv_profile_begin/_end)"}
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 scan bytes -> f32 points -> f64 cells / normal equations / LM", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sequences_per_gpu": S, "scans_per_step": S * world, "n_az": N_AZ, "n_range": N_RANGE,
                   "k_strongest": K_STRONGEST, "z_min": 60, "cell_radius_m": 3.0, "keyframes": 4, "cost": "P2L", "loss": "Huber(0.1)",
                   "weights": "combined", "sharding": f"sequences over {world} GPU(s), no collective", "places": N_PLACES,
                   "l2": "each step reads sequences_per_gpu x 1.5 MB of scans (>> 126 MB L2); no flush needed"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": step_bytes, "d2h_bytes_per_step": S * 80,
                "ms_per_step": round(e2e_s / K_e2e * 1e3, 4), "steps": K_e2e, "api": "tbv_odom_submit/tbv_odom_collect (pinned host scans, double-buffered)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "cpu_baseline": cpu_baseline, "parity_check": parity, "workload_stats": {k: round(v, 3) if isinstance(v, float) else v for k, v in stats.items()},
    }
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU path for the same workload: oracle/ (C++ port; the reference itself needs ROS/PCL/Ceres and cannot
    be built here), one worker per host thread over independent sequences — the reference's own scaling model."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py
    W, K = args.warmup, args.steps
    T = W + K
    threads = os.cpu_count() or oracle_py.hardware_threads() or 1
    n_seq = threads * max(1, int(round(20.0 / (0.0035 * T))))  # ~20 s of work per thread
    n_seq = min(n_seq, threads * 64)
    pool = make_pool(T + POOL_EXTRA, 0)
    first = first_offsets(n_seq, T + POOL_EXTRA, 0)
    sec, _ = oracle_py.odom_run_timed(oracle_py.default_odom_params(), pool, first, W, T, threads)
    value = n_seq * K / sec
    # Filtering stage alone, one thread: the reference's OWN radar_filters.cpp (oracle/_ref, built from /root/reference where that
    # exists) next to the oracle port of the same stage — shows which way the port's omitted overheads bias the baseline.
    filtering = None
    try:
        from oracle import ref_py
        if ref_py.available():
            n_f = min(48, len(pool))
            t0 = time.perf_counter(); ref_py.kstrongest_many(pool[:n_f], 60.0, K_STRONGEST, 2.5, 0.0438); t_ref = time.perf_counter() - t0
            t0 = time.perf_counter()
            for i in range(n_f):
                oracle_py.kstrongest(pool[i], z_min=60.0, k=K_STRONGEST)
            t_port = time.perf_counter() - t0
            filtering = {"reference_ms_per_scan": round(t_ref / n_f * 1e3, 3), "port_ms_per_scan": round(t_port / n_f * 1e3, 3), "scans": n_f,
                         "threads": 1, "note": "reference = unmodified radar_filters.cpp (constructor + both clouds) from oracle/_ref"}
    except Exception as e:  # the timing of one stage must never take the arm down
        filtering = {"error": str(e)[:200]}
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": round(sec / K * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 scan bytes -> f32 points -> f64 cells / normal equations / LM", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sequences": n_seq, "n_az": N_AZ, "n_range": N_RANGE, "k_strongest": K_STRONGEST, "z_min": 60,
                   "cell_radius_m": 3.0, "keyframes": 4, "cost": "P2L", "loss": "Huber(0.1)", "weights": "combined"},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_seq} independent sequences x {K} timed frames (after {W} warm-up frames) over {threads} host threads; "
                                   "a step = one frame of every sequence"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "filtering_stage": filtering,
    }))


_REAL_STDOUT = None


def emit(line: str):
    """The bench prints exactly ONE line on stdout.  Libraries write there too (NCCL's version banner, for one), so the process's
    fd 1 is pointed at stderr for the whole run and the result line goes to the saved original stdout."""
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line + "\n"); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
