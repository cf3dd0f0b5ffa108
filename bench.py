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

Two more legs run in every invocation and are reported as sub-objects of the same JSON line:
  loop_batch : batched loop-closure candidate registration (loopclosure.cpp:658-724), 1 024 and 256 candidates PER GPU per iteration over a
               replicated keyframe database, candidates sharded id_from mod N, accepted constraints all-gathered inside the library
               (tbv_loopdb_register_sharded: ncclAllGather + device merge); pairs/s, all-gather microseconds, CPU baseline, parity.
  mulran     : MulRan-shape odometry (400 x 3360, 0.0595238 m, ccw, range-major wire layout rotated on receipt on the device) alone and
               mixed with one 1 000-candidate-per-GPU loop batch per step on a second context; scans/s, pairs/s, K1 roofline for the shape.
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
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the loop_batch and mulran legs (profiling runs)")
    ap.add_argument("--no-overlap", action="store_true", help="single-stream steps with the compensation fused into the filter kernel (tbv_odom_set_overlap off)")
    return ap.parse_args()


N_PLACES = 8  # distinct stretches of the figure-8 the sequences start from
LOOP_PAIRS_PER_GPU = (1024, 256)   # loop_batch leg: candidates per GPU per iteration
LOOP_KF_PER_PLACE = 12             # keyframes of the loop database: the first frames of every stretch (2.5 m apart: overlapping views)
MU_N_RANGE, MU_RANGE_RES = 3360, 0.0595238
MU_PLACES, MU_SEQS, MU_WARMUP, MU_STEPS, MU_LOOP_PAIRS = 2, 592, 3, 8, 1000
SURVEY_K1_UNIT = 1_715_200         # K1 per OX scan, k = 40: 1 507 200 B read + 208 000 B written with every row full
K1_TRAFFIC_FILE = "profiles/r2g_full_k1_filter_fused.txt"   # ncu --set full summary of the kernel as it runs in this bench (tools/gpu_final.sh)


def make_pool(n_frames: int, rank: int, dataset=None, places: int = N_PLACES, with_gt: bool = False):
    """[places * n_frames] scans: `places` stretches of the synthetic world, n_frames consecutive frames each.  Every rank renders
    the same stretches; what differs per rank is which sequence drives which stretch (first_offsets), so every GPU gets the same mix
    of scene densities — weak scaling measures the machine, not the luck of one rank's neighbourhood."""
    from tbv_slam_public_b200 import synth
    streams = [synth.make_stream(n_frames, s0=137.0 * p, **({"dataset": dataset} if dataset else {})) for p in range(places)]
    scans = np.concatenate([st.scans for st in streams])
    return (scans, np.concatenate([st.gt for st in streams])) if with_gt else scans


def bench_config() -> dict:
    """The `config` object: the SAME keys and values in both arms (what differs between the arms — how many sequences a run drives — is
    reported outside it, under `run`)."""
    return {"workload": WORKLOAD, "n_az": N_AZ, "n_range": N_RANGE, "k_strongest": K_STRONGEST, "z_min": 60, "cell_radius_m": 3.0, "keyframes": 4,
            "cost": "P2L", "loss": "Huber(0.1)", "weights": "combined", "places": N_PLACES,
            "sharding": "sequences over the GPUs (no collective on the odometry path); loop candidates by id_from mod N with an all-gather of the accepted constraints",
            "l2": "each step reads sequences x 1.5 MB of scans (>> 126 MB L2); no flush needed"}


def k1_traffic_per_scan(path: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per scan from the committed ncu summary of k1_filter_fused (one CTA per scan: the grid's
    x dimension is the number of scans of the profiled launch).  Fails loudly when the file is missing or unreadable: the roofline must
    not carry a made-up number."""
    import re
    txt = open(os.path.join(ROOT, path)).read()
    blocks = [b for b in txt.split("== ")[1:] if "k1_filter_fused" in b.split("\n", 1)[0] and "grid (" in b.split("\n", 1)[0]]
    if not blocks:
        raise RuntimeError(f"{path}: no k1_filter_fused launch in the ncu summary")
    b = blocks[0]
    grid = int(re.search(r"grid \((\d+),", b).group(1))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(key) + r"\s+([0-9.]+)\s+(\w+)", b)
        tot += float(m.group(1)) * unit[m.group(2)]
    return tot / grid, grid


def first_offsets(n_seq: int, n_frames: int, rank: int = 0):
    """Index of every sequence's first scan in the pool: stretch (j + rank) mod N_PLACES, starting frame (5 j) mod POOL_EXTRA."""
    j = np.arange(n_seq)
    return (((j + rank) % N_PLACES) * n_frames + (j * 5) % POOL_EXTRA).astype(np.int32)


def bind_to_gpu_numa_node(gpu_index: int) -> dict:
    """Host-side placement for the e2e leg: run this rank's threads on, and take its pinned upload buffers from, the
    NUMA node the GPU's PCIe root complex hangs off — so that eight ranks do not all pull their scans from node 0 across the socket
    interconnect.  Reads the GPU's node from sysfs (nvidia-smi gives the PCI bus id); sched_setaffinity to the node's CPUs and
    set_mempolicy(MPOL_PREFERRED, node) before any pinned allocation.  On a guest that exposes a single node (this pool's 1-GPU boxes do:
    profiles/r2b_host_topology_1gpu_box.txt) there is nothing to choose and the record says so."""
    info = {"nodes": 1, "gpu_node": None, "bound": False}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        info["nodes"] = len(nodes)
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=20).stdout.strip().lower()
        if bus.startswith("0000") and len(bus) > 12:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        info["gpu_node"] = node
        if len(nodes) > 1 and node in nodes:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
            import ctypes
            libc = ctypes.CDLL("libc.so.6", use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238            # x86-64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 16 * 64)
            info["bound"] = bool(allowed) and rc == 0
            info["cpus"] = len(allowed)
    except Exception as e:  # placement is an optimisation: never fatal
        info["error"] = str(e)[:120]
    return info


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
        # compensation of an overlapped step: x, y, azimuth in, x, y out, both clouds
        "k_compensate_polar": 18 * npts * 1.33,
        # fused cells kernel: points (x,y f32 + I u8) in, one 16-double record per valid cell out (all tables in shared memory)
        "cells_fused": 9 * npts + 128 * ncell,
        # legacy multi-kernel cells path (fine voxel grids only)
        "c5_cells": 9 * npts + 128 * nsmp,
        "c4_centroids": 12 * npts + 8 * nsmp,
        # (moving + keyframes) cells x 6 doubles used (u, n, N, planarity) in, one result record out
        "k_register": (1 + nkf) * ncell * 48 + 136,
    }
    return per[kernel] * n_seq if kernel in per else None


# ---------------------------------------------------------------------------------------------------------------------------------------
# loop_batch leg: the candidate loop of ScanContextClosure::SearchAndAddConstraint (loopclosure.cpp:658-724)
# ---------------------------------------------------------------------------------------------------------------------------------------
def loop_keyframe_rows(n_frames: int) -> np.ndarray:
    """Pool rows of the loop database's keyframes: the first LOOP_KF_PER_PLACE frames of every stretch."""
    return np.concatenate([p * n_frames + np.arange(LOOP_KF_PER_PLACE) for p in range(N_PLACES)])


def loop_candidates(gt_kf: np.ndarray, n_total: int, world: int, seed: int = 7):
    """The replicated candidate list: n_total (from, to) keyframe pairs with |from - to| <= 3 inside one stretch (overlapping views, like a
    revisit), `to` at its true pose, `from` at its true pose perturbed by U[+-1.5 m, +-1.5 m, +-0.1 rad] (the Scan-Context guess error
    scale).  Candidate i has from mod world == i mod world, so every rank's share is exactly n_total / world (weak scaling)."""
    n_kf = len(gt_kf)
    rng = np.random.default_rng(seed)
    r = np.arange(n_total) % world
    per = n_kf // world
    fr = (rng.integers(0, per, n_total) * world + r).astype(np.int32)
    within = fr % LOOP_KF_PER_PLACE
    d = rng.choice([-3, -2, -1, 1, 2, 3], n_total)
    w2 = np.where((within + d < 0) | (within + d >= LOOP_KF_PER_PLACE), within - d, within + d)
    to = (fr - within + w2).astype(np.int32)
    err = np.stack([rng.uniform(-1.5, 1.5, n_total), rng.uniform(-1.5, 1.5, n_total), rng.uniform(-0.1, 0.1, n_total)], axis=1)
    return fr, to, gt_kf[fr] + err, gt_kf[to]


def build_loop_sets(ctx, pool, kf_rows):
    """Keyframe cell sets through the product path (k-strongest filter + surface points on the GPU)."""
    sets = []
    for i in range(0, len(kf_rows), 16):
        rows = kf_rows[i:i + 16]
        f, _ = ctx.StructuredKStrongest(pool[rows], peaks=False)
        for b in range(len(rows)):
            az, rg, inten, x, y = f.scan(b)
            cells, _ = ctx.MapPointNormal(x, y, inten.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)
            sets.append(cells)
    return sets


def run_loop_leg(api, ctx, db, sets, gt_kf, world, rank, barrier, reduce_max, want_cpu: bool, iters: int = 40, warmup: int = 3):
    out = {"metric": "loop-closure candidate registrations/sec (P2L, Huber 0.1, SetParameters(4,10))", "unit": "pairs/s", "scaling": "weak",
           "keyframes": len(sets), "mean_cells": round(float(np.mean([len(c) for c in sets])), 1),
           "sharding": f"candidates by id_from mod {world}; database replicated; ONE ncclAllGather of 128-byte constraint records + device merge per batch inside the library; two batches in flight (tbv_loopdb_submit_sharded / _collect_sharded), every batch's records collected on the host",
           "batches": {}}
    for P in LOOP_PAIRS_PER_GPU:
        fr, to, Tf, Tt = loop_candidates(gt_kf, P * world, world)
        for _ in range(warmup):
            rec = db.register_sharded(fr, to, Tf, Tt)
        ctx.synchronize()
        barrier()
        l0 = ctx.launch_count()
        ph = np.zeros(4)
        t0 = time.perf_counter()
        # two batches in flight (tbv_loopdb_submit_sharded / _collect_sharded): the exchange of a batch runs under the registration of the next
        db.submit_sharded(fr, to, Tf, Tt)
        for _ in range(iters - 1):
            db.submit_sharded(fr, to, Tf, Tt)
            rec, tm = db.collect_sharded(want_timing=True)
            ph += tm
        rec, tm = db.collect_sharded(want_timing=True)
        ph += tm
        ctx.synchronize()
        sec = reduce_max(time.perf_counter() - t0)
        launches = ctx.launch_count() - l0
        barrier()
        ctx.profile_begin()                                    # one more iteration with an event after every launch: who takes the time
        db.register_sharded(fr, to, Tf, Tt)
        prof = {}
        for name, ms in ctx.profile_end():
            prof[name] = prof.get(name, 0.0) + ms
        ph /= iters
        b = {"pairs_per_gpu": P, "pairs_per_iter": P * world, "iters": iters, "value": round(P * world * iters / sec, 1), "ms_per_iter": round(sec / iters * 1e3, 4),
             "accepted": int(len(rec)), "gpu_launches": int(launches),
             "device_ms": {"h2d_register_pack": round(float(ph[0]), 4), "allgather_merge": round(float(ph[1]), 4), "d2h_records": round(float(ph[2]), 4),
                           "call": round(float(ph[3]), 4)},
             "allgather_us": round(prof.get("nccl_all_gather", 0.0) * 1e3, 2), "merge_us": round(prof.get("k_merge_constraints", 0.0) * 1e3, 2),
             "k_register_ms": round(prof.get("k_register", 0.0), 4),
             # cells of both scans of every candidate in (6 doubles used per cell), one 128-byte record out
             "k_register_algorithmic_GBps": round(P * (2 * out["mean_cells"] * 48 + 128) / max(prof.get("k_register", 1e-9), 1e-9) / 1e6, 1)}
        if want_cpu and P == LOOP_PAIRS_PER_GPU[-1]:
            from oracle import oracle_py   # checker + CPU baseline only
            n_cpu = min(64, P)
            acc = {int(c["candidate"]): c for c in rec}
            t0 = time.perf_counter()
            ref = [oracle_py.loop_register(sets[fr[q]], sets[to[q]], Tf[q], Tt[q]) for q in range(n_cpu)]
            cpu_s = time.perf_counter() - t0
            dxy = dth = 0.0
            agree = True
            for q, (ok, Ta, Tr, itrs, score) in enumerate(ref):
                agree &= (q in acc) == bool(ok)
                if ok and q in acc:
                    d = acc[q]["t_be"] - Ta
                    dxy = max(dxy, float(np.abs(d[:2]).max())); dth = max(dth, float(abs(np.arctan2(np.sin(d[2]), np.cos(d[2])))))
            out["cpu_baseline"] = {"value": round(n_cpu / cpu_s, 1), "unit": "pairs/s", "cores": 1, "kind": "port",
                                   "sample": f"first {n_cpu} of the {P} candidates, oracle loop_register, 1 thread"}
            out["parity_check"] = {"pairs": n_cpu, "accept_decisions_agree": bool(agree), "max_abs_xy_m": dxy, "max_abs_yaw_rad": dth}
        out["batches"][str(P)] = b
    out["value"] = out["batches"][str(LOOP_PAIRS_PER_GPU[0])]["value"]
    return out


# ---------------------------------------------------------------------------------------------------------------------------------------
# mulran leg: MulRan-shape scans in their wire layout, rotated on receipt on the device; odometry alone and mixed with
# one MU_LOOP_PAIRS-candidate-per-GPU loop batch per step on a second context (the reference's loop-closure thread)
# ---------------------------------------------------------------------------------------------------------------------------------------
def mulran_params(api):
    par = api.default_odom_params(radar_ccw=1)
    par.filter.range_res = MU_RANGE_RES
    return par


def mulran_first(n_seq: int, n_frames: int, rank: int):
    j = np.arange(n_seq)
    return (((j + rank) % MU_PLACES) * n_frames + (j * 5) % POOL_EXTRA).astype(np.int32)


def run_mulran_leg(api, parallel, torch, ctx, sets, gt_kf, world, rank, local, barrier, reduce_max, peak_hbm, want_cpu: bool):
    from tbv_slam_public_b200 import synth
    W, K, S = MU_WARMUP, MU_STEPS, MU_SEQS
    T = W + K
    n_frames = T + POOL_EXTRA
    pool = make_pool(n_frames, rank, dataset=synth.MULRAN, places=MU_PLACES)           # [places * n_frames][400][3360] azimuth-major
    wire = np.ascontiguousarray(np.rot90(pool, -1, axes=(1, 2)))                         # what the driver receives: [3360][400] per scan
    first = mulran_first(S, n_frames, rank)
    scan_bytes = N_AZ * MU_N_RANGE
    dev = torch.empty((T, S * scan_bytes), dtype=torch.uint8, device="cuda")
    for t in range(T):
        dev[t].copy_(torch.from_numpy(np.take(wire, first + t, axis=0).reshape(-1)))
    torch.cuda.synchronize()
    fuser = api.OdometryKeyframeFuser(ctx, S, N_AZ, MU_N_RANGE, mulran_params(api))
    fuser.set_wire_layout(True)
    fuser.set_overlap(OVERLAP)
    stream = torch.cuda.ExternalStream(ctx.stream)
    # ---- odometry alone, scans resident in HBM ------------------------------------------------------------------------------------------
    for t in range(W):
        fuser.step_dev(dev[t].data_ptr())
    ctx.synchronize(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(W, T):
        fuser.step_dev(dev[t].data_ptr())
    e1.record(stream)
    e1.synchronize(); ctx.synchronize()
    ms_alone = reduce_max(e0.elapsed_time(e1))
    outs = fuser.fetch()
    final = api.poses(outs).copy()
    npts = float(np.mean([o.n_points for o in outs]))
    if max(abs(o.status) for o in outs) != 0:
        raise RuntimeError("mulran leg: a per-scan capacity was exceeded")
    barrier()
    # per-kernel pass for the shape's K1 roofline
    fuser.reset()
    for t in range(W):
        fuser.step_dev(dev[t].data_ptr())
    ctx.synchronize()
    ctx.profile_begin()
    for t in range(W, W + 2):
        fuser.step_dev(dev[t].data_ptr())
    kern = {}
    for name, ms in ctx.profile_end():
        kern[name] = kern.get(name, 0.0) + ms / 2
    k1_bytes = S * (scan_bytes + 13 * npts * 1.33)
    k1_gbps = k1_bytes / kern["k1_filter_fused"] / 1e6
    # ---- mixed: every step also registers MU_LOOP_PAIRS candidates per GPU on a second context (sharded, all-gathered) ------------------
    ctx2 = api.Context(local)
    parallel.init_comm(ctx2)
    db2 = api.LoopDB(ctx2, len(sets), max(len(c) for c in sets))
    db2.add(sets)
    fr, to, Tf, Tt = loop_candidates(gt_kf, MU_LOOP_PAIRS * world, world, seed=11)
    fuser.reset()
    for t in range(W):
        fuser.step_dev(dev[t].data_ptr())
        db2.register_sharded(fr, to, Tf, Tt)
    ctx.synchronize(); ctx2.synchronize(); barrier()
    t0 = time.perf_counter()
    for t in range(W, T):
        fuser.step_dev(dev[t].data_ptr())          # asynchronous: the odometry step runs while the loop batches below are registered
        db2.submit_sharded(fr, to, Tf, Tt)         # two loop batches in flight: the one submitted in the previous step is collected now
        if t > W:
            rec = db2.collect_sharded()
    rec = db2.collect_sharded()
    ctx.synchronize(); ctx2.synchronize()
    sec_mixed = reduce_max(time.perf_counter() - t0)
    barrier()
    same = np.array_equal(api.poses(fuser.fetch()), final)
    if not same:
        raise RuntimeError("mulran leg: the mixed run changed the odometry result")
    out = {"workload": "MulRan-shape odometry (400 x 3360, 0.0595238 m, ccw, wire layout rotated on receipt on the device), CFEAR-3 filter, 4 keyframes, P2L; "
                       f"mixed = the same steps with one {MU_LOOP_PAIRS}-candidate-per-GPU sharded loop batch per step on a second context (configs[4])",
           "sequences_per_gpu": S, "steps": K, "warmup": W, "unit": UNIT,
           "odometry_alone": {"value": round(S * K * world / (ms_alone * 1e-3), 1), "ms_per_step": round(ms_alone / K, 4)},
           "mixed": {"scans_per_s": round(S * K * world / sec_mixed, 1), "pairs_per_s": round(MU_LOOP_PAIRS * world * K / sec_mixed, 1),
                     "ms_per_step": round(sec_mixed / K * 1e3, 4), "accepted_per_batch": int(len(rec)), "timing": "wall clock between synchronisations, max over ranks"},
           "roofline": {"kernel": "k1_filter_fused", "bound": "hbm", "achieved": round(k1_gbps, 2), "peak": peak_hbm, "unit": "GB/s",
                        "frac": round(k1_gbps / peak_hbm, 5), "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": round(kern["k1_filter_fused"], 4),
                        "traffic": None},
           "kernels_ms": {k: round(v, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])},
           "workload_stats": {"n_points": round(npts, 1), "n_cells": round(float(np.mean([o.n_cells for o in outs])), 1)}}
    if want_cpu:
        from oracle import oracle_py
        n_cpu = 6
        par = oracle_py.default_odom_params(radar_ccw=1, range_res=MU_RANGE_RES)
        sec, ref = oracle_py.odom_run_timed(par, pool, first[:n_cpu], W, T, 1)
        d = final[:n_cpu] - ref[:, T - 1, :]
        out["cpu_baseline"] = {"value": round(n_cpu * K / sec, 1), "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"{n_cpu} of the {S} sequences x {K} timed frames, oracle/ C++ port, 1 thread (scans already azimuth-major)"}
        out["parity_check"] = {"sequences": n_cpu, "frames": T, "max_abs_xy_m": float(np.abs(d[:, :2]).max()),
                               "max_abs_yaw_rad": float(np.abs(np.arctan2(np.sin(d[:, 2]), np.cos(d[:, 2]))).max())}
    fuser.close(); db2.close(); ctx2.close()
    del dev
    return out


OVERLAP = True   # set from --no-overlap in run_ours


def run_ours(args):
    global OVERLAP
    OVERLAP = not args.no_overlap
    import torch
    import torch.distributed as dist
    from tbv_slam_public_b200 import api, parallel, statistics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)     # before any pinned allocation
    k1_traffic_per_scan(K1_TRAFFIC_FILE)    # fail before any GPU work (and on every rank alike) when the committed ncu summary is unreadable
    if world > 1:
        import datetime
        # a rank that dies must not leave the others waiting for the default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    W, K, S = args.warmup, args.steps, args.seqs
    T = W + K
    pool, gt_pool = make_pool(T + POOL_EXTRA, rank, with_gt=True)
    first = first_offsets(S, T + POOL_EXTRA, rank)

    ctx = api.Context(local)
    par = api.default_odom_params()
    fuser = api.OdometryKeyframeFuser(ctx, S, N_AZ, N_RANGE, par)
    fuser.set_overlap(OVERLAP)   # the filter of step t + 1 on a second stream under the registration of step t (scans are resident / uploaded)
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

    def reduce_max(v: float) -> float:
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

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
    h2d_gbs_rank = step_bytes * K_e2e / e2e_s / 1e9               # this rank's host->device rate inside the e2e leg
    if world > 1:
        tt = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(tt[0]), float(tt[1])
        hh = torch.zeros(world, dtype=torch.float64, device="cuda")
        hh[rank] = h2d_gbs_rank
        dist.all_reduce(hh, op=dist.ReduceOp.SUM)
        h2d_gbs = [round(float(v), 2) for v in hh.cpu()]
    else:
        h2d_gbs = [round(h2d_gbs_rank, 2)]
    total_scans = S * K * world
    value = total_scans / (ms_total * 1e-3)
    e2e_value = S * K_e2e * world / e2e_s

    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm, peak_key = pick_hbm_peak(peaks)

    # ---------------- the other two legs (their own sub-objects; not part of `value`) ---------------------------
    fuser.close()
    del dev, host
    pinned.free()
    torch.cuda.empty_cache()
    loop_leg = mulran_leg = None
    if not args.no_extra_legs:
        want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
        if world > 1:
            parallel.init_comm(ctx)                                # the context's own NCCL communicator (tbv_comm_init_rank)
        kf_rows = loop_keyframe_rows(T + POOL_EXTRA)
        sets = build_loop_sets(ctx, pool, kf_rows)
        gt_kf = gt_pool[kf_rows]
        db = api.LoopDB(ctx, len(sets), max(len(c) for c in sets))
        db.add(sets)
        loop_leg = run_loop_leg(api, ctx, db, sets, gt_kf, world, rank, barrier, reduce_max, want_cpu)
        db.close()
        mulran_leg = run_mulran_leg(api, parallel, torch, ctx, sets, gt_kf, world, rank, local, barrier, reduce_max, peak_hbm, want_cpu)

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
            ctx.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        return
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
    traffic_per_scan, traffic_scans = k1_traffic_per_scan(K1_TRAFFIC_FILE)   # raises when the committed ncu summary is missing
    survey_gbps = SURVEY_K1_UNIT * S / kern[dominant] / 1e6
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": round(achieved, 2), "peak": peak_hbm, "unit": "GB/s",
                "frac": round(achieved / peak_hbm, 5),
                # dram__bytes_read.sum + dram__bytes_write.sum of one k1_filter_fused launch, per scan, from the committed ncu --set full summary
                "traffic": round(traffic_per_scan * S, 0),
                "traffic_source": f"{K1_TRAFFIC_FILE} ({traffic_scans} scans per launch: {traffic_per_scan:.0f} B per scan), scaled to {S} scans",
                # the same launch against SURVEY 8d's K1 unit (1 507 200 B read + 208 000 B written when every row keeps k = 40 points;
                # the synthetic rows keep fewer, so fewer bytes are actually written than this unit assumes)
                "survey_unit": {"bytes_per_scan": SURVEY_K1_UNIT, "achieved": round(survey_gbps, 2), "frac": round(survey_gbps / peak_hbm, 5)},
                "algorithmic_bytes_per_launch": b_dom, "launch_ms": round(kern[dominant], 4),
                # the whole filter stage = this kernel + the compensation launch of an overlapped step (fused into this kernel when overlap is off):
                # the same scan-in / clouds-out bytes over the time of both launches
                "filter_stage": {"kernels": [k for k in (dominant, "k_compensate_polar") if k in kern],
                                 "ms": round(kern[dominant] + kern.get("k_compensate_polar", 0.0), 4),
                                 "frac": round(b_dom / (kern[dominant] + kern.get("k_compensate_polar", 0.0)) / 1e6 / peak_hbm, 5)},
                "peak_source": f"MEASURED_PEAKS.json {peak_key} (of measured)" if peak_key else "fallback 6650 GB/s (of fallback)",
                "share_of_step": kernels[dominant]["share"], "longest_kernel": longest, "longest_kernel_share": kernels[longest]["share"],
                "step": {"algorithmic_bytes": step_bytes_alg, "achieved": round(step_bytes_alg / (ms_total / K) / 1e6, 2),
                         "frac": round(step_bytes_alg / (ms_total / K) / 1e6 / peak_hbm, 5)},
                "note": "launch duration from CUDA events recorded after every launch on the library's stream (tbv_profile_begin/_end)"}
    stages = {}
    for name, ms in kern.items():
        st_name = statistics.STAGE_OF_KERNEL.get(name, name)
        stages[st_name] = round(stages.get(st_name, 0.0) + ms / S, 6)
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 scan bytes -> f32 points -> f64 cells / normal equations / LM", "data": "synthetic",
        "config": bench_config(),
        "run": {"sequences_per_gpu": S, "scans_per_step": S * world, "gpus": world,
                "step_overlap": ("filter of step t+1 on a second stream under the registration of step t, compensation as its own launch (tbv_odom_set_overlap); "
                                 "per-kernel times come from a serialised pass, so their sum exceeds ms_per_step") if OVERLAP else
                                "off: single-stream steps, compensation fused into the filter kernel"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": step_bytes, "d2h_bytes_per_step": S * 80,
                "ms_per_step": round(e2e_s / K_e2e * 1e3, 4), "steps": K_e2e, "api": "tbv_odom_submit/tbv_odom_collect (pinned host scans, double-buffered)",
                "h2d_gbs_per_gpu": h2d_gbs, "host_placement": numa,
                "bound": "host->device link: every scan byte crosses PCIe once (1.5 MB per scan); the kernels need ~5 % of the step's copy time"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        # device ms per scan under the reference's timing keys (odometrykeyframefuser.cpp:253-256, radar_driver.cpp:87); compensation is
        # fused into the filter kernel here, so it is part of "Filtering"
        "stages_ms_per_scan": stages,
        "loop_batch": loop_leg, "mulran": mulran_leg,
        "cpu_baseline": cpu_baseline, "parity_check": parity, "workload_stats": {k: round(v, 3) if isinstance(v, float) else v for k, v in stats.items()},
    }
    emit(json.dumps(out))
    if world > 1:
        ctx.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def reference_stages(pool, gt_pool, n_frames: int, n_scans: int = 12):
    """One host thread, the oracle's primitives one stage at a time on consecutive frames of the first stretch: ms per scan under the
    reference's own timing keys (odometrykeyframefuser.cpp:253-256 "compensate" / "build_normals" / "register", radar_driver.cpp:87
    "Filtering").  The registration is scan-to-4-keyframes from the ground-truth guess of the bench workload."""
    from oracle import oracle_py
    t = {"Filtering": 0.0, "compensate": 0.0, "build_normals": 0.0, "register": 0.0}
    cells, n_reg = [], 0
    rp = oracle_py.default_reg_params()
    for f in range(n_scans):
        img = pool[f]
        mot = np.zeros(3) if f == 0 else np.array([2.5, 0.0, float(gt_pool[f][2] - gt_pool[f - 1][2])])
        t0 = time.perf_counter(); r = oracle_py.kstrongest(img, z_min=60.0, k=K_STRONGEST); t["Filtering"] += time.perf_counter() - t0
        az, rg, inten, x, y = r["filtered"]
        t0 = time.perf_counter(); x, y = oracle_py.compensate(x, y, mot, False); t["compensate"] += time.perf_counter() - t0
        t0 = time.perf_counter(); c, _ = oracle_py.build_cells(x, y, inten.astype(np.float32), radius=3.0, weight_intensity=True); t["build_normals"] += time.perf_counter() - t0
        cells.append(c)
        if f >= 4:
            scans = cells[f - 4:f] + [c]
            T = np.array([gt_pool[g] for g in range(f - 4, f + 1)], np.float64)
            t0 = time.perf_counter(); oracle_py.register(scans, T, rp); t["register"] += time.perf_counter() - t0
            n_reg += 1
    out = {k: round(v / n_scans * 1e3, 4) for k, v in t.items()}
    out["register"] = round(t["register"] / max(n_reg, 1) * 1e3, 4)
    out["note"] = f"oracle/ C++ port, 1 thread, mean over {n_scans} consecutive scans ({n_reg} registrations against 4 keyframes)"
    return out


def run_reference(args):
    """The reference's CPU path for the same workload: oracle/ (C++ port; the reference itself needs ROS/PCL/Ceres and cannot
    be built here), one worker per host thread over independent sequences — the reference's own scaling model."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py
    W, K = args.warmup, args.steps
    T = W + K
    threads = os.cpu_count() or oracle_py.hardware_threads() or 1
    n_seq = threads * max(1, int(round(20.0 / (0.0035 * T))))  # ~20 s of work per thread
    n_seq = min(n_seq, threads * 64)
    pool, gt_pool = make_pool(T + POOL_EXTRA, 0, with_gt=True)
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
    stages = loop_leg = mulran_leg = None
    try:
        stages = reference_stages(pool, gt_pool, T + POOL_EXTRA)
        # loop_batch on the CPU: the oracle's loopclosure::Register over the same keyframes / candidate rule, all host threads
        kf_rows = loop_keyframe_rows(T + POOL_EXTRA)
        sets = []
        for r in kf_rows:
            az, rg, inten, x, y = oracle_py.kstrongest(pool[r], z_min=60.0, k=K_STRONGEST, peaks=False)["filtered"]
            sets.append(oracle_py.build_cells(x, y, inten.astype(np.float32), radius=3.0, weight_intensity=True)[0])
        gt_kf = gt_pool[kf_rows]
        n_pairs = min(LOOP_PAIRS_PER_GPU[0], threads * 24)
        fr, to, Tf, Tt = loop_candidates(gt_kf, n_pairs, 1)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:     # the C call releases the GIL
            oks = list(ex.map(lambda q: oracle_py.loop_register(sets[fr[q]], sets[to[q]], Tf[q], Tt[q])[0], range(n_pairs)))
        lsec = time.perf_counter() - t0
        loop_leg = {"metric": "loop-closure candidate registrations/sec (P2L, Huber 0.1, SetParameters(4,10))", "unit": "pairs/s", "value": round(n_pairs / lsec, 1),
                    "pairs": n_pairs, "accepted": int(sum(bool(o) for o in oks)), "cores": threads, "kind": "port"}
        # MulRan-shape odometry on the CPU (scans already azimuth-major: the rotation on receipt is not timed)
        from tbv_slam_public_b200 import synth
        Tm = MU_WARMUP + MU_STEPS
        mpool = make_pool(Tm + POOL_EXTRA, 0, dataset=synth.MULRAN, places=MU_PLACES)
        n_mu = threads * 2
        msec, _ = oracle_py.odom_run_timed(oracle_py.default_odom_params(radar_ccw=1, range_res=MU_RANGE_RES), mpool, mulran_first(n_mu, Tm + POOL_EXTRA, 0),
                                           MU_WARMUP, Tm, threads)
        mulran_leg = {"unit": UNIT, "odometry_alone": {"value": round(n_mu * MU_STEPS / msec, 1)}, "sequences": n_mu, "steps": MU_STEPS, "cores": threads, "kind": "port"}
    except Exception as e:  # the extra legs must never take the arm down
        stages = stages or {"error": str(e)[:200]}
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": round(sec / K * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 scan bytes -> f32 points -> f64 cells / normal equations / LM", "data": "synthetic",
        "config": bench_config(),
        "run": {"sequences": n_seq, "threads": threads},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_seq} independent sequences x {K} timed frames (after {W} warm-up frames) over {threads} host threads; "
                                   "a step = one frame of every sequence"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "filtering_stage": filtering, "stages_ms_per_scan": stages, "loop_batch": loop_leg, "mulran": mulran_leg,
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
