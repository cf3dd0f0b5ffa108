"""BASELINE configs[2] / [4]: batched loop-closure candidate registration (256 scan pairs per iteration by default), on 1 GPU
or sharded over N GPUs with the accepted constraints all-gathered over NCCL (tbv_slam_public_b200/parallel.py).

  python tests/tools/loop_bench.py [--pairs 256] [--iters 20] [--keyframes 64]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/loop_bench.py --pairs 1024

Keyframes: consecutive frames of the synthetic Oxford-shape stream, filtered (k=40, z_min=60) and turned into cells on the
GPU.  Candidates: (from, to) with |from - to| <= 3 (overlapping views, like a revisit), `to` placed at its true pose and
`from` at its true pose perturbed by U[+-1.5 m, +-1.5 m, +-0.1 rad] (the Scan-Context guess error scale).
Prints ONE JSON line on rank 0: pairs/s (device events, max over ranks), the CPU oracle's rate on a sample, and the pose
parity of that sample.  Not the headline bench (that is bench.py); committed under profiles/ as evidence for 8e."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tbv_slam_public_b200 import api, parallel, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--keyframes", type=int, default=64)
    ap.add_argument("--cpu-pairs", type=int, default=64)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    st = synth.make_stream(a.keyframes)
    # keyframe cells through the product path (filter + cells on the GPU); every rank holds the whole database
    sets = []
    for i in range(0, a.keyframes, 16):
        f, _ = ctx.StructuredKStrongest(st.scans[i:i + 16], peaks=False)
        for b in range(len(st.scans[i:i + 16])):
            az, rg, inten, x, y = f.scan(b)
            cells, _ = ctx.MapPointNormal(x, y, inten.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)
            sets.append(cells)
    db = api.LoopDB(ctx, a.keyframes, max(len(c) for c in sets))
    db.add(sets)
    rng = np.random.default_rng(7)
    fr = rng.integers(0, a.keyframes, a.pairs)
    to = np.clip(fr + rng.choice([-3, -2, -1, 1, 2, 3], a.pairs), 0, a.keyframes - 1)
    to = np.where(to == fr, np.where(fr > 0, fr - 1, fr + 1), to)
    err = np.stack([rng.uniform(-1.5, 1.5, a.pairs), rng.uniform(-1.5, 1.5, a.pairs), rng.uniform(-0.1, 0.1, a.pairs)], axis=1)
    Tf, Tt = st.gt[fr] + err, st.gt[to]

    slc = parallel.ShardedLoopClosure(db)
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(a.warmup):
        out = slc.register_candidates(fr, to, Tf, Tt)
    ctx.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters):
        out = slc.register_candidates(fr, to, Tf, Tt)      # includes the all-gather and the D2H of the gathered records
    e1.record(stream)
    e1.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        dist.barrier()
    if rank == 0:
        from oracle import oracle_py   # checker + CPU baseline only
        n_cpu = min(a.cpu_pairs, a.pairs)
        acc = {int(c["candidate"]): c for c in out}
        t0 = time.perf_counter()
        ref = [oracle_py.loop_register(sets[fr[p]], sets[to[p]], Tf[p], Tt[p]) for p in range(n_cpu)]
        cpu_s = time.perf_counter() - t0
        dxy = dth = 0.0
        agree = True
        for p, (ok, Ta, Tr, itrs, score) in enumerate(ref):
            agree &= (p in acc) == bool(ok)
            if ok and p in acc:
                d = acc[p]["t_be"] - Ta
                dxy = max(dxy, float(np.abs(d[:2]).max())); dth = max(dth, float(abs(np.arctan2(np.sin(d[2]), np.cos(d[2])))))
        print(json.dumps({
            "metric": "loop-closure candidate registrations/sec (P2L, Huber 0.1, SetParameters(4,10))", "value": round(a.pairs * a.iters / (ms * 1e-3), 1),
            "unit": "pairs/s", "n_gpus": world, "pairs_per_iter": a.pairs, "iters": a.iters, "ms_per_iter": round(ms / a.iters, 4),
            "accepted": int(len(out)), "keyframes": a.keyframes, "mean_cells": float(np.mean([len(c) for c in sets])),
            "gpu_launches": int(launches), "sharding": f"candidates by id_from mod {world}; database replicated; one all-gather of 128-byte constraint records per iteration",
            "cpu_baseline": {"value": round(n_cpu / cpu_s, 1), "unit": "pairs/s", "cores": 1, "kind": "port", "sample": f"first {n_cpu} pairs, oracle loop_register, 1 thread"},
            "parity_check": {"pairs": n_cpu, "accept_decisions_agree": bool(agree), "max_abs_xy_m": dxy, "max_abs_yaw_rad": dth}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
