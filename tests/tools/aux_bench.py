"""Per-call device time of the loop-closure side of the path at Oxford-sequence sizes (4.5 k keyframes, 5.4 k constraints,
10 ring-key candidates x 5 augmentations) next to the oracle's single-thread time for the same call.
Kernel times come from the library's per-launch events (tbv_profile_begin/_end) summed over the call's kernels; `call_ms` is the wall
time of the host-buffer C-ABI call (uploads + kernels + downloads)."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tbv_slam_public_b200 import api, synth  # noqa: E402
from oracle import oracle_py as o  # noqa: E402


def timed(ctx, fn, reps=5):
    fn()
    ctx.profile_begin()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    wall = (time.perf_counter() - t) / reps
    kern = {}
    for name, ms in ctx.profile_end():
        kern[name] = kern.get(name, 0.0) + ms / reps
    return wall * 1e3, kern


def cpu(fn, reps=3):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


def graph(n, rng):
    nodes = np.zeros((n, 7)); nodes[:, 6] = 1
    for i in range(n):
        th = 0.002 * i
        nodes[i, :3] = [i * 1.5 * math.cos(th), i * 1.5 * math.sin(th), 0]
        nodes[i, 3:] = [0, 0, math.sin(th / 2), math.cos(th / 2)]
    ids, meas = [], []
    for i in range(n - 1):
        pairs = [(i, i + 1, 0)] + ([(max(0, i - 2000), i + 1, 1)] if (i % 5 == 4 and i > 2000) else [])
        for (a, b, t) in pairs:
            qa, qb = nodes[a, 3:], nodes[b, 3:]
            tha, thb = 2 * math.atan2(qa[2], qa[3]), 2 * math.atan2(qb[2], qb[3])
            d = nodes[b, :3] - nodes[a, :3]
            c, s = math.cos(-tha), math.sin(-tha)
            dth = thb - tha + rng.normal(0, 0.01)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1] + rng.normal(0, 0.05), s * d[0] + c * d[1] + rng.normal(0, 0.05), 0, 0, 0, math.sin(dth / 2), math.cos(dth / 2)])
    return nodes, np.array(ids, np.int32), np.array(meas)


def main():
    ctx = api.Context(0)
    st = synth.make_stream(4)
    out = {}
    az, rg, I, x, y = o.kstrongest(st.scans[0], z_min=60.0, k=40)["peaks"]
    I = I.astype(np.float32)
    par, opar = api.default_sc_params(), o.default_sc_params()
    offs = ((0.0, 0.0), (0.0, -2.0), (0.0, 2.0), (0.0, -4.0), (0.0, 4.0))
    # K6: descriptor + keys, identity + 4 lateral augmentations (RadarScancontext.cpp:162-179)
    w, k = timed(ctx, lambda: api.sc_make(ctx, x, y, I, par, offs))
    c = cpu(lambda: [o.sc_make(x, y, I, opar, off) for off in offs])
    out["sc_make (5 descriptors, %d points)" % len(x)] = {"call_ms": w, "kernel_ms": sum(k.values()), "oracle_ms": c}
    # K7a: ring-key search over a 4.5 k-keyframe database, 5 queries
    rng = np.random.default_rng(0)
    n_db = 4500
    keys = rng.random((n_db, par.num_ring)).astype(np.float32)
    odom = np.cumsum(np.c_[np.full(n_db, 1.5), np.zeros(n_db), np.zeros(n_db)], axis=0)
    qk = rng.random((5, par.num_ring)).astype(np.float32)
    w, k = timed(ctx, lambda: api.sc_search(ctx, keys, odom, qk, np.full(5, n_db - 1, np.int32), par))
    out["sc_search (5 queries x %d keys)" % n_db] = {"call_ms": w, "kernel_ms": sum(k.values())}
    # K7b: 50 descriptor distances (10 candidates x 5 augmentations)
    desc = np.stack([api.sc_make(ctx, x + i, y, I, par)[0][0] for i in range(11)])
    qi, ci = np.zeros(50, np.int32), (np.arange(50) % 10 + 1).astype(np.int32)
    w, k = timed(ctx, lambda: api.sc_distance_batch(ctx, desc, desc, qi, ci, par))
    c = cpu(lambda: [o.sc_distance(desc[0], desc[j]) for j in ci])
    out["sc_distance_batch (50 pairs)"] = {"call_ms": w, "kernel_ms": sum(k.values()), "oracle_ms": c}
    # K8: pose-graph normal equations at Oxford size
    nodes, ids, meas = graph(4500, rng)
    w, k = timed(ctx, lambda: api.pgo_assemble(ctx, nodes, ids, meas), reps=3)
    c = cpu(lambda: o.pgo_assemble(nodes, ids, meas), reps=2)
    out["pgo_assemble (%d nodes, %d constraints)" % (len(nodes), len(ids))] = {"call_ms": w, "kernel_ms": sum(k.values()), "oracle_ms": c}
    # K8b: one LM step solve on the same graph (block-Jacobi PCG in one CTA) and the whole LM run
    _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas)
    info = {}
    def step():
        info["r"] = api.pgo_solve_step(ctx, ids, Hd, Ho, g, radius=1e4, rel_tol=1e-10)
    w, k = timed(ctx, step, reps=3)
    out["pgo_solve_step (%d nodes, rel_tol 1e-10, %d CG iterations)" % (len(nodes), info["r"][1])] = {"call_ms": w, "kernel_ms": sum(k.values())}
    import time as _t
    t0 = _t.perf_counter()
    _, S = api.pgo_optimize(ctx, nodes, ids, meas)
    out["pgo_optimize (%d nodes, Ceres default tolerances, %d LM iterations, %d CG iterations, %s)" % (
        len(nodes), S.iterations, S.cg_iterations, S.termination)] = {"call_ms": 1e3 * (_t.perf_counter() - t0), "cost_ratio": S.final_cost / S.initial_cost}
    # K1b: CA-CFAR on one Oxford scan
    w, k = timed(ctx, lambda: ctx.AzimuthCACFAR(st.scans[0], window_size=40, nb_guard_cells=10, capacity=65536))
    c = cpu(lambda: o.cacfar(st.scans[0], 40, 0.01, 10), reps=2)
    out["cacfar (one 400x3768 scan)"] = {"call_ms": w, "kernel_ms": sum(k.values()), "oracle_ms": c}
    for v in out.values():
        for kk in list(v):
            v[kk] = round(v[kk], 4)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
