"""Times the pose-graph path at Oxford size (4.5 k nodes): normal-equation assembly, one LM step solve, a whole LM run.
Prints one JSON object. Run on a B200: python tools/pgo_bench.py"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api  # noqa: E402


def timed(ctx, fn, reps=3):
    fn()
    ctx.synchronize()
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


def graph(n, rng):
    """A 1.5 m / keyframe arc with noisy odometry and, past 2000 keyframes, a loop constraint to the keyframe 2000 earlier every 5th."""
    nodes = np.zeros((n, 7)); nodes[:, 6] = 1
    th = 0.002 * np.arange(n)
    nodes[:, 0], nodes[:, 1] = 1.5 * np.arange(n) * np.cos(th), 1.5 * np.arange(n) * np.sin(th)
    nodes[:, 5], nodes[:, 6] = np.sin(th / 2), np.cos(th / 2)
    ids, meas = [], []
    back = min(2000, n // 2)
    for i in range(n - 1):
        for (a, b, t) in [(i, i + 1, 0)] + ([(i - back, i + 1, 1)] if (i % 5 == 4 and i > back) else []):
            d = nodes[b, :3] - nodes[a, :3]
            c, s = math.cos(-th[a]), math.sin(-th[a])
            dth = th[b] - th[a] + rng.normal(0, 0.01)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1] + rng.normal(0, 0.05), s * d[0] + c * d[1] + rng.normal(0, 0.05), 0, 0, 0, math.sin(dth / 2), math.cos(dth / 2)])
    return nodes, np.array(ids, np.int32), np.array(meas)


def main():
    ctx = api.Context(0)
    rng = np.random.default_rng(0)
    out = {}
    for n in (600, 4500):
        nodes, ids, meas = graph(n, rng)
        w, k = timed(ctx, lambda: api.pgo_assemble(ctx, nodes, ids, meas), reps=3)
        _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas)
        row = {"constraints": int(len(ids)), "assemble_call_ms": round(w, 3), "assemble_kernel_ms": round(sum(k.values()), 4)}
        for radius in (1e4, 1e2, 1e8):
            res = {}
            def step():
                res["r"] = api.pgo_solve_step(ctx, ids, Hd, Ho, g, radius=radius, rel_tol=1e-10)
            w, k = timed(ctx, step, reps=3)
            it = res["r"][1]
            row["solve_step_radius_%g" % radius] = {"cg_iterations": it, "call_ms": round(w, 3), "kernel_ms": round(k.get("pgo_pcg_cr", 0.0), 3),
                                                    "us_per_cg_iteration": round(1e3 * k.get("pgo_pcg_cr", 0.0) / max(it, 1), 3)}
        api.pgo_optimize_device(ctx, nodes, ids, meas)    # warm-up
        t0 = time.perf_counter()
        _, S = api.pgo_optimize_device(ctx, nodes, ids, meas)                  # Ceres defaults (200 iterations, 1e-6 / 1e-10 / 1e-8), whole loop on the device
        row["optimize_device"] = {"wall_ms": round(1e3 * (time.perf_counter() - t0), 2), "device_ms": round(S.device_ms, 3), "lm_iterations": S.iterations,
                                  "successful_steps": S.successful_steps, "cg_iterations": S.cg_iterations, "termination": S.termination,
                                  "cost": [S.initial_cost, S.final_cost]}
        t0 = time.perf_counter()
        _, Sh = api.pgo_optimize_ceres(ctx, nodes, ids, meas)                  # the same rules driven from the host, blocks crossing PCIe every iteration
        row["optimize_host_driven"] = {"wall_ms": round(1e3 * (time.perf_counter() - t0), 1), "lm_iterations": Sh.iterations, "cg_iterations": Sh.cg_iterations,
                                       "termination": Sh.termination, "cost": [Sh.initial_cost, Sh.final_cost]}
        row["reference_cpu_ms"] = "1230 (CeresLeastSquares on a 4.5 k-node Oxford graph, one CPU thread; SURVEY 6)" if n == 4500 else None
        out["%d nodes" % n] = row
        print(json.dumps({"%d nodes" % n: row}), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
