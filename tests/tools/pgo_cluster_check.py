"""First GPU run of the thread-block-cluster PCG (k_pgo.cu: pgo_pcg_cluster, opt-in via TBV_PGO_CLUSTER=1) — correctness against scipy's sparse
direct solve and timing against the one-CTA kernel.  Written when the round's GPU budget was spent; run it FIRST next round:

    TBV_PGO_CLUSTER=1 python tests/tools/pgo_cluster_check.py      # cluster kernel (block-Jacobi PCG on 8 CTAs)
    TBV_PGO_CHAIN=1 python tests/tools/pgo_cluster_check.py        # odometry-chain preconditioner (one CTA): expect ~5 CG iterations
    python tests/tools/pgo_cluster_check.py                        # one-CTA block-Jacobi kernel, same checks, for the comparison

Prints one JSON line; exits 1 on a parity failure.  If it passes and is faster, make the cluster kernel the default in tbv_pgo_solve_step,
add its cases to tests/test_loop_gpu.py and update DESIGN.md §4 / §7b."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api  # noqa: E402
from test_loop_gpu import _damped_system, _graph  # noqa: E402


def main():
    import scipy.sparse.linalg as spl
    ctx = api.Context(0)
    out = {"kernel": "pgo_pcg_cr", "cases": []}
    ok = True
    for n, radius, fixed in ((2, 1e4, 0), (7, 1e4, 3), (30, 1e4, 0), (600, 1e4, 0), (600, 1e2, 17), (600, 1e8, 17), (4500, 1e4, 0), (4500, 1e2, 0)):
        rng = np.random.default_rng(n)
        nodes, ids, meas = _graph(n, rng)
        _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas, fixed_node=fixed)
        ctx.profile_begin()
        delta, iters, rel = api.pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node=fixed, radius=radius, max_iters=20000, rel_tol=1e-12)
        prof = dict(ctx.profile_end())
        delta2, iters2, _ = api.pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node=fixed, radius=radius, max_iters=20000, rel_tol=1e-12)
        A, b, keep = _damped_system(ids, Hd, Ho, g, radius, fixed)
        ref = spl.spsolve(A, b)
        got = delta.reshape(-1)[keep]
        case = {"nodes": n, "radius": radius, "cg_iterations": iters, "rel_residual": rel,
                "residual_check": float(np.linalg.norm(A @ got - b) / np.linalg.norm(b)),
                "delta_error": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
                "deterministic": bool(iters2 == iters and np.array_equal(delta, delta2)), "fixed_zero": bool(np.all(delta[fixed] == 0)),
                "kernel_ms": sum(prof.values()), "us_per_cg_iteration": 1e3 * sum(prof.values()) / max(iters, 1), "kernels": sorted(prof)}
        case["ok"] = bool(0 < iters < 20000 and rel <= 1e-12 and case["residual_check"] <= 1e-11 and case["delta_error"] <= 1e-5
                          and case["deterministic"] and case["fixed_zero"])
        ok = ok and case["ok"]
        out["cases"].append(case)
    out["ok"] = ok
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
