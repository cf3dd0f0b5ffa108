"""Host side of the pose-graph optimiser (api.pgo_plus / pgo_optimize's trust-region bookkeeping) without a GPU.

The device calls (tbv_pgo_assemble, tbv_pgo_solve_step, tbv_pgo_solve_damped) are replaced HERE, in the test only, by the oracle's assembly and scipy's sparse
direct solve, so the Levenberg-Marquardt loop (ceresoptimizer.cpp:50-62 + Ceres 2.1.0 trust_region_minimizer defaults) is exercised on CPU;
the device versions are checked against the same checkers in tests/test_loop_gpu.py.
"""
import math

import numpy as np
import pytest

from tbv_slam_public_b200 import api


def _graph(n, rng, noise=1.0):
    nodes = np.zeros((n, 7)); nodes[:, 6] = 1
    for i in range(n):
        th = 0.07 * i
        nodes[i, :3] = [i * 1.2, 0.1 * i * i, 0]
        nodes[i, 3:] = [0, 0, math.sin(th / 2), math.cos(th / 2)]
    ids, meas = [], []
    for i in range(n - 1):
        for (a, b, t) in [(i, i + 1, 0)] + ([(max(0, i - 7), i + 1, 1)] if i % 3 == 2 else []):
            tha, thb = 0.07 * a, 0.07 * b
            d = nodes[b, :3] - nodes[a, :3]
            c, s = math.cos(-tha), math.sin(-tha)
            dth = thb - tha + noise * rng.normal(0, 0.01)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1] + noise * rng.normal(0, 0.05), s * d[0] + c * d[1] + noise * rng.normal(0, 0.05), 0, 0, 0,
                         math.sin(dth / 2), math.cos(dth / 2)])
    truth = nodes.copy()
    nodes[1:, :3] += rng.normal(0, 0.05, size=(n - 1, 3))
    return truth, nodes, np.array(ids, np.int32), np.array(meas)


@pytest.fixture()
def cpu_device(monkeypatch):
    """Stand-ins for the two device calls (test infrastructure only)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from oracle import oracle_py
    oracle_py.lib()

    def assemble(ctx, x, ids, meas, params=None, info=None, fixed_node=0):
        return oracle_py.pgo_assemble(x, ids, meas, fixed_node=fixed_node)

    def solve(ctx, ids, Hd, Ho, g, fixed_node=0, radius=1e4, max_iters=0, rel_tol=0):
        n, r6 = len(Hd), np.arange(6)
        bi = lambda i: (6 * i[:, None, None] + r6[None, :, None]) + 0 * r6[None, None, :]
        bj = lambda j: (6 * j[:, None, None] + r6[None, None, :]) + 0 * r6[None, :, None]
        nn, a, b = np.arange(n), ids[:, 0].astype(np.int64), ids[:, 1].astype(np.int64)
        A = sp.coo_matrix((np.r_[Hd.ravel(), Ho.ravel(), Ho.ravel()],
                           (np.r_[bi(nn).ravel(), bi(a).ravel(), bj(b).ravel()], np.r_[bj(nn).ravel(), bj(b).ravel(), bi(a).ravel()])),
                          shape=(6 * n, 6 * n)).tocsr()
        A = A + sp.diags(np.clip(A.diagonal(), 1e-6, 1e32) / radius)
        keep = np.r_[0:6 * fixed_node, 6 * fixed_node + 6:6 * n]
        x = np.zeros(6 * n)
        x[keep] = spl.spsolve(A[keep][:, keep].tocsc(), -g.reshape(-1)[keep])
        return x.reshape(n, 6), 1, 0.0

    def solve_damped(ctx, ids, Hd, Ho, g, damping, fixed_node=0, max_iters=0, rel_tol=0):
        # (H + diag(damping)) delta = -g: the radius interface with a diagonal chosen so that clamp(e) / 1 adds exactly `damping`
        n, idx = len(Hd), np.arange(6)
        Hd2 = np.array(Hd, np.float64).reshape(-1, 6, 6).copy()
        total = Hd2[:, idx, idx] + np.asarray(damping, np.float64).reshape(-1, 6)
        e = np.where(total / 2.0 >= 1e-6, total / 2.0, total - 1e-6)
        Hd2[:, idx, idx] = np.where(e > 1e32, total - 1e32, e)
        return solve(ctx, ids, Hd2.reshape(n, 36).reshape(n, 6, 6), Ho, g, fixed_node, 1.0)

    monkeypatch.setattr(api, "pgo_assemble", assemble)
    monkeypatch.setattr(api, "pgo_solve_step", solve)
    monkeypatch.setattr(api, "pgo_solve_damped", solve_damped)
    return oracle_py


def test_pgo_plus_matches_rotation_composition():
    rng = np.random.default_rng(0)
    q = rng.normal(size=(6, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    nodes = np.c_[rng.normal(size=(6, 3)), q]
    d = rng.normal(0, 0.3, size=(6, 6)); d[2] = 0

    def rot(qq):
        x, y, z, w = qq
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    out = api.pgo_plus(nodes, d)
    assert np.array_equal(out[2], nodes[2])
    assert np.allclose(np.linalg.norm(out[:, 3:], axis=1), 1.0, atol=1e-14)
    for i in range(6):
        th = np.linalg.norm(d[i, 3:])
        dq = np.r_[np.sin(th) * d[i, 3:] / th, np.cos(th)] if th > 0 else np.array([0, 0, 0, 1.0])
        assert np.allclose(rot(out[i, 3:]), rot(dq) @ rot(nodes[i, 3:]), atol=1e-12)
        assert np.allclose(out[i, :3], nodes[i, :3] + d[i, :3])


def test_pgo_plus_agrees_with_the_oracle_tangent_jacobian(cpu_device):
    """Finite differences of the oracle's cost along pgo_plus directions reproduce its tangent-space gradient: the (+) convention of the
    host update is the one the assembled normal equations are expressed in."""
    rng = np.random.default_rng(4)
    _, nodes, ids, meas = _graph(12, rng)
    c0, _, _, g, _ = cpu_device.pgo_assemble(nodes, ids, meas)
    for _ in range(6):
        d = np.zeros((12, 6)); i, k = rng.integers(1, 12), rng.integers(0, 6)
        h = 1e-6
        d[i, k] = h
        cp = cpu_device.pgo_assemble(api.pgo_plus(nodes, d), ids, meas)[0]
        cm = cpu_device.pgo_assemble(api.pgo_plus(nodes, -d), ids, meas)[0]
        assert abs((cp - cm) / (2 * h) - g[i, k]) <= 1e-5 * max(1.0, abs(g[i, k]))


def test_hessian_product_matches_dense(cpu_device):
    rng = np.random.default_rng(1)
    _, nodes, ids, meas = _graph(15, rng)
    _, Hd, Ho, g, _ = cpu_device.pgo_assemble(nodes, ids, meas)
    H = np.zeros((90, 90))
    for i in range(15):
        H[6 * i:6 * i + 6, 6 * i:6 * i + 6] = Hd[i]
    for c, (a, b, _t) in enumerate(ids):
        H[6 * a:6 * a + 6, 6 * b:6 * b + 6] += Ho[c]
        H[6 * b:6 * b + 6, 6 * a:6 * a + 6] += Ho[c].T
    x = rng.normal(size=(15, 6))
    assert np.allclose(api._pgo_hessian_times(ids, Hd, Ho, x).ravel(), H @ x.ravel(), rtol=1e-12, atol=1e-9)
    assert np.allclose(H, H.T)


def test_lm_loop_recovers_consistent_graph(cpu_device):
    rng = np.random.default_rng(7)
    truth, nodes, ids, meas = _graph(40, rng, noise=0.0)
    x, S = api.pgo_optimize(None, nodes, ids, meas, function_tolerance=1e-16, gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    assert S.final_cost <= 1e-14 * S.initial_cost and S.successful_steps >= 2
    assert np.abs(x[:, :3] - truth[:, :3]).max() <= 1e-6
    assert np.array_equal(x[0], nodes[0])


def test_lm_loop_terminations(cpu_device):
    rng = np.random.default_rng(8)
    _, nodes, ids, meas = _graph(40, rng)
    x, S = api.pgo_optimize(None, nodes, ids, meas)
    assert S.termination == "function_tolerance" and S.final_cost < S.initial_cost and 1 <= S.successful_steps <= S.iterations <= 200
    _, S1 = api.pgo_optimize(None, nodes, ids, meas, max_num_iterations=1)
    assert S1.iterations == 1 and S1.termination == "max_num_iterations" and S1.final_cost <= S1.initial_cost
    _, S2 = api.pgo_optimize(None, x, ids, meas, gradient_tolerance=1e30)        # already below the gradient tolerance: no iteration
    assert S2.iterations == 0 and S2.termination == "gradient_tolerance"
    xt, St = api.pgo_optimize(None, nodes, ids, meas, function_tolerance=1e-15, gradient_tolerance=1e-9, parameter_tolerance=1e-15)
    g = cpu_device.pgo_assemble(xt, ids, meas)[3]
    assert np.abs(g).max() <= 1e-8 and St.final_cost <= S.final_cost



def test_ceres_restatement_converges_and_does_not_take_its_last_step(cpu_device):
    rng = np.random.default_rng(7)
    truth, nodes, ids, meas = _graph(40, rng, noise=0.0)
    x, S = api.pgo_optimize_ceres(None, nodes, ids, meas, function_tolerance=1e-16, gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    assert S.final_cost <= 1e-14 * S.initial_cost and S.successful_steps >= 2 and np.abs(x[:, :3] - truth[:, :3]).max() <= 1e-6
    assert np.array_equal(x[0], nodes[0])
    rng = np.random.default_rng(8)
    _, nodes, ids, meas = _graph(40, rng)
    x, S = api.pgo_optimize_ceres(None, nodes, ids, meas)
    assert S.termination in ("function_tolerance", "gradient_tolerance", "parameter_tolerance") and S.final_cost < S.initial_cost
    # the returned point is the last ACCEPTED one: its cost is the summary's final cost (the candidate that triggered the stop is dropped)
    assert cpu_device.pgo_assemble(x, ids, meas)[0] == pytest.approx(S.final_cost, rel=1e-12)
    assert S.iterations >= S.successful_steps + (1 if S.termination != "gradient_tolerance" else 0)
    xd, Sd = api.pgo_optimize(None, nodes, ids, meas)
    assert S.final_cost == pytest.approx(Sd.final_cost, rel=1e-3)            # both drivers end in the same valley
    _, S1 = api.pgo_optimize_ceres(None, nodes, ids, meas, max_num_iterations=1)
    assert S1.iterations == 1 and S1.termination == "max_num_iterations"
    _, S0 = api.pgo_optimize_ceres(None, x, ids, meas, gradient_tolerance=1e30)
    assert S0.iterations == 0 and S0.termination == "gradient_tolerance"


def test_chain_preconditioned_solve_prototype(cpu_device):
    """tests/tools/pgo_chain_prototype.py — the block-tridiagonal (odometry chain) preconditioner planned for the next kernel revision — is an
    exact solver of the same damped system (vs scipy's sparse LU) and needs two orders of magnitude fewer CG iterations than block-Jacobi."""
    import os
    import sys
    import scipy.sparse.linalg as spl
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "tools"))
    import pgo_chain_prototype as P
    from test_loop_gpu import _damped_system, _graph as big_graph
    for n, radius, fixed, cap in ((2, 1e4, 0, 3), (30, 1e4, 0, 12), (600, 1e4, 0, 12), (600, 1e8, 17, 120), (1500, 1e4, 0, 12)):
        rng = np.random.default_rng(n)
        nodes, ids, meas = big_graph(n, rng)
        _, Hd, Ho, g, _ = cpu_device.pgo_assemble(nodes, ids, meas, fixed_node=fixed)
        delta, iters, rel = P.solve(ids, Hd, Ho, g, fixed, radius, rel_tol=1e-12)
        A, b, keep = _damped_system(ids, Hd, Ho, g, radius, fixed)
        ref = spl.spsolve(A, b)
        got = delta.reshape(-1)[keep]
        assert rel <= 1e-12 and iters <= cap, (n, radius, iters)
        assert np.all(delta[fixed] == 0) and np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref)   # the bar of tbv_pgo_solve_step
    # without loop constraints the chain IS the matrix: one iteration
    nodes, ids, meas = big_graph(200, np.random.default_rng(1))
    odo = ids[:, 2] == 0
    _, Hd, Ho, g, _ = cpu_device.pgo_assemble(nodes, ids[odo], meas[odo])
    _, iters, rel = P.solve(ids[odo], Hd, Ho, g, 0, 1e4, rel_tol=1e-10)
    assert iters <= 2 and rel <= 1e-10


def test_parallel_cyclic_reduction_equals_the_sweeps(cpu_device):
    """pgo_chain_prototype.pcr_*: the log2(n)-level, fully parallel form of the chain preconditioner gives the block-Thomas result (to the
    conditioning of the chain: 1e-8 at Ceres' initial radius, 1e-5 at radius 1e8) — the checker of the next kernel revision's sweeps."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "tools"))
    import pgo_chain_prototype as P
    from test_loop_gpu import _graph as big_graph
    for n, radius, fixed, tol in ((2, 1e4, 0, 1e-12), (3, 1e4, 1, 1e-12), (37, 1e4, 5, 1e-9), (600, 1e4, 0, 1e-8), (600, 1e8, 17, 1e-5), (1500, 1e4, 0, 1e-8)):
        rng = np.random.default_rng(n)
        nodes, ids, meas = big_graph(n, rng)
        _, Hd, Ho, g, _ = cpu_device.pgo_assemble(nodes, ids, meas, fixed_node=fixed)
        Ad, C = P.chain_blocks(ids, Hd, Ho, radius, fixed)
        r = rng.normal(size=(n, 6))
        want = P.apply(*P.factorise(Ad, C), r)
        levels, Dinv = P.pcr_factorise(Ad, C)
        assert len(levels) == int(np.ceil(np.log2(n)))
        got = P.pcr_apply(levels, Dinv, r)
        assert np.abs(got - want).max() <= tol * np.abs(want).max(), (n, radius)
        # and it is M^-1: multiplying back by the block-tridiagonal matrix returns r
        back = np.einsum("nij,nj->ni", Ad, got)
        back[1:] += np.einsum("nij,nj->ni", C, got[:-1])
        back[:-1] += np.einsum("nji,nj->ni", C, got[1:])
        scale = max(np.abs(Ad).max(), np.abs(C).max() if len(C) else 0.0) * np.abs(got).max()      # |M| |z|: the size of the terms that cancel
        assert np.abs(back - r).max() <= 1e-9 * scale + 1e3 * tol * np.abs(r).max()
