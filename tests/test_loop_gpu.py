"""K6/K7/K8 parity: Scan-Context descriptor / keys / candidate search / distance and pose-graph normal equations vs the oracle.

Bar: descriptor bins, ring keys, shifts and candidate indices are index work -> exact (descriptor: up to 2 points per
descriptor may land in the neighbouring sector, because the sector index goes through a float atan whose last bit is
libm-specific — glibc picks an FMA or non-FMA atanf by CPU); distances / similarities / PGO blocks within 1e-12 relative.
"""
import math

import numpy as np
import pytest

from tbv_slam_public_b200 import api, synth

pytestmark = pytest.mark.gpu


def _peaks(oracle, img):
    az, rg, I, x, y = oracle.kstrongest(img)["peaks"]
    return x, y, I.astype(np.float32)


def _desc_close(ref, got, max_moved_points=2):
    diff = np.abs(ref - got) > 1e-12
    assert diff.sum() <= 2 * max_moved_points, f"{diff.sum()} descriptor bins differ"
    assert abs(ref[ref > -0.5].sum() - got[got > -0.5].sum()) < 1e-9   # intensity mass is conserved


def test_sc_make_descriptor_and_keys(ctx, oracle, stream8):
    P, OP = api.default_sc_params(), oracle.default_sc_params()
    for i in (0, 3):
        x, y, I = _peaks(oracle, stream8.scans[i])
        desc, rk, sk = api.sc_make(ctx, x, y, I, P, api.AUGMENTS)
        for a, off in enumerate(api.AUGMENTS):
            d_ref, rk_ref, sk_ref = oracle.sc_make(x, y, I, OP, off)
            _desc_close(d_ref, desc[a])
            if np.array_equal(d_ref, desc[a]):
                assert np.array_equal(rk_ref, rk[a]) and np.array_equal(sk_ref, sk[a])
            else:
                assert np.allclose(rk_ref, rk[a], atol=1e-3) and np.allclose(sk_ref, sk[a], atol=1e-2)
            assert desc[a].min() == -1.0   # quirk: empty bins = -1000 / 1000


def test_sc_make_max_and_unit_divider(ctx, oracle, stream8):
    x, y, I = _peaks(oracle, stream8.scans[1])
    for fn, div in [(1, 1000.0), (0, 1.0), (1, 1.0)]:
        P, OP = api.default_sc_params(desc_function=fn, desc_divider=div, no_point=0.0), oracle.default_sc_params(desc_function=fn, desc_divider=div)
        desc, rk, sk = api.sc_make(ctx, x, y, I, P)
        d_ref, _, _ = oracle.sc_make(x, y, I, OP)
        diff = np.abs(d_ref - desc[0]) > 1e-12
        assert diff.sum() <= 4
        if div == 1.0:
            assert desc[0].min() == 0.0   # divider 1: the no-point reset does fire


def test_sc_distance_batch(ctx, oracle, stream8):
    descs = []
    for i in range(6):
        x, y, I = _peaks(oracle, stream8.scans[i])
        descs.append(oracle.sc_make(x, y, I)[0])
    rng = np.random.default_rng(0)
    rolled = [np.roll(d.reshape(120, 40), s, axis=0).reshape(-1) for d, s in zip(descs, (0, 5, 17, 60, 100, 119))]
    qi, ci = [], []
    for a in range(6):
        for b in range(6):
            qi.append(a); ci.append(b)
    dist, shift = api.sc_distance_batch(ctx, descs, rolled, qi, ci)
    for p, (a, b) in enumerate(zip(qi, ci)):
        d_ref, s_ref = oracle.sc_distance(descs[a], rolled[b])
        assert s_ref == shift[p]
        assert abs(d_ref - dist[p]) <= 1e-12 * max(1.0, abs(d_ref))
    for a, s in enumerate((0, 5, 17, 60, 100, 119)):
        p = a * 6 + a
        assert dist[p] < 1e-12 and shift[p] == (120 - s) % 120
    # degenerate: empty descriptors (every column has zero norm)
    z = np.zeros(4800)
    d, s = api.sc_distance_batch(ctx, [z], [z], [0], [0])
    assert (d[0], s[0]) == oracle.sc_distance(z, z)


def _trajectory(n):
    gt = synth.figure8(n, speed=10.0)
    return np.stack([synth.se2_mul(synth.se2_inv(gt[0]), g) for g in gt])


def test_rsc_manager_detects_like_the_oracle(ctx, oracle):
    """A revisiting trajectory: 40 keyframes along the figure-8, then the first 8 places again."""
    st = synth.make_stream(12)
    odom = _trajectory(12)
    order = list(range(12)) * 3 + list(range(8))      # revisits produce loop candidates
    rng = np.random.default_rng(1)
    for n_cand in (1, 3):
        OP, P = oracle.default_sc_params(n_candidates=n_cand), api.default_sc_params(n_candidates=n_cand)
        ref, gpu = oracle.RSC(OP), api.RSCManager(ctx, P)
        pose = np.zeros(3)
        n_with = n_override = 0
        for k, i in enumerate(order):
            x, y, I = _peaks(oracle, st.scans[i])
            pose = pose + np.array([2.5 * math.cos(0.05 * k), 2.5 * math.sin(0.05 * k), 0.05])   # a growing spiral: old places are far in odometry
            ref.add(x, y, I, pose); gpu.makeAndSaveScancontextAndKeysRadarCloud(x, y, I, pose)
            # K6's only tolerance (a point on a sector edge, see the module docstring) must not leak into the K7 comparison:
            # where a GPU descriptor differs from the oracle's by such a point, continue with the oracle's for this keyframe.
            desc, rk, offs = gpu.queries
            for a, off in enumerate(offs):
                d_ref, rk_ref, _ = oracle.sc_make(x, y, I, OP, off)
                if not np.array_equal(d_ref, desc[a]):
                    _desc_close(d_ref, desc[a])
                    n_override += 1
                    desc[a], rk[a] = d_ref, rk_ref
            gpu.polarcontexts[-1], gpu.ringkeys[-1] = desc[0].copy(), rk[0].copy()
            r = ref.detect(); g = gpu.detectLoopClosureID()
            assert len(r) == len(g), f"keyframe {k}"
            for a, b in zip(r, g):
                assert int(a[4]) == b["nn_idx"] and int(a[5]) == b["argmin_shift"] and int(a[6]) == b["aug_idx"]
                assert abs(a[0] - b["min_dist"]) < 1e-9 and abs(a[1] - b["min_dist_sc"]) < 1e-9 and abs(a[2] - b["min_dist_odom"]) < 1e-12
                assert abs(a[3] - b["yaw_diff_rad"]) < 1e-7
            n_with += len(g) > 0
        assert n_with > 20
        assert n_override <= 0.05 * 5 * len(order)


def test_sc_search_batched_as_of_queries(ctx, oracle):
    """One batched launch answering 'as of keyframe c' for many c == the incremental oracle, incl. the exclusion window."""
    rng = np.random.default_rng(7)
    n = 60
    keys = rng.uniform(0, 1, size=(n, 40)).astype(np.float32)
    odom = np.cumsum(np.c_[rng.uniform(0.5, 3.0, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-0.1, 0.1, n)], axis=0)
    ref = oracle.RSC(oracle.default_sc_params(augment_sc=0))
    want = {}
    cloud = (np.float32([1.0]), np.float32([1.0]), np.float32([100.0]))
    for c in range(n):
        ref.add(*cloud, odom[c])
        ne, sim = ref.state()
        want[c] = (ne, sim.copy())
    cur = np.arange(n, dtype=np.int32)
    ci, cs, ne = api.sc_search(ctx, keys, odom, keys, cur)
    for c in range(n):
        assert ne[c] == want[c][0]
        n_search = max(0, c - 1 - ne[c])
        sim = want[c][1]
        # independent numpy restatement of L2norm + top-10
        k41 = np.c_[keys[:n_search], (10 * sim[:n_search]).astype(np.float32)]
        q41 = np.r_[keys[c], np.float32(0)]
        l2 = np.zeros(n_search, np.float32)
        for i in range(41):
            err = (q41[i] - k41[:, i]).astype(np.float64)
            l2 = (l2.astype(np.float64) + err * err).astype(np.float32)
        top = sorted(range(n_search), key=lambda i: (l2[i], i))[:10]
        assert ci[c, :len(top)].tolist() == top and np.all(ci[c, len(top):] == -1)
        assert np.allclose(cs[c, :len(top)], sim[top], rtol=1e-12, atol=1e-15)


def _graph(n, rng):
    nodes = np.zeros((n, 7)); nodes[:, 6] = 1
    for i in range(n):
        th = 0.07 * i
        nodes[i, :3] = [i * 1.2, 0.1 * i * i, 0]
        nodes[i, 3:] = [0, 0, math.sin(th / 2), math.cos(th / 2)]
    ids, meas = [], []
    for i in range(n - 1):
        pairs = [(i, i + 1, 0)] + ([(max(0, i - 7), i + 1, 1)] if i % 3 == 2 else [])
        for (a, b, t) in pairs:
            qa, qb = nodes[a, 3:], nodes[b, 3:]
            tha, thb = 2 * math.atan2(qa[2], qa[3]), 2 * math.atan2(qb[2], qb[3])
            d = nodes[b, :3] - nodes[a, :3]
            c, s = math.cos(-tha), math.sin(-tha)
            dth = thb - tha + rng.normal(0, 0.01)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1] + rng.normal(0, 0.05), s * d[0] + c * d[1] + rng.normal(0, 0.05), 0, 0, 0, math.sin(dth / 2), math.cos(dth / 2)])
    nodes[:, :3] += rng.normal(0, 0.02, size=(n, 3))
    q = nodes[:, 3:] + rng.normal(0, 0.005, size=(n, 4))
    nodes[:, 3:] = q / np.linalg.norm(q, axis=1, keepdims=True)     # general (non-planar) unit quaternions
    return nodes, np.array(ids, np.int32), np.array(meas)


@pytest.mark.parametrize("n", [2, 30, 600])
def test_pgo_assemble(ctx, oracle, n):
    rng = np.random.default_rng(n)
    nodes, ids, meas = _graph(n, rng)
    c_ref, Hd_ref, Ho_ref, g_ref, r_ref = oracle.pgo_assemble(nodes, ids, meas)
    c, Hd, Ho, g, r = api.pgo_assemble(ctx, nodes, ids, meas)
    assert abs(c - c_ref) <= 1e-12 * abs(c_ref)
    for a, b in ((Hd, Hd_ref), (Ho, Ho_ref), (g, g_ref), (r, r_ref)):
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300)
    assert np.all(Hd[0] == 0) and np.all(g[0] == 0)


def test_pgo_full_information_matrices(ctx, oracle):
    rng = np.random.default_rng(5)
    nodes, ids, meas = _graph(20, rng)
    info = np.zeros((len(ids), 36))
    for c in range(len(ids)):
        a = rng.normal(size=(6, 6))
        info[c] = (a @ a.T + 6 * np.eye(6)).reshape(-1)
    OP, P = oracle.default_pgo_params(replace_cov_by_identity=0), api.default_pgo_params(replace_cov_by_identity=0)
    c_ref, Hd_ref, Ho_ref, g_ref, r_ref = oracle.pgo_assemble(nodes, ids, meas, OP, info=info, fixed_node=3)
    c, Hd, Ho, g, r = api.pgo_assemble(ctx, nodes, ids, meas, P, info=info, fixed_node=3)
    assert abs(c - c_ref) <= 1e-12 * abs(c_ref)
    assert np.abs(Hd - Hd_ref).max() <= 1e-11 * np.abs(Hd_ref).max() and np.abs(Ho - Ho_ref).max() <= 1e-11 * np.abs(Ho_ref).max()
    assert np.all(Hd[3] == 0)
    bad = info.copy(); bad[2] = -bad[2]
    with pytest.raises(api.TbvError):
        api.pgo_assemble(ctx, nodes, ids, meas, P, info=bad)


def _damped_system(ids, Hd, Ho, g, radius, fixed):
    """The same (H + D) delta = -g as a scipy CSC matrix without the fixed node's rows / columns (checker for tbv_pgo_solve_step)."""
    import scipy.sparse as sp
    n, r6 = len(Hd), np.arange(6)
    blk_i = lambda i: (6 * i[:, None, None] + r6[None, :, None]) + 0 * r6[None, None, :]
    blk_j = lambda j: (6 * j[:, None, None] + r6[None, None, :]) + 0 * r6[None, :, None]
    nn, a, b = np.arange(n), ids[:, 0].astype(np.int64), ids[:, 1].astype(np.int64)
    rows = np.r_[blk_i(nn).ravel(), blk_i(a).ravel(), blk_j(b).ravel()]
    cols = np.r_[blk_j(nn).ravel(), blk_j(b).ravel(), blk_i(a).ravel()]
    A = sp.coo_matrix((np.r_[Hd.ravel(), Ho.ravel(), Ho.ravel()], (rows, cols)), shape=(6 * n, 6 * n)).tocsr()
    A = A + sp.diags(np.clip(A.diagonal(), 1e-6, 1e32) / radius)
    keep = np.r_[0:6 * fixed, 6 * fixed + 6:6 * n]
    return A[keep][:, keep].tocsc(), -g.reshape(-1)[keep], keep


@pytest.mark.parametrize("n,radius,fixed", [(1, 1e4, 0), (2, 1e4, 0), (3, 1e4, 1), (7, 1e4, 3), (30, 1e4, 0), (600, 1e4, 0), (600, 1e2, 17), (600, 1e8, 17),
                                            (1025, 1e6, 1024), (4500, 1e4, 0), (4500, 1e8, 0)])
def test_pgo_solve_step_matches_sparse_direct_solve(ctx, n, radius, fixed):
    """tbv_pgo_solve_step (CG preconditioned with the odometry chain by block cyclic reduction, one 8-CTA cluster) against scipy's sparse LU
    on the same damped normal equations — including the large trust-region radii at which a block-Jacobi preconditioner does not converge
    (profiles/r2a_pgo_one_cta.json) and node counts around the powers of two of the reduction.
    Tolerance: relative residual <= 1e-11 (asked: 1e-12), |delta - direct| <= 1e-5 |direct|."""
    import scipy.sparse.linalg as spl
    rng = np.random.default_rng(n)
    nodes, ids, meas = _graph(n, rng)
    _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas, fixed_node=fixed)
    delta, iters, rel = api.pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node=fixed, radius=radius, max_iters=20000, rel_tol=1e-12)
    if n == 1:                     # only the fixed node: nothing to solve
        assert iters == 0 and np.all(delta == 0)
        return
    assert 0 < iters < 400 and rel <= 1e-12
    assert np.all(delta[fixed] == 0)
    A, b, keep = _damped_system(ids, Hd, Ho, g, radius, fixed)
    ref = spl.spsolve(A, b)
    got = delta.reshape(-1)[keep]
    assert np.linalg.norm(A @ got - b) <= 1e-11 * np.linalg.norm(b)
    assert np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref)
    # deterministic: the same call twice gives the same bits
    delta2, iters2, _ = api.pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node=fixed, radius=radius, max_iters=20000, rel_tol=1e-12)
    assert iters2 == iters and np.array_equal(delta, delta2)


def test_pgo_solve_step_iteration_cap_and_arguments(ctx):
    rng = np.random.default_rng(3)
    nodes, ids, meas = _graph(200, rng)
    _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas)
    d5, it5, rel5 = api.pgo_solve_step(ctx, ids, Hd, Ho, g, max_iters=1)
    assert it5 == 1 and 0 < rel5 < 1.0 and np.all(np.isfinite(d5))
    d0, it0, rel0 = api.pgo_solve_step(ctx, ids, Hd, Ho, g, max_iters=0)
    assert it0 == 0 and np.all(d0 == 0)
    dz, itz, relz = api.pgo_solve_step(ctx, ids, Hd, Ho, np.zeros_like(g))        # zero gradient: nothing to do
    assert itz == 0 and relz == 0 and np.all(dz == 0)
    bad = ids.copy(); bad[4, 1] = 200
    with pytest.raises(api.TbvError):
        api.pgo_solve_step(ctx, bad, Hd, Ho, g)
    with pytest.raises(api.TbvError):
        api.pgo_solve_step(ctx, ids, Hd, Ho, g, radius=0.0)


def test_pgo_plus_is_the_quaternion_left_update():
    rng = np.random.default_rng(0)
    q = rng.normal(size=(5, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    nodes = np.c_[rng.normal(size=(5, 3)), q]
    d = rng.normal(0, 0.3, size=(5, 6)); d[2] = 0
    out = api.pgo_plus(nodes, d)
    assert np.allclose(out[:, :3], nodes[:, :3] + d[:, :3]) and np.allclose(np.linalg.norm(out[:, 3:], axis=1), 1.0)
    assert np.array_equal(out[2], nodes[2])
    for i in range(5):       # rotation matrices: R(out) = R(exp(d)) R(q)
        def rot(qq):
            x, y, z, w = qq
            return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                             [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                             [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        th = np.linalg.norm(d[i, 3:])
        dq = np.r_[np.sin(th) * d[i, 3:] / th, np.cos(th)] if th > 0 else np.array([0, 0, 0, 1.0])
        assert np.allclose(rot(out[i, 3:]), rot(dq) @ rot(nodes[i, 3:]), atol=1e-12)


def test_pgo_optimize_recovers_a_consistent_graph(ctx, oracle):
    """Noise-free measurements, perturbed nodes: the optimum is the ground truth (cost 0) up to the fixed node's gauge."""
    n = 120
    truth = np.zeros((n, 7)); truth[:, 6] = 1
    for i in range(n):
        th = 0.05 * i
        truth[i, :3] = [2.0 * math.cos(th) * i / 4, 2.0 * math.sin(th) * i / 4, 0]
        truth[i, 3:] = [0, 0, math.sin(th / 2), math.cos(th / 2)]
    ids, meas = [], []
    for i in range(n - 1):
        for (a, b, t) in [(i, i + 1, 0)] + ([(max(0, i - 40), i + 1, 1)] if i % 5 == 4 else []):
            tha, thb = 0.05 * a, 0.05 * b
            d = truth[b, :3] - truth[a, :3]
            c, s = math.cos(-tha), math.sin(-tha)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1], s * d[0] + c * d[1], 0, 0, 0, math.sin((thb - tha) / 2), math.cos((thb - tha) / 2)])
    ids, meas = np.array(ids, np.int32), np.array(meas)
    rng = np.random.default_rng(9)
    start = truth.copy()
    start[1:, :2] += rng.normal(0, 0.3, size=(n - 1, 2))
    start[1:] = api.pgo_plus(start[1:], np.c_[np.zeros((n - 1, 5)), rng.normal(0, 0.05, size=n - 1)])
    x, S = api.pgo_optimize(ctx, start, ids, meas, function_tolerance=1e-16, gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    assert S.initial_cost > 1.0 and S.final_cost <= 1e-14 * S.initial_cost and S.successful_steps >= 3
    sign = np.sign(np.sum(x[:, 3:] * truth[:, 3:], axis=1))[:, None]
    assert np.abs(x[:, :3] - truth[:, :3]).max() <= 1e-6 and np.abs(sign * x[:, 3:] - truth[:, 3:]).max() <= 1e-7
    assert np.array_equal(x[0], start[0])            # the fixed node does not move


def _ring_graph(n, rng):
    """A spiral driven anticlockwise (radius 30 m, loops to the scan one lap earlier), noisy odometry, start = dead reckoning."""
    truth = np.zeros((n, 7)); truth[:, 6] = 1
    for i in range(n):
        th = 0.05 * i
        truth[i, :3] = [30 * math.cos(th) * (1 + 0.002 * i), 30 * math.sin(th) * (1 + 0.002 * i), 0]
        truth[i, 3:] = [0, 0, math.sin((th + math.pi / 2) / 2), math.cos((th + math.pi / 2) / 2)]
    ids, meas = [], []
    for i in range(n - 1):
        for (a, b, t) in [(i, i + 1, 0)] + ([(i + 1 - 125, i + 1, 1)] if (i % 5 == 4 and i + 1 >= 125) else []):
            tha, thb = 0.05 * a + math.pi / 2, 0.05 * b + math.pi / 2
            d = truth[b, :3] - truth[a, :3]
            c, s = math.cos(-tha), math.sin(-tha)
            dth = thb - tha + rng.normal(0, 0.002)
            ids.append((a, b, t))
            meas.append([c * d[0] - s * d[1] + rng.normal(0, 0.02), s * d[0] + c * d[1] + rng.normal(0, 0.02), 0, 0, 0, math.sin(dth / 2), math.cos(dth / 2)])
    ids, meas = np.array(ids, np.int32), np.array(meas)
    start = truth.copy()
    for c, (a, b, t) in enumerate(ids):
        if t == 0:
            tha, dth = 2 * math.atan2(start[a, 5], start[a, 6]), 2 * math.atan2(meas[c, 5], meas[c, 6])
            cc, ss = math.cos(tha), math.sin(tha)
            start[b, :2] = [start[a, 0] + cc * meas[c, 0] - ss * meas[c, 1], start[a, 1] + ss * meas[c, 0] + cc * meas[c, 1]]
            start[b, 3:] = [0, 0, math.sin((tha + dth) / 2), math.cos((tha + dth) / 2)]
    return truth, start, ids, meas


@pytest.mark.parametrize("loop_scaling", [500000.0, 1.0])
def test_pgo_optimize_reaches_a_stationary_point_of_the_oracle_cost(ctx, oracle, loop_scaling):
    """Noisy graph with Cauchy-robustified loop constraints (TBV's loop covariance scaling, and loops at full weight): at the returned nodes the
    ORACLE's gradient vanishes (<= 1e-6 of the start's; 1e-4 in the flat valley of the full-weight case) and its cost equals the summary's; with Ceres' default tolerances the run stops on
    function_tolerance no later."""
    rng = np.random.default_rng(2)
    truth, start, ids, meas = _ring_graph(300, rng)
    P, OP = api.default_pgo_params(loop_scaling=loop_scaling), oracle.default_pgo_params(loop_scaling=loop_scaling)
    x, S = api.pgo_optimize(ctx, start, ids, meas, P, function_tolerance=1e-15, gradient_tolerance=1e-9, parameter_tolerance=1e-14)
    c_ref, _, _, g_ref, _ = oracle.pgo_assemble(x, ids, meas, OP)
    c0, _, _, g0, _ = oracle.pgo_assemble(start, ids, meas, OP)
    if loop_scaling == 1.0:      # saturated Cauchy loops: a flat valley, the run ends on function_tolerance = 1e-15 within the 200 iterations
        assert S.termination in ("gradient_tolerance", "function_tolerance", "parameter_tolerance") and S.iterations <= 200
    else:
        assert S.termination == "gradient_tolerance" and S.iterations < 100
    assert abs(S.final_cost - c_ref) <= 1e-10 * c_ref and abs(S.initial_cost - c0) <= 1e-10 * c0 and c_ref < c0
    assert np.abs(g_ref).max() <= (1e-4 if loop_scaling == 1.0 else 1e-6) * np.abs(g0).max()
    assert np.allclose(np.linalg.norm(x[:, 3:], axis=1), 1.0, atol=1e-12)
    if loop_scaling == 1.0:      # loops at full weight pull the dead-reckoning drift in
        assert np.abs(x[:, :2] - truth[:, :2]).max() < 0.5 * np.abs(start[:, :2] - truth[:, :2]).max()
    _, S2 = api.pgo_optimize(ctx, start, ids, meas, P)                      # Ceres defaults
    assert S2.termination in ("function_tolerance", "gradient_tolerance") and S2.iterations <= S.iterations
    assert S2.final_cost <= S.final_cost * (1 + 1e-4)


def test_pgo_solve_damped_takes_an_arbitrary_damping(ctx):
    """tbv_pgo_solve_damped: (H + diag(damping)) delta = -g for ANY positive damping vector (what LevenbergMarquardtStrategy needs: its
    diagonal is scaled, clamped and reused), against a dense direct solve."""
    rng = np.random.default_rng(3)
    nodes, ids, meas = _graph(40, rng)
    _, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas)
    n = len(Hd)
    H = np.zeros((6 * n, 6 * n))
    for i in range(n):
        H[6 * i:6 * i + 6, 6 * i:6 * i + 6] = Hd[i]
    for c, (a, b, _t) in enumerate(ids):
        H[6 * a:6 * a + 6, 6 * b:6 * b + 6] += Ho[c]
        H[6 * b:6 * b + 6, 6 * a:6 * a + 6] += Ho[c].T
    keep = np.arange(6, 6 * n)
    for damping in (rng.uniform(0.1, 50.0, size=(n, 6)), np.full((n, 6), 1e-9), rng.uniform(1e-3, 1.0, size=(n, 6)) * Hd[:, np.arange(6), np.arange(6)].clip(1e-3)):
        delta, iters, rel = api.pgo_solve_damped(ctx, ids, Hd, Ho, g, damping, fixed_node=0)
        A = (H + np.diag(damping.reshape(-1)))[np.ix_(keep, keep)]
        want = np.linalg.solve(A, -g.reshape(-1)[keep])
        assert iters > 0 and np.all(delta[0] == 0) and np.allclose(delta.reshape(-1)[keep], want, rtol=1e-6, atol=1e-9 * np.abs(want).max())
    with pytest.raises(ValueError):
        api.pgo_solve_damped(ctx, ids, Hd, Ho, g, np.zeros((n, 6)))


@pytest.mark.parametrize("loop_scaling", [500000.0, 1.0])
def test_pgo_optimize_on_the_device_equals_the_host_driven_loop(ctx, oracle, loop_scaling):
    """tbv_pgo_optimize (every LM iteration on the device) against api.pgo_optimize_ceres (the same Ceres 2.1.0 iteration rules driven from the
    host, one device call at a time): same iterations, termination and accepted steps; nodes to 1e-9; and the oracle's gradient vanishes there."""
    rng = np.random.default_rng(2)
    truth, start, ids, meas = _ring_graph(300, rng)
    P, OP = api.default_pgo_params(loop_scaling=loop_scaling), oracle.default_pgo_params(loop_scaling=loop_scaling)
    kw = dict(function_tolerance=1e-14, gradient_tolerance=1e-9, parameter_tolerance=1e-14) if loop_scaling != 1.0 else {}
    xh, Sh = api.pgo_optimize_ceres(ctx, start, ids, meas, P, **kw)
    xd, Sd = api.pgo_optimize_device(ctx, start, ids, meas, P, **kw)
    assert (Sd.iterations, Sd.successful_steps, Sd.termination) == (Sh.iterations, Sh.successful_steps, Sh.termination)
    assert abs(Sd.final_cost - Sh.final_cost) <= 1e-9 * Sh.final_cost and abs(Sd.initial_cost - Sh.initial_cost) <= 1e-12 * Sh.initial_cost
    assert np.abs(xd - xh).max() <= 1e-9 and np.array_equal(xd[0], start[0])
    c_ref, _, _, g_ref, _ = oracle.pgo_assemble(xd, ids, meas, OP)
    c0, _, _, g0, _ = oracle.pgo_assemble(start, ids, meas, OP)
    assert abs(Sd.final_cost - c_ref) <= 1e-10 * c_ref and c_ref < c0
    if loop_scaling != 1.0:
        assert np.abs(g_ref).max() <= 1e-6 * np.abs(g0).max()
    assert Sd.device_ms > 0


def test_pgo_optimize_full_sequence_graph_beats_the_reference_time(ctx, oracle):
    """SURVEY 6 / VERDICT r1: the reference's CeresLeastSquares needs 1.23 s for an Oxford sequence graph (4.5 k nodes, 5.2 k constraints) on one
    CPU thread.  A same-size graph (noisy odometry chain + a loop every 5th node one lap back: 875 loop constraints, TBV's loop weighting,
    Ceres' default tolerances) is optimised on the device, whole LM loop included, in less than half of that (measured on a B200: 0.55 s for
    48 LM iterations / 4 285 chain-preconditioned CG iterations of ~0.13 ms; profiles/r2c_pgo_bench.json has a 153-iteration graph at 0.63 s),
    and the run must stop on a Ceres tolerance at a point where the oracle's cost equals the summary's.  The time bound is stated per unit of
    work (LM iterations and CG iterations) so that it does not depend on how hard this particular graph is."""
    rng = np.random.default_rng(5)
    truth, start, ids, meas = _ring_graph(4500, rng)
    assert len(ids) > 5200
    x, S = api.pgo_optimize_device(ctx, start, ids, meas)            # warm-up (pool allocations)
    x, S = api.pgo_optimize_device(ctx, start, ids, meas)
    assert S.termination in ("function_tolerance", "gradient_tolerance", "parameter_tolerance") and S.final_cost < S.initial_cost
    c_ref = oracle.pgo_assemble(x, ids, meas)[0]
    assert abs(S.final_cost - c_ref) <= 1e-9 * c_ref
    assert S.device_ms < 615.0 and S.device_ms < 2.0 * S.iterations + 0.2 * S.cg_iterations, S
