"""K4/K5 parity: correspondences + cost / J^T J / J^T r, GetCost, Register and batched loop registrations vs the oracle.

Bar: associations bit-exact (index work); cost / gradient / Hessian within 1e-12 relative (the reduction order differs:
the reference sums sequentially, the kernel in a fixed tree); registered poses within 1e-5 m / 1e-6 rad (north_star),
with the same number of association and LM iterations.
"""
import numpy as np
import pytest

from tbv_slam_public_b200 import api

pytestmark = pytest.mark.gpu

POS_TOL, ANG_TOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def cellsets(oracle, stream8):
    sets = []
    for i in range(6):
        az, rg, I, x, y = oracle.kstrongest(stream8.scans[i])["filtered"]
        c, _ = oracle.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)
        sets.append(c)
    return sets


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _ang(a, b):
    d = a - b
    return abs(np.arctan2(np.sin(d), np.cos(d)))


@pytest.mark.parametrize("cost", [api.P2L, api.P2P, api.P2D])
@pytest.mark.parametrize("loss,wopt", [(api.HUBER, api.W_COMBINED), (api.CAUCHY, api.W_UNIFORM), (api.LOSS_NONE, api.W_SIM_N),
                                       (api.TUKEY, api.W_SIM_DIR), (api.SOFTLONE, api.W_SIM_SCALE), (api.COMBINED, api.W_UNIFORM)])
def test_pair_normal_eq(ctx, oracle, cellsets, cost, loss, wopt):
    tgt, src = cellsets[0], cellsets[1]
    Tt, Ts = (0.0, 0.0, 0.0), (2.4, 0.05, 0.003)
    for itr in (1, 2):
        kw = dict(cost=cost, loss=loss, weight_opt=wopt, loss_limit=0.1, cov_scale=1.0, regularization=0.1)
        ref = oracle.pair_normal_eq(tgt, Tt, src, Ts, oracle.default_reg_params(**kw), itr=itr)
        got = ctx.pair_normal_eq(tgt, Tt, src, Ts, api.default_reg_params(**kw), itr=itr)
        assert np.array_equal(ref["assoc"], got["assoc"])
        assert ref["n_res"] == got["n_res"] and ref["n_res"] > 50
        assert abs(ref["cost"] - got["cost"]) <= 1e-12 * abs(ref["cost"])
        assert _rel(got["g"], ref["g"]) < 1e-11
        assert _rel(got["H"], ref["H"]) < 1e-12


def test_get_cost(ctx, oracle, cellsets):
    scans = [cellsets[0], cellsets[1], cellsets[2]]
    T = [(0, 0, 0), (2.5, 0, 0), (5.05, -0.04, -0.006)]
    for itr in (0, 1):
        kw = dict(cost=api.P2L, loss=api.HUBER, loss_limit=0.3)
        n, score, cost, res = oracle.get_cost(scans, T, oracle.default_reg_params(**kw), itr=itr)
        gn, gscore, gcost, gres = ctx.GetCost(scans, T, api.default_reg_params(**kw), itr=itr)
        assert n == gn and n > 100
        assert abs(cost - gcost) <= 1e-12 * cost and abs(score - gscore) <= 1e-12 * score
        assert np.allclose(res, gres, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("cost,loss,wopt", [(api.P2L, api.HUBER, api.W_COMBINED), (api.P2P, api.HUBER, api.W_COMBINED),
                                            (api.P2D, api.CAUCHY, api.W_COMBINED), (api.P2L, api.CAUCHY, api.W_UNIFORM)])
def test_register_multi_keyframe(ctx, oracle, stream8, cellsets, cost, loss, wopt):
    """scan 4 against keyframes 0..3 placed at the ground-truth poses, perturbed initial guess."""
    gt = stream8.gt
    T = [tuple(gt[i]) for i in range(4)] + [(gt[4][0] + 0.4, gt[4][1] - 0.3, gt[4][2] + 0.02)]
    scans = cellsets[:5]
    kw = dict(cost=cost, loss=loss, weight_opt=wopt, loss_limit=0.1, cov_scale=1.0, regularization=0.1)
    Tr, sr = oracle.register(scans, T, oracle.default_reg_params(**kw))
    Tg, sg = ctx.Register(scans, T, api.default_reg_params(**kw))
    assert sr.success == 1 and sg.success == 1
    assert (sr.itrs, sr.lm_iterations, sr.num_residuals, sr.last_n_iterations, sr.termination) == \
           (sg.itrs, sg.lm_iterations, sg.num_residuals, sg.last_n_iterations, sg.termination)
    assert np.abs(Tr[-1, :2] - Tg[-1, :2]).max() < POS_TOL and _ang(Tr[-1, 2], Tg[-1, 2]) < ANG_TOL
    assert abs(sr.score - sg.score) <= 1e-9 * abs(sr.score)
    # the estimate is near the ground truth (sanity of the whole chain, not a parity statement)
    assert np.abs(Tg[-1, :2] - gt[4][:2]).max() < 0.5


def test_register_failure_and_empty(ctx, oracle, cellsets):
    far = [(0, 0, 0), (500.0, 500.0, 1.0)]  # no correspondences -> BuildOptimizationProblem fails
    Tr, sr = oracle.register(cellsets[:2], far)
    Tg, sg = ctx.Register(cellsets[:2], far)
    assert sr.success == 0 and sg.success == 0 and sr.itrs == sg.itrs
    assert np.allclose(Tr, Tg, atol=1e-12)
    empty = np.zeros((0, 16))
    Tg, sg = ctx.Register([cellsets[0], empty], [(0, 0, 0), (1, 0, 0)])
    assert sg.success == 0


def test_register_batch_loop_candidates(ctx, oracle, stream8, cellsets):
    """loopclosure::Register for a batch of (from, to) pairs with initial errors of the Scan-Context scale."""
    rng = np.random.default_rng(3)
    gt = stream8.gt
    fs, ts, Tf, Tt = [], [], [], []
    for _ in range(24):
        a, b = rng.choice(6, 2, replace=False)
        err = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(-0.1, 0.1)])
        fs.append(a); ts.append(b); Tf.append(gt[a] + err); Tt.append(gt[b])
    Tr, Ta, summ = ctx.RegisterBatch(cellsets, fs, ts, Tf, Tt)
    n_ok = 0
    for p in range(24):
        ok, Ta_ref, Tr_ref, itrs, score = oracle.loop_register(cellsets[fs[p]], cellsets[ts[p]], Tf[p], Tt[p])
        assert bool(summ[p].success) == ok and summ[p].itrs == itrs
        assert np.abs(Tr_ref[:2] - Tr[p, :2]).max() < POS_TOL and _ang(Tr_ref[2], Tr[p, 2]) < ANG_TOL
        if ok:
            n_ok += 1
            assert np.abs(Ta_ref[:2] - Ta[p, :2]).max() < POS_TOL and _ang(Ta_ref[2], Ta[p, 2]) < ANG_TOL
            assert abs(score - summ[p].score) <= 1e-9 * abs(score)
    assert n_ok >= 20


def _np_cov_from_samples(samples, score_scale, scaler):
    """odometrykeyframefuser.cpp:315-378 restated with numpy (lstsq = minimum-norm least squares, like bdcSvd().solve())."""
    x, y, z, c = samples.T
    A = np.stack([x * x, y * y, z * z, x * y, y * z, z * x, x, y, z, np.ones_like(x)], axis=1)
    # the columns span 12 orders of magnitude: condition the solve like the library does (column scaling leaves the minimiser unchanged)
    s = np.linalg.norm(A, axis=0)
    q = np.linalg.lstsq(A / s, c, rcond=None)[0] / s
    H = np.array([[2 * q[0], q[3], q[5]], [q[3], 2 * q[1], q[4]], [q[5], q[4], 2 * q[2]]])
    if not (np.linalg.eigvalsh(H) > 0).all():
        return False, None
    C3 = 2.0 * np.linalg.inv(H) * score_scale * scaler
    cov = np.eye(6)
    cov[:2, :2] = C3[:2, :2]
    cov[5, 5] = C3[2, 2]
    cov[0, 5], cov[1, 5], cov[5, 0], cov[5, 1] = C3[0, 2], C3[1, 2], C3[2, 0], C3[2, 1]
    return True, cov


@pytest.mark.parametrize("n_axis,cost", [(3, api.P2L), (5, api.P2L), (3, api.P2P)])
def test_covariance_by_cost_sampling(ctx, oracle, cellsets, n_axis, cost):
    """approximateCovarianceBySampling (odometrykeyframefuser.cpp:261-380): the n^3 GetCost samples come from ONE launch and must
    equal the oracle's n^3 sequential GetCost calls; the fitted covariance must equal a numpy restatement of the fit."""
    scans = [cellsets[0], cellsets[1], cellsets[2], cellsets[3]]
    T = np.array([(0, 0, 0), (2.5, 0, 0), (5.0, 0.0, 0.0), (7.4, 0.03, 0.002)], float)
    kw = dict(cost=cost, loss=api.HUBER, loss_limit=0.1, weight_opt=api.W_COMBINED)
    # register first, as processFrame does; the sampling runs around the registered pose with the object's itr_ (> 1)
    Treg, summary = ctx.Register(scans, T.copy(), api.default_reg_params(**kw))
    ref = oracle.cost_samples(scans, Treg, oracle.default_reg_params(**kw), itr=2, n_per_axis=n_axis)
    score_scale = summary.final_cost / (summary.num_residuals - 3)
    ok, cov, got = ctx.approximateCovarianceBySampling(scans, Treg, score_scale, api.default_reg_params(**kw), itr=2, samples_per_axis=n_axis)
    assert got.shape == ref.shape == (n_axis ** 3, 4)
    assert np.array_equal(got[:, :3], ref[:, :3])                       # the sampling grid, in the reference's order
    assert np.abs(got[:, 3] - ref[:, 3]).max() <= 1e-11 * np.abs(ref[:, 3]).max()
    assert ref[:, 3].min() > 0 and np.ptp(ref[:, 3]) > 0
    ok_np, cov_np = _np_cov_from_samples(ref, score_scale, 4.0)
    assert ok == ok_np
    if ok:
        assert np.allclose(cov, cov_np, rtol=1e-6, atol=0)
        assert np.allclose(cov, cov.T) and (np.linalg.eigvalsh(cov[np.ix_([0, 1, 5], [0, 1, 5])]) > 0).all()


def test_cov_from_cost_samples_recovers_a_known_quadric(ctx):
    """Samples of f = 1/2 d^T H d + g^T d + c on the 3x3x3 grid give back 2 H^-1 * scale exactly; a saddle is reported not convex."""
    import ctypes as C
    H = np.array([[40.0, 3.0, 200.0], [3.0, 25.0, -150.0], [200.0, -150.0, 9.0e5]])
    g = np.array([0.3, -0.2, 5.0])
    xy = np.linspace(-0.2, 0.2, 3)
    th = np.linspace(-0.00218125, 0.00218125, 3)
    S = np.array([(x, y, t, 0.0) for t in th for x in xy for y in xy])
    d = S[:, :3]
    S[:, 3] = 0.5 * np.einsum("ni,ij,nj->n", d, H, d) + d @ g + 7.0
    cov = np.zeros((6, 6)); ok = C.c_int(0)
    L = api.lib()
    assert L.tbv_cov_from_cost_samples(S.ctypes.data_as(C.c_void_p), len(S), C.c_double(0.5), C.c_double(4.0), cov.ctypes.data_as(C.c_void_p), C.byref(ok)) == 0
    assert ok.value == 1
    want = 2.0 * np.linalg.inv(H) * 0.5 * 4.0
    assert np.allclose(cov[np.ix_([0, 1, 5], [0, 1, 5])], want, rtol=1e-7)
    S[:, 3] = 0.5 * np.einsum("ni,ij,nj->n", d, np.diag([40.0, -25.0, 9e5]), d)
    assert L.tbv_cov_from_cost_samples(S.ctypes.data_as(C.c_void_p), len(S), C.c_double(0.5), C.c_double(4.0), cov.ctypes.data_as(C.c_void_p), C.byref(ok)) == 0
    assert ok.value == 0


def test_cfear_quality_batch(ctx, oracle, cellsets):
    """CFEARQuality (AlignmentQuality.cpp:330-352) for a batch of pairs in one launch == one oracle GetCost per pair (P2L, Huber 0.3,
    uniform weights, itr_ = 0): score within 1e-12 relative, residual counts exact; a pair too far apart to match gives {0, 0, 0}."""
    pairs = [(1, 0), (2, 1), (3, 2), (5, 0), (4, 4)]
    T = {0: (0, 0, 0), 1: (2.5, 0, 0), 2: (5.0, 0.02, 0.001), 3: (7.5, 0.0, 0.0), 4: (10.0, 0, 0), 5: (900.0, 0, 0)}
    Ts = np.array([T[s] for s, _ in pairs], float); Tr = np.array([T[r] for _, r in pairs], float)
    To = np.zeros((len(pairs), 3)); To[1] = (0.4, -0.3, 0.02)
    q = ctx.CFEARQualityBatch(cellsets, [p[0] for p in pairs], [p[1] for p in pairs], Ts, Tr, To)
    kw = dict(cost=api.P2L, loss=api.HUBER, loss_limit=0.3, weight_opt=api.W_UNIFORM)
    for k, (s, r) in enumerate(pairs):
        c, sn = np.cos(Ts[k, 2]), np.sin(Ts[k, 2])
        A = np.array([[c, -sn, Ts[k, 0]], [sn, c, Ts[k, 1]], [0, 0, 1]])
        co, so = np.cos(To[k, 2]), np.sin(To[k, 2])
        B = np.array([[co, -so, To[k, 0]], [so, co, To[k, 1]], [0, 0, 1]])
        M = A @ B
        pose = (M[0, 2], M[1, 2], np.arctan2(M[1, 0], M[1, 1]))
        n, score, cost, res = oracle.get_cost([cellsets[r], cellsets[s]], [Tr[k], pose], oracle.default_reg_params(**kw), itr=0)
        if n <= 1:
            assert np.array_equal(q[k], [0, 0, 0])
        else:
            assert q[k, 1] == n and q[k, 2] == (len(cellsets[s]) + len(cellsets[r])) / 2.0
            assert abs(q[k, 0] - cost) <= 1e-12 * abs(cost)
    assert q[0, 1] > 100 and np.array_equal(q[3], [0, 0, 0])
