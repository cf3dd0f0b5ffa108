"""CorAl alignment quality: CUDA CorAlRadarQuality, batched over pairs, vs the oracle's restatement of
coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:8-229, through the C-ABI.

Bar: the per-point decisions (enough neighbours in the other cloud, more than two in the own one, finite entropies) and therefore
count_valid / overlap / valid are exact (float radius test = index work); the entropies are logs of 2x2 covariance determinants
summed in a different order than the reference's Eigen products.  For a nearly singular neighbourhood (collinear peaks) the entropy
1/2 log(2 pi e det + 1e-8) is conditioned by the 1e-8 floor: |dH| <= 1/2 * 2 pi e * |d det| / 1e-8 ~ 1e9 * |d det|, and det carries a few
ulp of c00 * c11 -> 1e-7 absolute per point, 1e-8 on the means (the means are over >= 100 points, mostly well conditioned)."""
import numpy as np
import pytest

from tbv_slam_public_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clouds(oracle, stream8):
    out = {"peaks": [], "dense": []}
    for i in range(6):
        r = oracle.kstrongest(stream8.scans[i], z_min=60.0, k=40)
        az, rg, I, x, y = r["peaks"]
        out["peaks"].append((x, y, I.astype(np.float32)))
        az, rg, I, x, y = oracle.kstrongest(stream8.scans[i], z_min=70.0, k=12)["filtered"]
        out["dense"].append((x, y, I.astype(np.float32)))
    return out


def _check(ctx, oracle, cl, pairs, T_src, T_ref, T_off, tol_point=1e-7, tol_mean=1e-8, **kw):
    res, pp = api.CorAlRadarQuality(ctx, cl, [p[0] for p in pairs], [p[1] for p in pairs], T_src, T_ref, T_off, per_point=True, **kw)
    row = 0
    for k, (s, r) in enumerate(pairs):
        ref = oracle.coral_quality(cl[s], cl[r], T_src[k], T_ref[k], Toffset=T_off[k] if T_off is not None else (0, 0, 0), per_point=True, **kw)
        g = res[k]
        m = ref["merged_size"]
        got_pp = pp[row:row + m]
        row += m
        assert g.merged_size == m
        assert np.array_equal(got_pp[:, 2], ref["per_point"][:, 2]), f"pair {k}: per-point validity differs"
        assert np.abs(got_pp[:, :2] - ref["per_point"][:, :2]).max() < tol_point, f"pair {k}: per-point entropies differ"
        assert g.count_valid == ref["count_valid"] and g.valid == int(ref["valid"])
        assert g.overlap == ref["overlap"]
        assert abs(g.joint - ref["joint"]) < tol_mean and abs(g.sep - ref["sep"]) < tol_mean
    return res


@pytest.mark.parametrize("kind", ["peaks", "dense"])
def test_coral_batch_vs_oracle(ctx, oracle, stream8, clouds, kind):
    cl = clouds[kind]
    gt = stream8.gt
    pairs = [(1, 0), (2, 1), (3, 1), (5, 4), (4, 4), (0, 5)]
    T_src = np.array([gt[s] for s, _ in pairs])
    T_ref = np.array([gt[r] for _, r in pairs])
    rng = np.random.default_rng(3)
    T_off = np.zeros((len(pairs), 3))
    T_off[2] = (1.0, -0.6, 0.04)          # a misaligned candidate
    T_off[3] = rng.normal(0, 0.2, 3) * (1, 1, 0.1)
    res = _check(ctx, oracle, cl, pairs, T_src, T_ref, T_off)
    if kind == "dense":
        assert all(r.count_valid > 100 for r in res)
        # a misaligned pair has a clearly larger joint entropy than the same scans at their true poses
        aligned = _check(ctx, oracle, cl, [(3, 1)], T_src[2:3], T_ref[2:3], None)[0]
        assert res[2].joint > aligned.joint + 0.1


def test_coral_intensity_weights_and_radius(ctx, oracle, stream8, clouds):
    cl = clouds["dense"]
    gt = stream8.gt
    pairs = [(1, 0), (2, 0)]
    T_src = np.array([gt[1], gt[2]])
    T_ref = np.array([gt[0], gt[0]])
    _check(ctx, oracle, cl, pairs, T_src, T_ref, None, weight_res_intensity=True)
    # radius 3: moments of up to 9 m^2 per point -> |d det| (and with it the floor-conditioned entropy error) grows ~ 10x
    _check(ctx, oracle, cl, pairs, T_src, T_ref, None, tol_point=2e-6, tol_mean=1e-7, radius=3.0)


def test_coral_degenerate_inputs(ctx, oracle):
    """Clouds far apart (no overlap: nothing valid), collinear points (zero determinant -> entropy of the 1e-8 floor), tiny clouds."""
    line = (np.linspace(0, 5, 60, dtype=np.float32), np.zeros(60, np.float32), np.full(60, 90, np.float32))
    far = (line[0] + 500.0, line[1], line[2])
    tiny = (np.array([0.1, 0.2], np.float32), np.array([0.0, 0.1], np.float32), np.array([80, 81], np.float32))
    cl = [line, far, tiny]
    T = np.zeros((3, 3))
    res = _check(ctx, oracle, cl, [(0, 1), (0, 0), (2, 0)], T, T, None)
    assert res[0].count_valid == 0 and res[0].valid == 0
    assert res[1].count_valid == 120 and abs(res[1].joint - 0.5 * np.log(1e-8)) < 1e-6


def test_coral_capacity_is_reported(ctx):
    big = (np.zeros(5000, np.float32), np.zeros(5000, np.float32), np.zeros(5000, np.float32))
    with pytest.raises(api.TbvError):
        api.CorAlRadarQuality(ctx, [big], [0], [0], np.zeros((1, 3)), np.zeros((1, 3)))


def test_verify_candidates_composes_both_batched_kernels(ctx, oracle, stream8, clouds):
    """verification.verify_candidates: CorAl + CFEAR features of all candidates from two launches -> alignment score -> probability."""
    from tbv_slam_public_b200 import verification as V
    cells = []
    for i in range(4):
        az, rg, I, x, y = oracle.kstrongest(stream8.scans[i])["filtered"]
        cells.append(oracle.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)[0])
    gt = stream8.gt
    src, ref = [1, 2, 3], [0, 0, 1]
    Ts, Tr = np.array([gt[s] for s in src]), np.array([gt[r] for r in ref])
    clf = V.LogisticRegression(-8.42595, [-15.2287, 7.47573, -0.0680198, -1.74182, 0.0945444, 0.022217])
    p, X, quality = V.verify_candidates(ctx, clouds["peaks"][:4], cells, src, ref, Ts, Tr, sc_sim=[0.1, 0.2, 0.1], odom_bounds=[0.0, 0.0, 0.3],
                                        alignment_classifier=clf)
    assert X.shape == (3, 6) and p.shape == (3,) and np.all((p > 0) & (p < 1))
    for k in range(3):
        c = oracle.coral_quality(clouds["peaks"][src[k]], clouds["peaks"][ref[k]], Ts[k], Tr[k])
        n, score, cost, res = oracle.get_cost([cells[ref[k]], cells[src[k]]], [Tr[k], Ts[k]],
                                              oracle.default_reg_params(cost=oracle.P2L, loss=oracle.HUBER, loss_limit=0.3, weight_opt=oracle.W_UNIFORM), itr=0)
        want = np.array([c["joint"], c["sep"], c["overlap"], cost, n, (len(cells[src[k]]) + len(cells[ref[k]])) / 2.0])
        assert np.allclose(X[k], want, rtol=1e-9, atol=1e-8)
        assert abs(quality[k] - (want @ clf.coef_ + clf.intercept_)) < 1e-6
