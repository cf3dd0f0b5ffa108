"""K1/K2 parity: CUDA k-strongest (+ peaks, + Cartesian cloud) and compensation vs the oracle, through the C-ABI.

Bar: bit-exact (azimuth, range, intensity) lists in the reference's order, bit-exact float x,y for the filter;
compensation within 1 float ulp (device atan2/sin/cos are not glibc's), mismatches counted.
"""
import numpy as np
import pytest

from tbv_slam_public_b200 import synth

pytestmark = pytest.mark.gpu


def _assert_same(o, g, what):
    names = ["azimuth", "range", "intensity", "x", "y"]
    assert len(o[0]) == len(g[0]), f"{what}: {len(o[0])} oracle points vs {len(g[0])} gpu"
    for a, b, n in zip(o, g, names):
        if a.dtype.kind == "f":
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{what}: {n} differs (bitwise)"
        else:
            assert np.array_equal(a, b), f"{what}: {n} differs"


def _check_scan(ctx, oracle, img, **kw):
    okw = dict(kw)
    if "k_strongest" in okw:
        okw["k"] = okw.pop("k_strongest")
    ref = oracle.kstrongest(img, **okw)
    f, p = ctx.StructuredKStrongest(img, **kw)
    _assert_same(ref["filtered"], f.scan(0), "filtered")
    _assert_same(ref["peaks"], p.scan(0), "peaks")
    return len(ref["filtered"][0]), len(ref["peaks"][0])


@pytest.mark.parametrize("k,z", [(12, 70.0), (40, 60.0), (1, 60.0), (128, 30.0)])
def test_radar_like_scans(ctx, oracle, stream8, k, z):
    for i in range(3):
        n, npk = _check_scan(ctx, oracle, stream8.scans[i], z_min=z, k_strongest=k)
        assert n > 0 and npk > 0


@pytest.mark.parametrize("kind", ["uniform", "equal", "zeros", "ramp", "sparse"])
@pytest.mark.parametrize("k,z", [(40, 60.0), (12, 0.0), (40, 255.0), (128, 1.0)])
def test_stress_distributions(ctx, oracle, kind, k, z):
    img = synth.stress_image(kind, seed=3)
    _check_scan(ctx, oracle, img, z_min=z, k_strongest=k)


@pytest.mark.parametrize("shape", [(400, 3360), (7, 64), (3, 5), (16, 4096), (5, 8191), (33, 1000), (2, 17)])
def test_shapes_and_alignment(ctx, oracle, shape):
    rng = np.random.default_rng(shape[1])
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    img[rng.random(shape) < 0.9] = 10  # mostly below threshold, ties at 10 when z_min is low
    _check_scan(ctx, oracle, img, z_min=60.0, k_strongest=12, min_distance=0.1)
    _check_scan(ctx, oracle, img, z_min=5.0, k_strongest=12, min_distance=0.1)


@pytest.mark.parametrize("shape", [(3, 5), (4, 17), (6, 64), (5, 130), (2, 600)])
@pytest.mark.parametrize("k", [1, 12, 128])
def test_zero_threshold_rows_shorter_and_longer_than_the_list(ctx, oracle, shape, k):
    """z_min = 0: every bin is a candidate; such rows always take the exact dense path, whatever their length."""
    rng = np.random.default_rng(shape[1] + k)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    img[:, ::3] = 0                      # many ties at the threshold itself
    _check_scan(ctx, oracle, img, z_min=0.0, k_strongest=k, min_distance=0.0)
    _check_scan(ctx, oracle, np.zeros(shape, np.uint8), z_min=0.0, k_strongest=k, min_distance=0.0)


@pytest.mark.parametrize("n_real", [0, 5, 12, 13, 300])
def test_high_threshold_rows_flagged_everywhere_but_holding_few_candidates(ctx, oracle, n_real):
    """z_min > 128: the conservative vector test flags every byte >= 128, so rows full of values in [128, z_min) overflow the queue of
    flagged vectors and take the dense path although they hold fewer than k real candidates (count(>= z_min) < k at the start of the
    threshold search) — or exactly k, or more."""
    rng = np.random.default_rng(100 + n_real)
    img = rng.integers(130, 200, size=(6, 3768), dtype=np.uint8)
    for b in range(6):
        at = rng.choice(3768, n_real, replace=False)
        img[b, at] = rng.integers(200, 256, n_real)
    img[5, :] = 199                                          # all-equal row just below the threshold
    _check_scan(ctx, oracle, img, z_min=200.0, k_strongest=12, min_distance=0.0)
    _check_scan(ctx, oracle, img, z_min=200.0, k_strongest=40, min_distance=0.0)


def test_edge_bins_and_row_crossing(ctx, oracle):
    """Kept bins within 6 of either row end: NMS reads across the row edge (flat cv::Mat indexing)."""
    img = np.full((6, 128), 20, np.uint8)
    for b in range(6):
        for r in (0, 1, 2, 3, 5, 60, 61, 122, 124, 125, 126, 127):
            img[b, r] = 100 + 7 * b + (r % 5)
    _check_scan(ctx, oracle, img, z_min=60.0, k_strongest=12, min_distance=0.0)
    _check_scan(ctx, oracle, img, z_min=60.0, k_strongest=5, min_distance=0.0)


def test_batch_matches_single(ctx, oracle, stream8):
    f, p = ctx.StructuredKStrongest(stream8.scans[:4], z_min=60.0, k_strongest=40)
    for b in range(4):
        ref = oracle.kstrongest(stream8.scans[b], z_min=60.0, k=40)
        _assert_same(ref["filtered"], f.scan(b), f"filtered[{b}]")
        _assert_same(ref["peaks"], p.scan(b), f"peaks[{b}]")


def test_mulran_rotation(ctx, oracle):
    rng = np.random.default_rng(9)
    src = rng.integers(0, 256, size=(3360, 400), dtype=np.uint8)  # range-major MONO8 as delivered (radar_driver.cpp:74-90)
    assert np.array_equal(ctx.rotate90ccw(src), oracle.rotate90ccw(src))
    assert np.array_equal(ctx.rotate90ccw(src), np.rot90(src, 1))


def test_full_size_properties(ctx):
    """BASELINE-size batch (64 Oxford scans): sortedness, thresholds, top-k dominance via numpy (no oracle loop)."""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 120, size=(64, 400, 3768), dtype=np.uint8)
    k, z = 40, 60
    f, _ = ctx.StructuredKStrongest(img, z_min=float(z), k_strongest=k, min_distance=0.0)
    for b in (0, 17, 63):
        az, rg, I, x, y = f.scan(b)
        assert np.all(I >= z)
        assert np.array_equal(I, img[b][az, rg])
        key = I.astype(np.int64) * 65536 + rg
        same = az[1:] == az[:-1]
        assert np.all(az[1:] >= az[:-1]) and np.all(key[1:][same] > key[:-1][same])
        for a in (0, 199, 399):
            row = img[b, a].astype(np.int64)
            full = row * 65536 + np.arange(row.size)
            full = full[(row >= z)]
            want = np.sort(full)[-k:]
            want = want[(want % 65536) > 0]  # min_range_bin = ceil(0/res) = 0 -> bins > 0
            assert np.array_equal(key[az == a], want)


def test_compensate(ctx, oracle, stream8):
    ref = oracle.kstrongest(stream8.scans[0])["filtered"]
    x, y = ref[3], ref[4]
    for mot, ccw in [((2.5, 0.03, 0.02), False), ((-1.0, 0.5, -0.1), True), ((0.0, 0.0, 0.0), False)]:
        ox, oy = oracle.compensate(x, y, mot, ccw)
        gx, gy = ctx.Compensate(x, y, mot, ccw)
        dx = np.abs(ox.view(np.int32).astype(np.int64) - gx.view(np.int32).astype(np.int64))
        dy = np.abs(oy.view(np.int32).astype(np.int64) - gy.view(np.int32).astype(np.int64))
        assert dx.max() <= 1 and dy.max() <= 1, "compensation differs by more than 1 float ulp"
        assert (dx > 0).mean() < 1e-3 and (dy > 0).mean() < 1e-3


@pytest.mark.parametrize("w,g,pfa,zmin", [(40, 10, 0.01, 20.0), (100, 10, 1e-3, 20.0), (20, 0, 0.1, 0.0), (500, 2, 1e-4, 60.0)])
def test_cacfar(ctx, oracle, stream8, w, g, pfa, zmin):
    """K1b: CA-CFAR detections (azimuth, range, intensity) and x,y bit-exact vs the oracle, in the reference's order."""
    for img in (stream8.scans[1], synth.stress_image("uniform", n_az=16, n_range=1000, seed=8)):
        ref = oracle.cacfar(img, window_size=w, false_alarm_rate=pfa, nb_guard_cells=g, static_threshold=zmin)
        out = ctx.AzimuthCACFAR(img, window_size=w, false_alarm_rate=pfa, nb_guard_cells=g, static_threshold=zmin)
        _assert_same(ref, out.scan(0), "cfar")


def test_cacfar_batch_and_row_ends(ctx, oracle):
    rng = np.random.default_rng(12)
    img = rng.integers(0, 60, size=(3, 8, 300), dtype=np.uint8)
    img[:, :, :4] = 200     # bins whose trailing window is empty (0/0 -> rejected)
    img[:, :, -5:] = 220    # bins whose leading window is clipped / empty
    out = ctx.AzimuthCACFAR(img, window_size=10, nb_guard_cells=3, min_distance=0.0, static_threshold=20.0)
    for b in range(3):
        ref = oracle.cacfar(img[b], window_size=10, nb_guard_cells=3, min_distance=0.0, static_threshold=20.0)
        _assert_same(ref, out.scan(b), f"cfar[{b}]")
