"""K3 parity: CUDA oriented surface points ("cells", MapPointNormal) vs the oracle, through the C-ABI.

Bar: identical cell count and order, identical neighbour counts (the voxel-grid samples and the float radius search
are integer / index work: bit-exact), means / covariances / eigen-decomposition bit-exact (same operation order, no FMA),
planarity (a log) within 4 ulp (device log is not glibc's).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

U0, U1, C00, C01, C10, C11, SCALE, N0, N1, O0, O1, LMIN, LMAX, SUMI, AVGI, NS = range(16)


def _cloud(oracle, img, **kw):
    az, rg, I, x, y = oracle.kstrongest(img, **kw)["filtered"]
    return x, y, I.astype(np.float32)


def _compare(ref, got):
    assert got.shape == ref.shape, f"{ref.shape[0]} oracle cells vs {got.shape[0]} gpu"
    assert np.array_equal(ref[:, NS], got[:, NS]), "neighbour counts differ"
    exact = [U0, U1, C00, C01, C10, C11, N0, N1, O0, O1, LMIN, LMAX, SUMI, AVGI]
    for f in exact:
        assert np.array_equal(ref[:, f], got[:, f]), f"field {f} differs: max abs {np.abs(ref[:, f] - got[:, f]).max()}"
    assert np.allclose(ref[:, SCALE], got[:, SCALE], rtol=1e-15 * 4, atol=0), "planarity differs by more than 4 ulp"


@pytest.mark.parametrize("k,z,radius,wint", [(40, 60.0, 3.0, True), (12, 70.0, 3.5, False), (12, 60.0, 3.0, True)])
def test_cells_radar_like(ctx, oracle, stream8, k, z, radius, wint):
    for i in (0, 5):
        x, y, I = _cloud(oracle, stream8.scans[i], k=k, z_min=z)
        ref, ns_ref = oracle.build_cells(x, y, I, radius=radius, weight_intensity=wint)
        got, ns = ctx.MapPointNormal(x, y, I, radius=radius, weight_intensity=wint)
        assert ns == ns_ref
        assert len(ref) > 50
        _compare(ref, got)


def test_cells_downsample_factor_and_origin(ctx, oracle, stream8):
    x, y, I = _cloud(oracle, stream8.scans[2])
    for f, origin in [(2.0, (0.0, 0.0)), (1.0, (10.0, -5.0)), (0.5, (0.0, 0.0))]:
        ref, ns_ref = oracle.build_cells(x, y, I, radius=3.0, downsample_factor=f, origin=origin)
        got, ns = ctx.MapPointNormal(x, y, I, radius=3.0, downsample_factor=f, origin=origin)
        assert ns == ns_ref
        _compare(ref, got)


def test_cells_random_clouds(ctx, oracle):
    rng = np.random.default_rng(5)
    for n in (1, 5, 6, 7, 50, 3000):
        x = rng.uniform(-40, 40, n).astype(np.float32)
        y = rng.uniform(-40, 40, n).astype(np.float32)
        I = rng.integers(55, 200, n).astype(np.float32)
        ref, ns_ref = oracle.build_cells(x, y, I, radius=3.0)
        got, ns = ctx.MapPointNormal(x, y, I, radius=3.0)
        assert ns == ns_ref
        _compare(ref, got)


def test_cells_degenerate(ctx, oracle):
    # collinear points (zero smallest eigenvalue -> invalid), duplicated points, all-weights-zero neighbourhoods
    t = np.linspace(0, 30, 400).astype(np.float32)
    x, y = t, (0.5 * t).astype(np.float32)
    I = np.full(400, 100, np.float32)
    ref, _ = oracle.build_cells(x, y, I, radius=3.0)
    got, _ = ctx.MapPointNormal(x, y, I, radius=3.0)
    _compare(ref, got)
    x2 = np.repeat(np.float32([1.0, 2.0, 3.0]), 10); y2 = np.repeat(np.float32([1.0, 1.5, 0.2]), 10)
    I2 = np.full(30, 90, np.float32)
    ref, _ = oracle.build_cells(x2, y2, I2, radius=3.0)
    got, _ = ctx.MapPointNormal(x2, y2, I2, radius=3.0)
    _compare(ref, got)
    got, ns = ctx.MapPointNormal(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32))
    assert len(got) == 0 and ns == 0
