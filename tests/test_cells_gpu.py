"""K3 parity: CUDA oriented surface points ("cells", MapPointNormal) vs the oracle, through the C-ABI.

Bar: identical cell count and order, identical neighbour counts and weight sums (the voxel-grid samples and the float radius
search are integer / index work: bit-exact).  Means / covariances are sums of ~100 fp64 terms: the kernel adds them in a
fixed lane tree, the reference in neighbour order -> tolerance 1e-12 relative (SURVEY §7 "hard parts": cells are
tolerance-parity given identical neighbour sets); eigenvalues / normals / planarity inherit that through a 2x2 eigen-solve.
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
    for f in (NS, SUMI, AVGI):
        assert np.array_equal(ref[:, f], got[:, f]), f"field {f} (integer-valued) differs"
    if len(ref) == 0:
        return
    scale = np.abs(ref[:, [U0, U1]]).max()
    assert np.abs(ref[:, [U0, U1]] - got[:, [U0, U1]]).max() <= 1e-13 * max(scale, 1.0), "means differ"
    cov_scale = np.abs(ref[:, C00:C11 + 1]).max(axis=1, keepdims=True)
    assert (np.abs(ref[:, C00:C11 + 1] - got[:, C00:C11 + 1]) <= 1e-11 * cov_scale).all(), "covariances differ"
    for f in (LMIN, LMAX):
        assert np.allclose(ref[:, f], got[:, f], rtol=1e-9, atol=0), f"eigenvalue {f} differs"
    # eigenvectors: conditioning ~ 1 / (relative eigenvalue gap)
    gap = (ref[:, LMAX] - ref[:, LMIN]) / ref[:, LMAX]
    tol = 1e-11 / np.maximum(gap, 1e-6)
    assert (np.abs(ref[:, N0] - got[:, N0]) <= tol).all() and (np.abs(ref[:, N1] - got[:, N1]) <= tol).all(), "normals differ"
    assert (np.abs(np.abs(ref[:, O0] * got[:, O0] + ref[:, O1] * got[:, O1]) - 1) <= tol).all(), "orth normals differ"
    assert np.allclose(ref[:, SCALE], got[:, SCALE], rtol=1e-9, atol=1e-12), "planarity differs"


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
