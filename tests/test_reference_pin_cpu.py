"""Pins the oracle's filter stage to the REFERENCE'S OWN CODE.

(1) tests/golden/filters_ref.npz holds clouds produced by the reference's unmodified radar_filters.cpp / cfar.cpp
    (tests/golden/make_golden.py; compiled where they lie against the container stand-ins of oracle/ref_shim/).  The oracle
    must reproduce every case bit for bit — this runs anywhere.
(2) Where oracle/_ref/libtbv_ref_filters.so exists (the build container; it also travels to the GPU box), the oracle is
    compared live against the reference on full-size Oxford / MulRan scans and random stress rows, and the committed fixture
    is checked to be what the reference produces today.
"""
import numpy as np
import pytest

from tbv_slam_public_b200 import synth

import golden_util as G


def _cases(kind):
    _, manifest = G.load()
    return [m["name"] for m in manifest if m["kind"] == kind]


@pytest.fixture(scope="module")
def golden():
    z, manifest = G.load()
    return z, {m["name"]: m for m in manifest}


@pytest.mark.parametrize("name", _cases("ks"))
def test_oracle_kstrongest_equals_reference_golden(oracle, golden, name):
    z, man = golden
    m = man[name]
    p = m["params"]
    r = oracle.kstrongest(z[m["image"]], z_min=p["z_min"], k=p["k"], min_distance=p["min_distance"], range_res=p["range_res"])
    for which in ("filtered", "peaks"):
        az, rg, I, x, y = r[which]
        G.assert_cloud([z[f"{name}.{which}.{c}"] for c in "xyi"], x, y, I, f"{name}.{which}")


@pytest.mark.parametrize("name", _cases("cfar"))
def test_oracle_cacfar_equals_reference_golden(oracle, golden, name):
    z, man = golden
    m = man[name]
    p = m["params"]
    az, rg, I, x, y = oracle.cacfar(z[m["image"]], p["window_size"], G.f32(p["false_alarm_rate"]), p["nb_guard_cells"], G.f32(p["range_res"]),
                                    G.f32(p["static_threshold"]), G.f32(p["min_distance"]), 400.0)
    G.assert_cloud([z[f"{name}.{c}"] for c in "xyi"], x, y, I, name)


ref_py = pytest.importorskip("oracle.ref_py")
needs_ref = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built here (needs /root/reference; `make -C oracle ref`)")


@needs_ref
def test_golden_fixture_is_what_the_reference_produces(golden):
    z, man = golden
    for name, m in man.items():
        p = m["params"]
        if m["kind"] == "ks":
            r = ref_py.kstrongest(z[m["image"]], **p)
            for which in ("filtered", "peaks"):
                for c, a in zip("xyi", r[which]):
                    assert G.same_bits(z[f"{name}.{which}.{c}"], a), (name, which, c)
        else:
            r = ref_py.cacfar(z[m["image"]], max_distance=400.0, **p)
            for c, a in zip("xyi", r):
                assert G.same_bits(z[f"{name}.{c}"], a), (name, c)


@needs_ref
@pytest.mark.parametrize("dataset,k,zmin", [("oxford", 40, 60.0), ("oxford", 12, 70.0), ("mulran", 40, 60.0), ("mulran", 12, 70.0)])
def test_oracle_equals_live_reference_full_scan(oracle, dataset, k, zmin):
    st = synth.make_stream(3)
    rr = 0.0438 if dataset == "oxford" else 0.0595238
    for f in range(3):
        img = st.scans[f] if dataset == "oxford" else np.ascontiguousarray(st.scans[f][:, :3360])
        a = oracle.kstrongest(img, z_min=zmin, k=k, min_distance=2.5, range_res=rr)
        b = ref_py.kstrongest(img, z_min=zmin, k=k, min_distance=2.5, range_res=rr)
        for which in ("filtered", "peaks"):
            az, rg, I, x, y = a[which]
            G.assert_cloud(b[which], x, y, I, f"{dataset} frame {f} {which}")
            assert len(x) > 100


@needs_ref
@pytest.mark.parametrize("kind", ["uniform", "equal", "ramp", "sparse", "zeros"])
def test_oracle_equals_live_reference_stress(oracle, kind):
    img = synth.stress_image(kind, n_az=40, n_range=3768, seed=9)
    for k, zmin in [(40, 60.0), (12, 0.0), (40, 255.0), (128, 1.0)]:
        a = oracle.kstrongest(img, z_min=zmin, k=k, min_distance=2.5, range_res=0.0438)
        b = ref_py.kstrongest(img, z_min=zmin, k=k, min_distance=2.5, range_res=0.0438)
        for which in ("filtered", "peaks"):
            az, rg, I, x, y = a[which]
            G.assert_cloud(b[which], x, y, I, f"{kind} k={k} z={zmin} {which}")


@needs_ref
def test_oracle_cacfar_equals_live_reference_full_scan(oracle):
    img = synth.make_stream(1).scans[0]
    for w, g, pfa, zt in [(40, 10, 0.01, 20.0), (10, 20, 0.01, 60.0), (200, 10, 0.001, 20.0)]:
        az, rg, I, x, y = oracle.cacfar(img, w, G.f32(pfa), g, G.f32(0.0438), zt, 2.5, 400.0)
        G.assert_cloud(ref_py.cacfar(img, w, pfa, g, 0.0438, zt, 2.5, 400.0), x, y, I, f"cfar w={w}")
        assert len(x) > 100
