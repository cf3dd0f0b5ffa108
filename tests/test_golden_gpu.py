"""CUDA filter stage vs the reference's own output (tests/golden/filters_ref.npz, produced by the reference's unmodified
radar_filters.cpp / cfar.cpp — tests/golden/make_golden.py), through the C-ABI.  Bar: bit-exact clouds in the reference's order."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


def _cases(kind):
    _, manifest = G.load()
    return [m["name"] for m in manifest if m["kind"] == kind]


@pytest.fixture(scope="module")
def golden():
    z, manifest = G.load()
    return z, {m["name"]: m for m in manifest}


@pytest.mark.parametrize("name", _cases("ks"))
def test_cuda_kstrongest_equals_reference_golden(ctx, golden, name):
    z, man = golden
    m = man[name]
    p = m["params"]
    f, pk = ctx.StructuredKStrongest(z[m["image"]], z_min=p["z_min"], k_strongest=p["k"], min_distance=p["min_distance"], range_res=p["range_res"])
    for which, cloud in (("filtered", f), ("peaks", pk)):
        az, rg, I, x, y = cloud.scan(0)
        G.assert_cloud([z[f"{name}.{which}.{c}"] for c in "xyi"], x, y, I, f"{name}.{which}")


@pytest.mark.parametrize("name", _cases("cfar"))
def test_cuda_cacfar_equals_reference_golden(ctx, golden, name):
    z, man = golden
    m = man[name]
    p = m["params"]
    out = ctx.AzimuthCACFAR(z[m["image"]], window_size=p["window_size"], false_alarm_rate=G.f32(p["false_alarm_rate"]),
                            nb_guard_cells=p["nb_guard_cells"], range_res=G.f32(p["range_res"]), static_threshold=G.f32(p["static_threshold"]),
                            min_distance=G.f32(p["min_distance"]), max_distance=400.0)
    az, rg, I, x, y = out.scan(0)
    G.assert_cloud([z[f"{name}.{c}"] for c in "xyi"], x, y, I, name)
