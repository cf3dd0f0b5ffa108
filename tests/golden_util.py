"""Loader for tests/golden/filters_ref.npz — clouds produced by the reference's own radar_filters.cpp / cfar.cpp
(tests/golden/make_golden.py).  Shared by the CPU pin of the oracle and the GPU parity test."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filters_ref.npz")


def load():
    z = np.load(PATH)
    manifest = json.loads(bytes(z["manifest"]).decode())
    return z, manifest


def f32(v) -> float:
    """A float parameter of radarDriver::Parameters widened to double at the call (radar_driver.h:40-45)."""
    return float(np.float32(v))


def same_bits(a, b) -> bool:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def assert_cloud(golden_xyz, x, y, intensity, what):
    gx, gy, gi = golden_xyz
    assert len(gx) == len(x), f"{what}: reference has {len(gx)} points, got {len(x)}"
    assert same_bits(gx, x), f"{what}: x differs (bitwise)"
    assert same_bits(gy, y), f"{what}: y differs (bitwise)"
    assert np.array_equal(gi, np.asarray(intensity).astype(np.float32)), f"{what}: intensity differs"
