"""The timing table under the reference's key names: format pinned to the reference's own statistics.cpp (through oracle/_ref when it
is present), stage mapping covers every kernel the library launches."""
import os
import re

import pytest

from tbv_slam_public_b200 import statistics as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _triplets(text):
    lines = [l for l in text.split("\n") if l]
    return sorted(tuple(lines[i:i + 3]) for i in range(0, len(lines), 3))


def test_format_matches_the_reference_statistics_class():
    ref_py = pytest.importorskip("oracle.ref_py")
    if not ref_py.available():
        pytest.skip("oracle/_ref not built here")
    pairs = [("Filtering", 1.5), ("register", 3.25), ("Filtering", 2.5), ("build_normals", 0.125), ("Filtering", 4.0), ("Pose grapgh optimization", 1230.0)]
    st = S.statistics()
    for n, v in pairs:
        st.Document(n, v)
    assert _triplets(st.GetStatistics()) == _triplets(ref_py.statistics(pairs))     # the reference iterates an unordered_map: order is free


def test_every_launched_kernel_has_a_stage():
    names = set()
    for f in os.listdir(os.path.join(ROOT, "tbv_slam_public_b200", "csrc")):
        if f.endswith(".cu"):
            names |= set(re.findall(r'launched\(ctx, "([a-z0-9_]+)"\)', open(os.path.join(ROOT, "tbv_slam_public_b200", "csrc", f)).read()))
    helpers = {"cells_aos_to_soa", "cells_soa_to_aos", "k_odom_reset"}                 # data movement / reset: no reference stage
    missing = sorted(n for n in names - helpers if n not in S.STAGE_OF_KERNEL)
    assert not missing, missing


def test_document_profile_folds_kernels_into_stages():
    st = S.statistics()
    for step in range(2):
        S.document_profile(st, [("k1_filter_fused", 0.34), ("cells_fused", 0.2), ("k_register", 0.44), ("k_odom_update", 0.02)])
    S.document_profile(st, [("k_register", 0.4), ("k_pack_constraints", 0.01)], loop_registration=True)
    assert all(abs(v - 0.34) < 1e-12 for v in st.t["Filtering"]) and len(st.t["Filtering"]) == 2
    assert st.t["register"] == [0.44, 0.44] and abs(st.t["Register"][0] - 0.41) < 1e-12
    assert "Filtering avg, 0.340000\nFiltering dev [σ], 0.000000\nFiltering count, 2\n" in st.GetStatistics()
