"""The oracle against the REFERENCE's own MapPointNormal / n_scan_normal_reg / RSCManager on the committed synthetic scans
(tests/golden/ref_downstream.npz, produced inside the reference's docker image by tools/make_ref_fixtures.sh — VERDICT r1 item 6a).

The fixture cannot be produced in this repository's build container (no Eigen / PCL / FLANN / Ceres / ROS), so until a maintainer runs the
recipe the comparison SKIPS, and DESIGN.md §2 keeps saying "parity unpinned" for rows a5-a20.  What runs here regardless: the comparison
harness on records recomputed by the oracle itself (it must accept them and must reject a perturbed copy), so that the day the fixture
arrives the test is known to work."""
import os

import numpy as np
import pytest

import ref_downstream_util as U


@pytest.fixture(scope="module")
def oracle_records(oracle):
    return U.compute_records(U.OracleBackend(oracle))


def test_harness_accepts_identical_records_and_rejects_perturbed_ones(oracle_records):
    rec = oracle_records
    assert len(rec) == 6 * 3 + 5 * 3 + 6 * 2
    assert rec["cells_0"].size % 14 == 0 and rec["cells_0"].size // 14 > 100 and rec["register_3"][0] == 1.0 and rec["sc_desc_0"].size == 4800
    U.compare(rec, rec)
    for key, delta in (("register_3", [0, 2e-5, 0, 0, 0, 0]), ("cells_2", None), ("get_cost_2", None)):
        bad = dict(rec)
        v = rec[key].copy()
        if delta is None:
            v[1] += 1e-6
        else:
            v += np.array(delta)
        bad[key] = v
        with pytest.raises(AssertionError):
            U.compare(bad, rec)


@pytest.mark.skipif(not os.path.exists(U.FIXTURE), reason="tests/golden/ref_downstream.npz absent: run tools/make_ref_fixtures.sh inside the reference's docker image")
def test_oracle_equals_the_reference_downstream_of_the_filter(oracle_records):
    ref = dict(np.load(U.FIXTURE))
    U.compare(oracle_records, ref)
