"""The CUDA path against the REFERENCE's own MapPointNormal / n_scan_normal_reg / RSCManager records (tests/golden/ref_downstream.npz, produced by
tools/make_ref_fixtures.sh inside the reference's docker image) — and, always, against the oracle's records through the same harness."""
import os

import numpy as np
import pytest

import ref_downstream_util as U
from tbv_slam_public_b200 import api

pytestmark = pytest.mark.gpu


class GpuBackend(U.OracleBackend):
    """Every stage through the C-ABI (api.Context); only cells_as_reference_layout is inherited (a column shuffle)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def kstrongest(self, scan):
        f, p = self.ctx.StructuredKStrongest(scan, peaks=True)
        out = {}
        for key, buf in (("filtered", f), ("peaks", p)):
            az, rg, I, x, y = buf.scan(0)
            out[key] = (x.copy(), y.copy(), I.astype(np.float32))
        return out

    def compensate(self, x, y, mot):
        return self.ctx.Compensate(x, y, mot, False) if len(x) else (x, y)

    def build_cells(self, x, y, I):
        return self.ctx.MapPointNormal(x, y, I, radius=3.0, weight_intensity=True, capacity=2048)[0]

    def register(self, scans, T, loop):
        P = api.loop_reg_params() if loop else api.default_reg_params(weight_opt=api.W_COMBINED)
        To, s = self.ctx.Register(scans, T, P)
        return bool(s.success), To[-1], s.score, s.itrs

    def get_cost(self, scans, T):
        n, score, cost, res = self.ctx.GetCost(scans, T, api.default_reg_params(cost=api.P2L, loss=api.HUBER, loss_limit=0.3, weight_opt=api.W_UNIFORM), itr=0)
        return n > 1, cost, res

    def scan_context(self, peaks, T):
        rsc, out = api.RSCManager(self.ctx), []
        for (x, y, I), t in zip(peaks, T):
            rsc.makeAndSaveScancontextAndKeysRadarCloud(x, y, I, t)
            out.append((rsc.polarcontexts[-1], [(c["nn_idx"], c["argmin_shift"], c["min_dist"], c["min_dist_sc"], c["min_dist_odom"], c["yaw_diff_rad"])
                                                for c in rsc.detectLoopClosureID()]))
        return out


@pytest.fixture(scope="module")
def gpu_records(ctx):
    return U.compute_records(GpuBackend(ctx))


def test_gpu_equals_the_oracle_through_the_reference_fixture_harness(gpu_records, oracle):
    U.compare(gpu_records, U.compute_records(U.OracleBackend(oracle)))


@pytest.mark.skipif(not os.path.exists(U.FIXTURE), reason="tests/golden/ref_downstream.npz absent: run tools/make_ref_fixtures.sh inside the reference's docker image")
def test_gpu_equals_the_reference_downstream_of_the_filter(gpu_records):
    U.compare(gpu_records, dict(np.load(U.FIXTURE)))
