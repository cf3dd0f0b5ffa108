"""Dress rehearsal of the GPU tests that were written after the round's GPU budget was spent (tests/test_offline_odometry_gpu.py, the batched /
sharded part of tests/test_tbv_slam_gpu.py): their bodies are run HERE against an oracle-backed imitation of the ctypes API objects they use
(api.Context methods, api.OdometryKeyframeFuser, api.RSCManager, api.LoopDB, api.CorAlRadarQuality, parallel.ShardedLoopClosure), so that a
wrong attribute name, array shape, index or tolerance in the test code shows up on CPU instead of on the GPU box.  Test infrastructure only:
nothing in the product imports this; on the GPU the same test bodies run against the real library."""
import types

import numpy as np
import pytest

from tbv_slam_public_b200 import api, parallel
import test_offline_odometry_gpu as OG
import test_tbv_slam_gpu as SG
from test_tbv_slam_cpu import OracleLoopDevice, drive  # noqa: F401  (fixture)


class _Buf:
    def __init__(self, tup):
        self.t = tup

    def scan(self, b):
        assert b == 0
        return self.t


class FakeCtx:
    """The subset of api.Context the rehearsed tests touch, computed by the oracle."""

    def __init__(self):
        from oracle import oracle_py as O
        O.lib()
        self.O, self.n = O, 0
        self.stream = 0

    def launch_count(self):
        return self.n

    def StructuredKStrongest(self, polar, z_min=60.0, k_strongest=40, min_distance=2.5, range_res=0.0438, peaks=True, n_range=None):
        self.n += 2
        r = self.O.kstrongest(np.asarray(polar), z_min, k_strongest, min_distance, range_res, peaks=peaks)
        return _Buf(r["filtered"]), (_Buf(r["peaks"]) if peaks else None)

    def AzimuthCACFAR(self, polar, window_size=40, false_alarm_rate=0.01, nb_guard_cells=10, range_res=0.0438, static_threshold=20.0,
                      min_distance=2.5, max_distance=400.0, capacity=None):
        self.n += 2
        return _Buf(self.O.cacfar(np.asarray(polar), window_size, false_alarm_rate, nb_guard_cells))

    def Compensate(self, x, y, mot, ccw=False):
        self.n += 1
        return self.O.compensate(x, y, mot, ccw)

    def MapPointNormal(self, x, y, intensity, radius=3.0, downsample_factor=1.0, weight_intensity=True, origin=(0.0, 0.0), capacity=None):
        self.n += 1
        return self.O.build_cells(x, y, intensity, radius=radius, downsample_factor=downsample_factor, weight_intensity=weight_intensity)

    def Register(self, scans, T, params=None):
        self.n += 1
        rp = params or api.default_reg_params()
        P = self.O.default_reg_params(cost=rp.cost, loss=rp.loss, weight_opt=rp.weight_opt, loss_limit=rp.loss_limit, cov_scale=rp.cov_scale,
                                      regularization=rp.regularization, max_itr_association=rp.max_itr_association, max_itr_solver=rp.max_itr_solver)
        return self.O.register(scans, T, P)

    def CFEARQualityBatch(self, sets, src_set, ref_set, T_src, T_ref, T_offset=None, params=None):
        self.n += 1
        return OracleLoopDevice().cfear(sets, src_set, ref_set, T_src, T_ref, T_offset)


class FakeFuser:
    """api.OdometryKeyframeFuser for n_seq = 1."""

    def __init__(self, ctx, n_seq, n_az, n_range, params=None):
        assert n_seq == 1
        self.ctx, self.od = ctx, ctx.O.Odometry()

    def pointcloudCallback(self, polar):
        assert polar.ndim == 3 and polar.shape[0] == 1
        self.ctx.n += 6
        o = self.od.step(polar[0])
        return [types.SimpleNamespace(pose=[o.pose[0], o.pose[1], o.pose[2]], n_points=o.n_points, n_cells=o.n_cells, itrs=o.itrs, reg_ok=o.reg_ok,
                                      is_keyframe=o.is_keyframe, n_keyframes=o.n_keyframes, lm_iterations=0, num_residuals=0, status=0, n_samples=0,
                                      score=0.0)]

    def cells(self, seq, keyframe=-1, capacity=8192):
        poses, _ = self.od.keyframes()
        return self.od.keyframe_cells(keyframe), poses[keyframe]

    def close(self):
        pass


class FakeRSC:
    def __init__(self, ctx, params=None):
        self.ctx, self.dev = ctx, OracleLoopDevice()

    def makeAndSaveScancontextAndKeysRadarCloud(self, x, y, intensity, Todom):
        self.ctx.n += 1
        self.dev.rsc.add(x, y, intensity, Todom)

    def detectLoopClosureID(self):
        self.ctx.n += 2
        return self.dev.detect()


class FakeLoopDB:
    def __init__(self, ctx, max_keyframes, cell_capacity=1024):
        self.ctx, self.dev = ctx, OracleLoopDevice()

    def add(self, sets):
        self.ctx.n += 1
        first = len(self.dev.cells)
        for s in sets:
            self.dev.add_keyframe(s)
        return first

    def register_candidates(self, id_from, id_to, T_from, T_to, candidate_index=None, quality=None, params=None, max_score=0.0, want_summaries=False):
        self.ctx.n += 2
        res = self.dev.register(list(id_from), list(id_to), np.asarray(T_from).reshape(-1, 3), np.asarray(T_to).reshape(-1, 3))
        out = np.zeros(len(res), api.CONSTRAINT_DTYPE)
        summ, n = [], 0
        for p, (ok, t, cov, score) in enumerate(res):
            summ.append(types.SimpleNamespace(success=int(ok), score=score, itrs=0))
            if ok:
                out[n]["candidate"], out[n]["id_begin"], out[n]["id_end"], out[n]["type"] = p, id_from[p], id_to[p], 1
                out[n]["t_be"], out[n]["cov"], out[n]["score"] = t, cov, score
                n += 1
        return (out[:n], summ) if want_summaries else out[:n]

    def close(self):
        pass


class FakeSharded:
    def __init__(self, db, group=None):
        self.db = db

    def register_candidates(self, id_from, id_to, T_from, T_to, quality=None, params=None, max_score=0.0):
        return self.db.register_candidates(id_from, id_to, T_from, T_to)


@pytest.fixture()
def fake_api(monkeypatch):
    dev = OracleLoopDevice()
    monkeypatch.setattr(api, "OdometryKeyframeFuser", FakeFuser)
    monkeypatch.setattr(api, "RSCManager", FakeRSC)
    monkeypatch.setattr(api, "LoopDB", FakeLoopDB)
    monkeypatch.setattr(parallel, "ShardedLoopClosure", FakeSharded)

    def coral(ctx, clouds, src_cloud, ref_cloud, T_src, T_ref, T_offset=None, radius=1.0, weight_res_intensity=False, per_point=False):
        ctx.n += 1
        q = dev.coral(clouds, src_cloud, ref_cloud, T_src, T_ref, T_offset)
        return [types.SimpleNamespace(joint=a, sep=b, overlap=c) for a, b, c in q]

    monkeypatch.setattr(api, "CorAlRadarQuality", coral)

    def pgo_assemble(ctx, nodes, ids, meas, params=None, info=None, fixed_node=0):
        ctx.n += 3
        O = dev.O
        P = O.default_pgo_params() if params is None else O.default_pgo_params(
            odom_vxx=params.odom_vxx, odom_vyy=params.odom_vyy, odom_vtt=params.odom_vtt, loop_scaling=params.loop_scaling,
            replace_cov_by_identity=params.replace_cov_by_identity, loop_cauchy=params.loop_cauchy)
        return O.pgo_assemble(nodes, ids, meas, P, info=info, fixed_node=fixed_node)

    def pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node=0, radius=1e4, max_iters=20000, rel_tol=1e-12):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl
        ctx.n += 1
        n, r6 = len(Hd), np.arange(6)
        bi = lambda i: (6 * i[:, None, None] + r6[None, :, None]) + 0 * r6[None, None, :]
        bj = lambda j: (6 * j[:, None, None] + r6[None, None, :]) + 0 * r6[None, :, None]
        nn, a, b = np.arange(n), ids[:, 0].astype(np.int64), ids[:, 1].astype(np.int64)
        A = sp.coo_matrix((np.r_[Hd.ravel(), Ho.ravel(), Ho.ravel()],
                           (np.r_[bi(nn).ravel(), bi(a).ravel(), bj(b).ravel()], np.r_[bj(nn).ravel(), bj(b).ravel(), bi(a).ravel()])),
                          shape=(6 * n, 6 * n)).tocsr()
        A = A + sp.diags(np.clip(A.diagonal(), 1e-6, 1e32) / radius)
        keep = np.r_[0:6 * fixed_node, 6 * fixed_node + 6:6 * n]
        x = np.zeros(6 * n)
        x[keep] = spl.spsolve(A[keep][:, keep].tocsc(), -g.reshape(-1)[keep])
        return x.reshape(n, 6), 1, 0.0

    def pgo_solve_damped(ctx, ids, Hd, Ho, g, damping, fixed_node=0, max_iters=0, rel_tol=0):
        # (H + diag(damping)) delta = -g through the radius interface: a diagonal chosen so that clamp(e) / 1 adds exactly `damping`
        n, idx = len(Hd), np.arange(6)
        Hd2 = np.array(Hd, np.float64).reshape(-1, 6, 6).copy()
        total = Hd2[:, idx, idx] + np.asarray(damping, np.float64).reshape(-1, 6)
        e = np.where(total / 2.0 >= 1e-6, total / 2.0, total - 1e-6)
        Hd2[:, idx, idx] = np.where(e > 1e32, total - 1e32, e)
        return pgo_solve_step(ctx, ids, Hd2, Ho, g, fixed_node, 1.0)

    def pgo_optimize_device(ctx, *a, **kw):
        # tbv_pgo_optimize runs the loop of pgo_optimize_ceres on the device; the rehearsal drives the same loop from the host over the fakes
        return api.pgo_optimize_ceres(ctx, *a, **kw)

    monkeypatch.setattr(api, "pgo_assemble", pgo_assemble)
    monkeypatch.setattr(api, "pgo_solve_step", pgo_solve_step)
    monkeypatch.setattr(api, "pgo_solve_damped", pgo_solve_damped)
    monkeypatch.setattr(api, "pgo_optimize_device", pgo_optimize_device)
    return FakeCtx()


def test_rehearse_reader_on_the_gpu(fake_api, tmp_path):
    OG.test_reader_on_the_gpu_matches_the_oracle_backed_reader(fake_api, tmp_path)


def test_rehearse_cloud_interface_fuser_on_the_gpu(fake_api):
    OG.test_cloud_interface_fuser_on_the_gpu_equals_the_fused_frame(fake_api)


def test_rehearse_offline_slam_on_the_gpu(fake_api, drive):
    SG.test_offline_slam_closes_the_loop_on_the_gpu(fake_api, drive)


def test_rehearse_batched_and_sharded_search_on_the_gpu(fake_api, drive):
    SG.test_batched_and_sharded_search_equal_the_per_keyframe_search_on_the_gpu(fake_api, drive)
