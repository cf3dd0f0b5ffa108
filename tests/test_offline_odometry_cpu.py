"""offline_odometry.radarReader — the scan-stream -> trajectory + simple graph loop (offline_odometry.cpp:57-141) — without a GPU: the device
argument is an oracle-backed stand-in with the interface of offline_odometry.GpuOdometryDevice (test infrastructure only).  The strongest
check is self-consistency: the cells a keyframe node stores must be exactly the cells one gets by building surface points from the
(compensated) filtered cloud the same node stores — which holds only if the reader pairs every frame with the motion the fuser used."""
import math

import numpy as np
import pytest

from tbv_slam_public_b200 import graph_io as G, offline_odometry as OO, synth, trajectory_io as TIO


class OracleOdometryDevice:
    def __init__(self, **kw):
        from oracle import oracle_py as O
        O.lib()
        self.O, self.params = O, O.default_odom_params(**kw)
        self.od = O.Odometry(self.params)

    def step(self, scan):
        o = self.od.step(scan)
        return dict(pose=np.array([o.pose[0], o.pose[1], o.pose[2]]), is_keyframe=bool(o.is_keyframe), n_keyframes=int(o.n_keyframes),
                    n_cells=int(o.n_cells), score=0.0, itrs=int(o.itrs))          # the oracle's record carries no score

    def newest_keyframe_cells(self, n_keyframes):
        return self.od.keyframe_cells(n_keyframes - 1)

    def clouds(self, scan, motion_xyt):
        f = self.params                                              # the oracle's parameter record is flat
        r = self.O.kstrongest(scan, f.z_min, f.k_strongest, f.min_distance, f.range_res, peaks=True)
        out = []
        for key in ("filtered", "peaks"):
            _, _, I, x, y = r[key]
            if self.params.compensate and len(x):
                x, y = self.O.compensate(x, y, motion_xyt, bool(self.params.radar_ccw))
            out.append(np.c_[x, y, np.zeros(len(x), np.float32), I.astype(np.float32)].astype(np.float32).reshape(-1, 4))
        return out[0], out[1]


@pytest.fixture(scope="module")
def run():
    st = synth.make_stream(26, speed=4.0)                     # 1 m per frame: a keyframe every other frame (threshold 1.5 m)
    dev = OracleOdometryDevice()
    rd = OO.radarReader(dev)
    outs = [rd.process(st.scans[i], stamp_ns=1_000_000_000 + 250_000_000 * i, gt=st.gt[i]) for i in range(len(st.scans))]
    return st, dev, rd, outs


def test_every_frame_is_evaluated_and_keyframes_become_nodes(run):
    st, dev, rd, outs = run
    n_kf = sum(o["is_keyframe"] for o in outs)
    assert len(rd.est) == len(st.scans) == len(rd.stamps) == len(rd.gt)
    assert outs[0]["is_keyframe"] and 8 <= n_kf < len(st.scans) and len(rd.graph) == n_kf
    assert rd.keyframe_rows == [i for i, o in enumerate(outs) if o["is_keyframe"]]
    first, cons0 = rd.graph.graph[0]
    assert cons0 == [] and np.array_equal(first.T, [0, 0, 0, 0, 0, 0, 1]) and np.array_equal(first.motion_, np.eye(4))
    for r in range(1, n_kf):
        scan, cons = rd.graph.graph[r]
        assert scan.idx_ == r and scan.stamp_ == rd.stamps[rd.keyframe_rows[r]] and len(cons) == 1
        c = cons[0]
        assert (c.id_begin, c.id_end, c.type) == (r, r - 1, G.ODOMETRY)
        assert np.allclose(scan.GetPose() @ G.pose3d_to_matrix(c.t_be), rd.graph.graph[r - 1][0].GetPose(), atol=1e-12)
        assert np.allclose(G.pose3d_to_xyt(scan.T), outs[rd.keyframe_rows[r]]["pose"], atol=1e-12)
        d = np.linalg.norm(scan.T[:2] - rd.graph.graph[r - 1][0].T[:2])
        assert d > 1.5 or abs(G.pose3d_to_xyt(c.t_be)[2]) > math.radians(5.0)            # the keyframe rule


def test_stored_cells_are_the_cells_of_the_stored_cloud(run):
    st, dev, rd, outs = run
    for r, (scan, _) in enumerate(rd.graph.graph):
        c = scan.cloud_nopeaks_
        rebuilt, _ = dev.O.build_cells(c[:, 0], c[:, 1], c[:, 3], radius=3.0, weight_intensity=True)
        assert len(rebuilt) == len(scan.cloud_normal_) > 100, r
        assert np.array_equal(rebuilt, scan.cloud_normal_), r                           # same points, same motion -> same bits
        assert 0 < len(scan.cloud_peaks_) <= len(c) and np.array_equal(scan.downsampled_, scan.cloud_normal_[:, :2].astype(np.float32))
    # and they would NOT match with the wrong motion (the check has teeth): re-derive one node's cloud without compensation
    k = rd.keyframe_rows[5]
    raw, _ = OracleOdometryDevice(compensate=0).clouds(st.scans[k], np.zeros(3))
    assert not np.array_equal(raw[:, :2], rd.graph.graph[5][0].cloud_nopeaks_[:, :2])


def test_save_writes_what_the_back_end_and_the_evaluation_read(run, tmp_path):
    st, dev, rd, outs = run
    paths = rd.Save(str(tmp_path), "07")
    assert set(paths) == {"est", "gt", "graph"} and paths["est"].endswith("est/07.txt")
    est, gt = TIO.read_kitti(paths["est"]), TIO.read_kitti(paths["gt"])
    assert len(est) == len(gt) == len(st.scans)
    g = G.load_simple_graph(paths["graph"])
    assert len(g) == len(rd.graph) and all(s.has_Tgt_ for s, _ in g.graph)
    for r, (s, _) in enumerate(g.graph):
        assert np.allclose(G.pose3d_to_xyt(s.Tgt)[:2], st.gt[rd.keyframe_rows[r]][:2], atol=1e-12)
    # odometry quality on the synthetic world: the estimate follows the ground truth (both re-based on their first pose)
    res = TIO.evaluate(gt, est, "6dof", step_size=1)
    assert res["ate"] < 0.25 and res["rpe_trans"] < 0.1
    nodes, ids, meas, info, _ = g.pgo_arrays()
    assert len(ids) == len(g) - 1 and np.all(ids[:, 2] == 0)


def test_run_assigns_sensor_clock_stamps():
    st = synth.make_stream(4)
    rd = OO.radarReader(OracleOdometryDevice()).run(st.scans)
    assert rd.stamps == [250_000_000 * (i + 1) for i in range(4)] and len(rd.graph) == 4 and not rd.gt
    paths_free = rd.graph.graph[3][0]
    assert paths_free.stamp_ == 1_000_000_000 and not paths_free.has_Tgt_


# ---- the fuser with the reference's own (cloud) interface, over primitive device calls ----------------------------------------------------
class OraclePrimitiveDevice:
    def __init__(self):
        from oracle import oracle_py as O
        O.lib()
        self.O = O

    def compensate(self, x, y, mot, ccw):
        return self.O.compensate(x, y, mot, ccw)

    def build_cells(self, x, y, intensity, radius, downsample_factor, weight_intensity):
        return self.O.build_cells(x, y, intensity, radius=radius, downsample_factor=downsample_factor, weight_intensity=weight_intensity)[0]

    def register(self, scans, T, rp):
        P = self.O.default_reg_params(cost=rp.cost, loss=rp.loss, weight_opt=rp.weight_opt, loss_limit=rp.loss_limit, cov_scale=rp.cov_scale,
                                      regularization=rp.regularization, max_itr_association=rp.max_itr_association, max_itr_solver=rp.max_itr_solver)
        Tio, s = self.O.register(scans, T, P)
        return Tio, bool(s.success), int(s.itrs), float(s.score)


def test_cloud_interface_fuser_equals_the_fused_frame():
    """OdometryKeyframeFuser.pointcloudCallback(cloud) = compensate + build_cells + register + the reference's pose / keyframe bookkeeping on the
    host must reproduce the fused frame (oracle.Odometry.step, the checker of tbv_odom_step) when it is fed the k-strongest cloud: poses to
    1e-12 (one atan2 / sincos round trip apart), same keyframe decisions, same association-iteration counts, same window."""
    from tbv_slam_public_b200 import api
    dev = OraclePrimitiveDevice()
    st = synth.make_stream(18, speed=4.0)
    fuser = OO.OdometryKeyframeFuser(dev, api.default_odom_params())
    ref = dev.O.Odometry()
    n_kf = 0
    for i in range(len(st.scans)):
        _, _, I, x, y = dev.O.kstrongest(st.scans[i], 60.0, 40, 2.5, 0.0438, peaks=False)["filtered"]
        pose, (cx, cy), _ = fuser.pointcloudCallback(x, y, I.astype(np.float32))
        o = ref.step(st.scans[i])
        assert np.abs(pose - np.array(o.pose[:])).max() < 1e-12, i
        assert (fuser.updated, fuser.last_itrs, len(fuser.keyframes_), len(fuser.last_cells)) == (bool(o.is_keyframe), o.itrs, o.n_keyframes, o.n_cells), i
        n_kf += fuser.updated
    assert 6 <= n_kf < len(st.scans) and len(fuser.keyframes_) == 4
    kp, _ = ref.keyframes()
    for (k, cells), want in zip(fuser.keyframes_, kp):
        assert np.abs(k.xyt() - want).max() < 1e-12


def test_cloud_interface_fuser_runs_on_ca_cfar_clouds():
    """The reason the cloud interface exists: odometry on CA-CFAR detections (radar_driver.cpp:52-56; the kstrong_vs_cfar presets, SURVEY
    Appendix A) — different points, same fuser.  The estimate must follow the ground truth of the synthetic drive."""
    from tbv_slam_public_b200 import api
    dev = OraclePrimitiveDevice()
    st = synth.make_stream(12)
    par = api.default_odom_params(weight_intensity=0)
    par.reg = api.default_reg_params(cost=api.P2P, weight_opt=api.W_UNIFORM, regularization=1.0)
    fuser = OO.OdometryKeyframeFuser(dev, par)
    est = []
    for i in range(len(st.scans)):
        _, _, I, x, y = dev.O.cacfar(st.scans[i], 40, 0.01, 10)
        assert len(x) > 500
        pose, _, _ = fuser.pointcloudCallback(x, y, I.astype(np.float32))
        est.append(pose)
    est = np.array(est)
    rel_gt = np.array([synth.se2_mul(synth.se2_inv(st.gt[0]), p) for p in st.gt])
    assert np.hypot(*(est[:, :2] - rel_gt[:, :2]).T).max() < 0.5 and np.abs(est[:, 2] - rel_gt[:, 2]).max() < 0.02
    assert np.hypot(*est[-1, :2]) > 20.0                                                   # it did move: 11 steps of 2.5 m


def test_affine_helper_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = OO._Affine.from_xyt(rng.normal(size=3)), OO._Affine.from_xyt(rng.normal(size=3))
        m = lambda t: np.array([[t.r00, t.r01, t.tx], [t.r10, t.r11, t.ty], [0, 0, 1.0]])
        assert np.allclose(m(a @ b), m(a) @ m(b), atol=1e-15) and np.allclose(m(a.inverse()), np.linalg.inv(m(a)), atol=1e-14)
        v = rng.normal(size=3); v[2] = math.remainder(v[2], 2 * math.pi)
        assert np.allclose(OO._Affine.from_xyt(v).xyt(), v, atol=1e-15)


def test_frame_motion_stays_small_when_the_heading_crosses_pi():
    """device poses are atan2-normalised; composing them with raw yaw sums gave a motion yaw near -2 pi for the frame after the
    heading crosses +-pi, and Compensate scales that yaw by the per-point time fraction.  The motion must be the small relative turn."""
    class Stub:
        def __init__(self, poses):
            self.poses, self.k, self.motions = poses, 0, []

        def step(self, scan):
            p = self.poses[self.k]
            self.k += 1
            return dict(pose=np.array(p, float), is_keyframe=True, n_keyframes=self.k, n_cells=0, score=0.0, itrs=0)

        def newest_keyframe_cells(self, n):
            return np.zeros((0, 16))

        def clouds(self, scan, motion_xyt):
            self.motions.append(np.array(motion_xyt, float))
            e = np.zeros((0, 4), np.float32)
            return e, e

    yaws = [3.10, 3.13, -3.13, -3.10, -3.07]                          # +0.03 rad per frame across the branch cut
    dev = Stub([(0.5 * i, 0.0, y) for i, y in enumerate(yaws)])
    rd = OO.radarReader(dev)
    for i in range(len(yaws)):
        rd.process(np.zeros((4, 8), np.uint8), stamp_ns=i + 1)
    for m in dev.motions[2:]:
        assert abs(m[2]) < 0.05, m                                     # was -6.25 for the frame after the crossing
    assert abs(dev.motions[3][2] - (2 * math.pi - 6.26)) < 1e-9
    assert abs(OO._mul((0, 0, 3.0), (0, 0, 1.0))[2] - (4.0 - 2 * math.pi)) < 1e-12 and abs(OO._inv((0, 0, -math.pi))[2] - math.pi) < 1e-12
