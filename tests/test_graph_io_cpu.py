"""graph_io: the simple-graph hand-off — construction rules of OdometryKeyframeFuser::AddToGraph, the .tbvg layout, and
the views that feed tbv_pgo_assemble / tbv_loopdb_add."""
import math
import os
import struct

import numpy as np
import pytest

from tbv_slam_public_b200 import graph_io as G
from tbv_slam_public_b200 import trajectory_io as TIO


def _drive(n, rng, sampled_cov=True):
    g = G.SimpleGraph()
    pose = np.zeros(3)
    for i in range(n):
        if i:
            pose = pose + [1.6 * math.cos(pose[2]), 1.6 * math.sin(pose[2]), 0.04 + 0.01 * rng.normal()]
        a = rng.normal(size=(3, 3))
        cov = 1e-3 * (a @ a.T + 3 * np.eye(3)) if sampled_cov else None
        cells = rng.normal(size=(5 + i % 3, 16))
        g.AddToGraph(pose, cov, stamp_ns=1_547_000_000_000_000_000 + 250_000_000 * i, motion_xyt=(0.4, 0.01, 0.01),
                     cloud_peaks=rng.normal(size=(7, 3)).astype(np.float32), cloud_nopeaks=rng.normal(size=(11, 4)).astype(np.float32),
                     cells=cells, radius=3.0, weight_intensity=True)
    return g


def test_pose_conversions_round_trip_including_half_turns():
    for t in (0.0, 0.3, -2.9, math.pi - 1e-9, -math.pi + 1e-9, 3.1, 2.0):
        pq = G.pose3d_from_xyt((1.0, -2.0, t))
        m = G.pose3d_to_matrix(pq)
        assert np.allclose(m[:2, :2], [[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]], atol=1e-15)
        back = G.pose3d_from_matrix(m)
        assert np.allclose(G.pose3d_to_matrix(back), m, atol=1e-14) and abs(np.linalg.norm(back[3:]) - 1) < 1e-14
        assert np.allclose(G.pose3d_to_xyt(back), (1.0, -2.0, math.atan2(math.sin(t), math.cos(t))), atol=1e-8)
    rng = np.random.default_rng(0)                     # general rotations: all four Shepperd branches
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        m = G.pose3d_to_matrix(np.r_[0, 0, 0, q])
        b = G.pose3d_from_matrix(m)
        assert np.allclose(b[3:], q, atol=1e-12) or np.allclose(b[3:], -q, atol=1e-12)


def test_add_to_graph_follows_the_reference_rules():
    rng = np.random.default_rng(1)
    g = _drive(6, rng)
    assert len(g) == 6 and g.graph[0][1] == []
    for i in range(1, 6):
        scan, cons = g.graph[i]
        assert len(cons) == 1 and scan.idx_ == i
        c = cons[0]
        assert (c.id_begin, c.id_end, c.type) == (i, i - 1, G.ODOMETRY)
        Tfrom, Tto = scan.GetPose(), g.graph[i - 1][0].GetPose()
        assert np.allclose(Tfrom @ G.pose3d_to_matrix(c.t_be), Tto, atol=1e-12)           # t_be = Tfrom^-1 Tto
        assert np.allclose(c.information, c.information.T, atol=1e-6 * np.abs(c.information).max())
        cov = np.linalg.inv(c.information)
        assert np.all(np.isfinite(cov)) and np.all(np.linalg.eigvalsh(0.5 * (cov + cov.T)) > 0)
        assert cov[2, 2] == pytest.approx(1.0) and cov[3, 3] == pytest.approx(1.0)          # identity in the unobserved axes
    # the 6x6 the sampling path builds
    c3 = np.array([[2e-3, 1e-4, 3e-5], [1e-4, 4e-3, -2e-5], [3e-5, -2e-5, 5e-5]])
    c6 = G.cov6_from_xyt(c3)
    assert np.array_equal(c6[:2, :2], c3[:2, :2]) and c6[5, 5] == c3[2, 2] and c6[0, 5] == c3[0, 2] and c6[5, 1] == c3[2, 1]
    assert np.array_equal(c6[2:5, 2:5], np.eye(3)) and c6[0, 2] == 0


def test_information_is_the_rotated_covariance_inverse():
    g = G.SimpleGraph()
    g.AddToGraph((0, 0, 0))
    c3 = np.diag([0.04, 0.01, 0.002])
    th = 0.7
    g.AddToGraph((2.0, 1.0, th), c3)
    info = g.graph[1][1][0].information
    R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
    want = np.linalg.inv(R.T @ c3[:2, :2] @ R)                                            # world-frame xy covariance seen from the scan
    assert np.allclose(info[:2, :2], want, rtol=1e-12) and info[5, 5] == pytest.approx(1 / 0.002)


def test_default_registration_covariance_is_singular_like_the_reference():
    g = _drive(3, np.random.default_rng(2), sampled_cov=False)
    assert np.array_equal(G.DEFAULT_REG_COV, np.diag([0.1 * 0.1, 0.1 * 0.1, 0, 0, 0, 0.01 * 0.01]))
    assert np.all(np.isnan(g.graph[1][1][0].information))                                  # only usable with replace_cov_by_identity


def test_tbvg_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    g = _drive(9, rng)
    g.AddGroundTruth([s.stamp_ for s, _ in g.graph][::2], [np.array([i, 2.0 * i, 0.1 * i]) for i in range(5)])
    g.AddConstraint(G.Constraint3d(8, 1, G.pose3d_from_xyt((0.3, -0.2, 0.05)), 7 * np.eye(6), G.LOOP_APPEARANCE,
                                   {"Coral": 0.93, "CFEAR": 0.12, "odom-bounds": 0.5}, "guess 2, verified ✓"))
    g.AddConstraint(G.Constraint3d(5, 2, G.pose3d_from_xyt((0, 0, 0)), np.eye(6), G.CANDIDATE))
    p = str(tmp_path / "simple_graph.tbvg")
    G.save_simple_graph(p, g)
    h = G.load_simple_graph(p)
    assert len(h) == len(g)
    for (a, ca), (b, cb) in zip(g.graph, h.graph):
        for f in ("T", "Tgt", "motion_", "cloud_peaks_", "cloud_nopeaks_", "cloud_normal_", "downsampled_"):
            x, y = getattr(a, f), getattr(b, f)
            assert x.dtype == y.dtype and np.array_equal(x, y), f
        assert (a.has_Tgt_, a.idx_, a.stamp_, a.radius_, a.weight_intensity_) == (b.has_Tgt_, b.idx_, b.stamp_, b.radius_, b.weight_intensity_)
        assert len(ca) == len(cb)
        for c, d in zip(ca, cb):
            assert (c.id_begin, c.id_end, c.type, c.quality, c.info) == (d.id_begin, d.id_end, d.type, d.quality, d.info)
            assert np.array_equal(c.t_be, d.t_be) and np.array_equal(c.information, d.information)
    assert sum(s.has_Tgt_ for s, _ in h.graph) == 5 and h.graph[2][0].has_Tgt_ and not h.graph[1][0].has_Tgt_
    assert h.graph[0][0].cloud_peaks_.shape == (7, 4) and np.all(h.graph[0][0].cloud_peaks_[:, 2] == 0)      # [n,3] input = x y I
    # byte-stable: saving the loaded graph reproduces the file
    q = str(tmp_path / "again.tbvg")
    G.save_simple_graph(q, h)
    assert open(p, "rb").read() == open(q, "rb").read()
    # header as documented
    raw = open(p, "rb").read()
    assert raw[:4] == b"TBVG" and struct.unpack_from("<II", raw, 4) == (1, 9)
    assert np.array_equal(np.frombuffer(raw, "<f8", 7, 12), g.graph[0][0].T)


def test_tbvg_rejects_foreign_and_truncated_files(tmp_path):
    g = _drive(3, np.random.default_rng(4))
    p = str(tmp_path / "g.tbvg")
    G.save_simple_graph(p, g)
    raw = open(p, "rb").read()
    for bad in (b"22 serialization::archive 17" + raw[28:], raw[:len(raw) // 2], raw + b"\0", raw[:4] + struct.pack("<I", 2) + raw[8:]):
        q = str(tmp_path / "bad.tbvg")
        open(q, "wb").write(bad)
        with pytest.raises(ValueError):
            G.load_simple_graph(q)
    empty = G.SimpleGraph()
    G.save_simple_graph(p, empty)
    assert len(G.load_simple_graph(p)) == 0 and os.path.getsize(p) == 12


def test_pgo_arrays_feed_the_assembly():
    """The graph's own odometry constraints are satisfied by its own poses: the oracle's assembly gives zero cost and zero gradient, with
    identity weights and with the stored information matrices; a loop constraint that disagrees raises both."""
    from oracle import oracle_py as O
    O.lib()
    g = _drive(12, np.random.default_rng(5))
    g.AddConstraint(G.Constraint3d(4, 0, G.pose3d_from_xyt((0, 0, 0)), np.eye(6), G.MINI_LOOP))          # skipped by the optimiser
    nodes, ids, meas, info, idx = g.pgo_arrays()
    assert nodes.shape == (12, 7) and ids.shape == (11, 3) and meas.shape == (11, 7) and info.shape == (11, 36) and list(idx) == list(range(12))
    assert np.all(ids[:, 2] == 0) and np.array_equal(ids[:, 0], np.arange(1, 12)) and np.array_equal(ids[:, 1], np.arange(0, 11))
    c, Hd, Ho, grad, r = O.pgo_assemble(nodes, ids, meas)
    assert c <= 1e-24 and np.abs(grad).max() <= 1e-10 and np.abs(r).max() <= 1e-12
    c, *_ = O.pgo_assemble(nodes, ids, meas, O.default_pgo_params(replace_cov_by_identity=0), info=info)
    assert c <= 1e-20
    Tb, Te = g.graph[10][0].GetPose(), g.graph[1][0].GetPose()
    off = G.pose3d_to_matrix(G.pose3d_from_xyt((0.5, -0.3, 0.02)))
    g.AddConstraint(G.Constraint3d(10, 1, G.pose3d_from_matrix(np.linalg.inv(Tb) @ Te @ off), np.eye(6), G.LOOP_APPEARANCE))
    nodes, ids, meas, info, _ = g.pgo_arrays()
    assert len(ids) == 12 and sorted(ids[:, 2]) == [0] * 11 + [1]
    c2, _, _, g2, _ = O.pgo_assemble(nodes, ids, meas)
    assert c2 > 0 and np.abs(g2).max() > 0
    # writing optimised poses back
    moved = nodes.copy(); moved[:, 0] += 1.0
    g.set_poses(moved)
    assert np.allclose(g.poses_xyt()[:, 0], moved[:, 0]) and len(g.loopdb_sets()) == 12 and g.loopdb_sets()[3].shape[1] == 16


def test_to_string_is_the_graph_txt_line(tmp_path):
    g = _drive(4, np.random.default_rng(6))
    p = str(tmp_path / "graph.txt")
    TIO.write_graph_txt(p, [s.GetPose() for s, _ in g.graph], [s.stamp_ for s, _ in g.graph])
    text = open(p).read()
    for s, _ in g.graph:
        assert s.ToString().strip() in text
    assert [G.Constraint2String(t) for t in range(4)] == ["odometry", "loop_apperance", "loop_candidate", "loop_candidate"]   # sic
