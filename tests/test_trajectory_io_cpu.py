"""est/NN.txt writer / reader / KITTI drift metric — checked against the reference's OWN reader and metric
(radar_kitti_benchmark/python/kitti_odometry.py, imported from /root/reference when present; it needs matplotlib only for plots, which
is stubbed) and against golden strings of the C++ writer's format."""
import os
import sys
import types

import numpy as np
import pytest

from tbv_slam_public_b200 import synth, trajectory_io as tio

REF_PY = "/root/reference/radar_kitti_benchmark/python"


def test_writer_format_matches_std_fixed():
    """MatToString (types.cpp:64-73): std::fixed -> exactly 6 decimals, 12 numbers, single spaces, negative zero kept as C++ prints it."""
    m = tio.pose_matrix((1.5, -2.25, 0.0))
    assert tio.mat_to_string(m) == "1.000000 -0.000000 0.000000 1.500000 0.000000 1.000000 0.000000 -2.250000 0.000000 0.000000 1.000000 0.000000"
    s = tio.mat_to_string(tio.pose_matrix((123.4567891, 0.0000004, np.pi / 2)))
    assert len(s.split(" ")) == 12 and all(len(t.split(".")[1]) == 6 for t in s.split(" "))


def test_roundtrip_and_drift_of_a_perfect_trajectory(tmp_path):
    gt = synth.make_stream(4).gt
    traj = synth.figure8(900)            # 900 poses at 2.5 m spacing: 2.2 km
    p = str(tmp_path / "00.txt")
    tio.write_kitti(p, traj)
    back = tio.read_kitti(p)
    assert len(back) == len(traj)
    for i, xyt in enumerate(traj):
        assert np.abs(back[i] - tio.pose_matrix(xyt)).max() < 5.1e-7   # 6 decimals
    t, r, n = tio.kitti_drift(back, back)
    assert n > 100 and t < 1e-9 and r < 1e-9
    assert len(gt) == 4


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="/root/reference not present (GPU box)")
def test_against_the_references_own_reader_and_metric(tmp_path):
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF_PY)
    try:
        from kitti_odometry import KittiEvalOdom
    finally:
        sys.path.remove(REF_PY)
    traj = synth.figure8(900)
    rng = np.random.default_rng(1)
    # an estimate with a slow yaw-rate bias and a scale error: drift of the order the reference reports (~1 %)
    est = []
    T = np.eye(4)
    for a, b in zip(traj[:-1], traj[1:]):
        d = np.linalg.inv(tio.pose_matrix(a)) @ tio.pose_matrix(b)
        d[:3, 3] *= 1.01
        yaw = np.arctan2(d[1, 0], d[0, 0]) + 2e-4 + rng.normal(0, 1e-4)
        d[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        est.append(T.copy())
        T = T @ d
    est.append(T.copy())
    gt_m = [np.linalg.inv(tio.pose_matrix(traj[0])) @ tio.pose_matrix(p) for p in traj]
    pg, pe = str(tmp_path / "gt.txt"), str(tmp_path / "est.txt")
    tio.write_kitti(pg, gt_m)
    tio.write_kitti(pe, est)
    tool = KittiEvalOdom(10)
    ref_gt, ref_est = tool.load_poses_from_txt(pg), tool.load_poses_from_txt(pe)     # the reference parses our files
    mine_gt, mine_est = tio.read_kitti(pg), tio.read_kitti(pe)
    assert sorted(ref_gt) == sorted(mine_gt)
    assert all(np.array_equal(ref_gt[k], mine_gt[k]) and np.array_equal(ref_est[k], mine_est[k]) for k in ref_gt)
    err = tool.calc_sequence_errors(ref_gt, ref_est)
    ave_t, ave_r = tool.compute_overall_err(err)
    t, r, n = tio.kitti_drift(mine_gt, mine_est)
    assert n == len(err) and n > 100
    assert abs(t - ave_t) < 1e-12 and abs(r - ave_r) < 1e-12
    assert 0.005 < t < 0.05


def test_graph_txt_roundtrip(tmp_path):
    traj = synth.figure8(20)
    stamps = [1547131046353776000 + 250000000 * i for i in range(20)]
    p = str(tmp_path / "graph.txt")
    tio.write_graph_txt(p, traj, stamps)
    raw = open(p).read().split("\n")
    assert raw[1] == "" and raw[0].endswith(" 1547131046353776000") and len(raw[0].split(" ")) == 13
    poses, st = tio.read_graph_txt(p)
    assert st == stamps and all(np.abs(P - tio.pose_matrix(x)).max() < 5.1e-7 for P, x in zip(poses, traj))


@pytest.mark.parametrize("alignment", [None, "6dof", "7dof"])
def test_evaluate_matches_the_references_numbers(tmp_path, alignment):
    """trajectory_io.evaluate against tests/golden/eval_ref.json — numbers the reference's own kitti_odometry.py produced for the same two
    pose files (tests/golden/make_eval_golden.py; the files are rewritten here from the same seed). Tolerance 1e-9 relative: the two
    implementations differ only in summation order."""
    import json
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    try:
        import make_eval_golden as M
    finally:
        sys.path.pop(0)
    gt, est = M.trajectories()
    pg, pe = str(tmp_path / "gt.txt"), str(tmp_path / "est.txt")
    tio.write_kitti(pg, gt)
    tio.write_kitti(pe, est)
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_ref.json")))[alignment or "none"]
    got = tio.evaluate(tio.read_kitti(pg), tio.read_kitti(pe), alignment)
    assert got["n_segments"] == want["n_segments"]
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-9 * max(abs(v), 1e-3), (k, got[k], v)
    if os.path.isdir(REF_PY):                                   # and live, where the reference is present
        live = M.reference_numbers(pg, pe, alignment)
        for k, v in live.items():
            assert abs(got[k] - v) <= 1e-9 * max(abs(v), 1e-3), (k, got[k], v)


def test_umeyama_recovers_a_known_similarity():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 40))
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    y = 1.3 * R @ x + np.array([[2.0], [-1.0], [0.5]])
    r, t, c = tio.umeyama_alignment(x, y, True)
    assert np.allclose(r, R, atol=1e-12) and np.allclose(t, [2.0, -1.0, 0.5], atol=1e-12) and abs(c - 1.3) < 1e-12
    r, t, c = tio.umeyama_alignment(x, R @ x, False)
    assert c == 1.0 and np.allclose(r, R, atol=1e-12) and np.allclose(t, 0, atol=1e-12)
    with pytest.raises(ValueError):
        tio.umeyama_alignment(x, y[:, :5])
