"""Golden numbers for trajectory_io.evaluate, produced by the reference's OWN evaluation code
(/root/reference/radar_kitti_benchmark/python/kitti_odometry.py: load_poses_from_txt, the alignment block of eval(), calc_sequence_errors,
compute_overall_err, compute_ATE, compute_RPE) on a seeded synthetic trajectory pair.  Run in the build container:
    python tests/golden/make_eval_golden.py       -> tests/golden/eval_ref.json
matplotlib is only needed by the reference's plotting and is stubbed."""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tbv_slam_public_b200 import synth, trajectory_io as tio  # noqa: E402

REF_PY = "/root/reference/radar_kitti_benchmark/python"


def trajectories(seed=3, n=700):
    traj = synth.figure8(n)
    rng = np.random.default_rng(seed)
    est, T = [], tio.pose_matrix((3.0, -2.0, 0.4))         # the estimate starts somewhere else: alignment has work to do
    for a, b in zip(traj[:-1], traj[1:]):
        d = np.linalg.inv(tio.pose_matrix(a)) @ tio.pose_matrix(b)
        d[:3, 3] *= 1.008
        d[0, 3] += rng.normal(0, 0.01); d[1, 3] += rng.normal(0, 0.01)
        yaw = np.arctan2(d[1, 0], d[0, 0]) + 1.5e-4 + rng.normal(0, 2e-4)
        d[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        est.append(T.copy())
        T = T @ d
    est.append(T.copy())
    return [tio.pose_matrix(p) for p in traj], est


def reference_numbers(pg, pe, alignment):
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF_PY)
    try:
        import kitti_odometry as K
    finally:
        sys.path.remove(REF_PY)
    tool = K.KittiEvalOdom(10)
    poses_result, poses_gt = tool.load_poses_from_txt(pe), tool.load_poses_from_txt(pg)
    # eval() lines 708-737, verbatim in effect (the function itself also plots and writes files)
    idx_0 = sorted(list(poses_result.keys()))[0]
    pred_0, gt_0 = poses_result[idx_0], poses_gt[idx_0]
    for cnt in poses_result:
        poses_result[cnt] = np.linalg.inv(pred_0) @ poses_result[cnt]
        poses_gt[cnt] = np.linalg.inv(gt_0) @ poses_gt[cnt]
    if alignment in ("6dof", "7dof"):
        xyz_gt = np.asarray([[poses_gt[c][0, 3], poses_gt[c][1, 3], poses_gt[c][2, 3]] for c in poses_result]).transpose(1, 0)
        xyz_result = np.asarray([[poses_result[c][0, 3], poses_result[c][1, 3], poses_result[c][2, 3]] for c in poses_result]).transpose(1, 0)
        r, t, scale = K.umeyama_alignment(xyz_result, xyz_gt, alignment != "6dof")
        A = np.eye(4); A[:3, :3] = r; A[:3, 3] = t
        for cnt in poses_result:
            poses_result[cnt][:3, 3] *= scale
            poses_result[cnt] = A @ poses_result[cnt]
    seq_err = tool.calc_sequence_errors(poses_gt, poses_result)
    ave_t, ave_r = tool.compute_overall_err(seq_err)
    ate = tool.compute_ATE(poses_gt, poses_result)
    with tempfile.TemporaryDirectory() as d:
        rpe_trans, rpe_rot, rpe_trans_dev, rpe_rot_dev, bias_x, bias_y, bias_theta, rmse = tool.compute_RPE(poses_gt, poses_result, d)
    return {"t_err_percent": ave_t * 100, "r_err_deg_per_100m": ave_r / np.pi * 180 * 100, "n_segments": len(seq_err), "ate": float(ate),
            "rpe_trans": float(rpe_trans), "rpe_rot": float(rpe_rot), "rpe_trans_dev": float(rpe_trans_dev), "rpe_rot_dev": float(rpe_rot_dev),
            "bias_x": float(bias_x), "bias_y": float(bias_y), "bias_theta": float(bias_theta), "rmse_trans": float(rmse)}


def main():
    gt, est = trajectories()
    with tempfile.TemporaryDirectory() as d:       # the pose files are a pure function of the seed: the test rewrites them, they are not committed
        pg, pe = os.path.join(d, "eval_gt.txt"), os.path.join(d, "eval_est.txt")
        tio.write_kitti(pg, gt)
        tio.write_kitti(pe, est)
        out = {a or "none": reference_numbers(pg, pe, a) for a in (None, "6dof", "7dof")}
    with open(os.path.join(HERE, "eval_ref.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
