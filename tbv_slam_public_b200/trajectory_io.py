"""Trajectory files either side of the path: the `est/NN.txt` / `gt/NN.txt` pose lists the reference's
evaluation reads.

  writer : EvalTrajectory::Write + MatToString (cfear_radarodometry/src/cfear_radarodometry/eval_trajectory.cpp:169-183,
           types.cpp:64-73): one pose per line, the row-major top 3x4 of the 4x4 matrix, `std::fixed` (6 decimals), single spaces.
  reader : KittiEvalOdom.load_poses_from_txt (radar_kitti_benchmark/python/kitti_odometry.py:93-121): 12 floats per line,
           optionally preceded by a frame index.
  metric : KittiEvalOdom.calc_sequence_errors / compute_overall_err (kitti_odometry.py:197-262): mean translational (fraction) and
           rotational (rad/m) drift over sub-sequences of 100 ... 800 m, the number the reference's papers quote.

  graph  : `graph.txt` (posegraph.cpp:177-187): the same 12 numbers + the node's time stamp, one empty line between nodes.

Host-side Python, no GPU: this is a file format, not a kernel.  `simple_graph.sgh` (a Boost binary archive) is out of scope.
"""
from __future__ import annotations

import math

import numpy as np

LENGTHS = (100, 200, 300, 400, 500, 600, 700, 800)   # kitti_odometry.py:44


def pose_matrix(xyt) -> np.ndarray:
    """(x, y, yaw) -> 4x4, as vectorToAffine3d builds it (registration.cpp:129-135: AngleAxis about z, z = 0)."""
    x, y, t = (float(v) for v in xyt)
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, 0.0, x], [s, c, 0.0, y], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


def mat_to_string(m: np.ndarray) -> str:
    """MatToString (types.cpp:64-73): `ss << std::fixed << std::showpoint` -> 6 decimals."""
    return " ".join("%.6f" % float(m[r, c]) for r in range(3) for c in range(4))


def write_kitti(path: str, poses) -> None:
    """poses: iterable of (x, y, yaw) or 4x4 matrices — EvalTrajectory::Write."""
    with open(path, "w") as f:
        for p in poses:
            m = np.asarray(p, float)
            f.write(mat_to_string(m if m.shape == (4, 4) else pose_matrix(m)) + "\n")


def write_graph_txt(path: str, poses, stamps_ns) -> None:
    """PoseGraph::SaveGraphString + RadarScan::ToString (tbv_slam/src/tbv_slam/posegraph.cpp:177-187,
    cfear_radarodometry/src/cfear_radarodometry/types.cpp:93-102): the 12 pose numbers, the node's stamp (std::to_string of the integer),
    and — ToString ends its line, SaveGraphString adds another — an empty line after every node."""
    with open(path, "w") as f:
        for p, t in zip(poses, stamps_ns):
            m = np.asarray(p, float)
            f.write(mat_to_string(m if m.shape == (4, 4) else pose_matrix(m)) + " " + str(int(t)) + "\n\n")


def read_graph_txt(path: str):
    """-> (list of 4x4 poses, list of stamps) from a graph.txt."""
    poses, stamps = [], []
    with open(path) as f:
        for line in f:
            v = line.strip().split(" ")
            if len(v) != 13:
                continue
            P = np.eye(4)
            P[:3, :4] = np.asarray([float(t) for t in v[:12]]).reshape(3, 4)
            poses.append(P)
            stamps.append(int(v[12]))
    return poses, stamps


def read_kitti(path: str) -> dict:
    """{frame index: 4x4} with the reference reader's rules (13 numbers -> the first is the index)."""
    poses = {}
    with open(path) as f:
        for cnt, line in enumerate(f.readlines()):
            v = [float(i) for i in line.strip().split(" ") if i != ""]
            if not v:
                continue
            off = 1 if len(v) == 13 else 0
            P = np.eye(4)
            P[:3, :4] = np.asarray(v[off:off + 12]).reshape(3, 4)
            poses[v[0] if off else cnt] = P
    return poses


def kitti_drift(poses_gt: dict, poses_est: dict, step_size: int = 10):
    """(mean translational drift [fraction], mean rotational drift [rad/m], n segments) — calc_sequence_errors + compute_overall_err."""
    keys = sorted(poses_gt.keys())
    dist = [0.0]
    for a, b in zip(keys[:-1], keys[1:]):
        d = poses_gt[a][:3, 3] - poses_gt[b][:3, 3]
        dist.append(dist[-1] + float(np.sqrt((d ** 2).sum())))
    t_err = r_err = 0.0
    n = 0
    for first in range(0, len(poses_gt), step_size):
        for length in LENGTHS:
            last = -1
            for i in range(first, len(dist)):
                if dist[i] > dist[first] + length:
                    last = i
                    break
            if last == -1 or last not in poses_est or first not in poses_est:
                continue
            d_gt = np.linalg.inv(poses_gt[first]) @ poses_gt[last]
            d_est = np.linalg.inv(poses_est[first]) @ poses_est[last]
            e = np.linalg.inv(d_est) @ d_gt
            rot = math.acos(max(min(0.5 * (e[0, 0] + e[1, 1] + e[2, 2] - 1.0), 1.0), -1.0))
            t_err += float(np.sqrt((e[:3, 3] ** 2).sum())) / length
            r_err += rot / length
            n += 1
    return (t_err / n, r_err / n, n) if n else (0.0, 0.0, 0)


# ---- the rest of the reference's odometry evaluation (radar_kitti_benchmark/python/kitti_odometry.py: eval(), `--align 6dof` as
# tbv_slam/script/run_eval.sh runs it): first-frame normalisation, Umeyama alignment, ATE, RPE -----------------------------------------
def umeyama_alignment(x: np.ndarray, y: np.ndarray, with_scale: bool = False):
    """kitti_odometry.py:32-84: (r, t, c) minimising |y - (c r x + t)|; x, y are m x n. Kabsch sign fix on the last axis."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    if x.shape != y.shape:
        raise ValueError("x.shape not equal to y.shape")
    m, n = x.shape
    mean_x, mean_y = x.mean(axis=1), y.mean(axis=1)
    sigma_x = 1.0 / n * (np.linalg.norm(x - mean_x[:, None]) ** 2)
    outer_sum = np.zeros((m, m))
    for i in range(n):                                   # same accumulation order as the reference's loop
        outer_sum += np.outer(y[:, i] - mean_y, x[:, i] - mean_x)
    cov_xy = np.multiply(1.0 / n, outer_sum)
    u, d, v = np.linalg.svd(cov_xy)
    s = np.eye(m)
    if np.linalg.det(u) * np.linalg.det(v) < 0.0:
        s[m - 1, m - 1] = -1
    r = u.dot(s).dot(v)
    c = 1 / sigma_x * np.trace(np.diag(d).dot(s)) if with_scale else 1.0
    t = mean_y - np.multiply(c, r.dot(mean_x))
    return r, t, c


def align_trajectories(poses_gt: dict, poses_est: dict, alignment: str | None = "6dof"):
    """eval() lines 708-737: both trajectories re-based on their first frame (the first key of the ESTIMATE), then the estimate aligned to
    the ground truth: None, "6dof" or "7dof" (Umeyama without / with scale). Returns new dicts; only keys of the estimate are touched."""
    idx_0 = sorted(poses_est.keys())[0]
    inv_e, inv_g = np.linalg.inv(poses_est[idx_0]), np.linalg.inv(poses_gt[idx_0])
    gt = dict(poses_gt)
    est = {}
    for k in poses_est:
        est[k] = inv_e @ poses_est[k]
        gt[k] = inv_g @ poses_gt[k]
    if alignment in ("6dof", "7dof"):
        xyz_gt = np.asarray([gt[k][:3, 3] for k in est]).transpose(1, 0)
        xyz_est = np.asarray([est[k][:3, 3] for k in est]).transpose(1, 0)
        r, t, scale = umeyama_alignment(xyz_est, xyz_gt, alignment != "6dof")
        A = np.eye(4)
        A[:3, :3], A[:3, 3] = r, t
        for k in est:
            e = est[k].copy()
            e[:3, 3] *= scale
            est[k] = A @ e
    elif alignment is not None:
        raise ValueError("alignment: None, '6dof' or '7dof'")
    return gt, est


def compute_ATE(gt: dict, pred: dict) -> float:
    """kitti_odometry.py:477-505: RMSE of the position differences over the estimate's keys."""
    errors = [np.sqrt(np.sum((gt[i][:3, 3] - pred[i][:3, 3]) ** 2)) for i in pred]
    return float(np.sqrt(np.mean(np.asarray(errors) ** 2)))


def compute_RPE(gt: dict, pred: dict) -> dict:
    """kitti_odometry.py:508-583 without the text dumps: frame-to-frame relative pose errors over consecutive keys of the estimate."""
    keys = list(pred.keys())[:-1]
    t_abs, t_sq, r_abs, ex, ey, ez = [], [], [], [], [], []
    for i in keys:
        gt_rel = np.linalg.inv(gt[i]) @ gt[i + 1]
        pred_rel = np.linalg.inv(pred[i]) @ pred[i + 1]
        e = np.linalg.inv(gt_rel) @ pred_rel
        ex.append(e[0, 3]); ey.append(e[1, 3])
        beta = -np.arcsin(e[2, 0])                                                      # rot2eul (:14-18) returns (alpha, beta, gamma) and the
        ez.append(np.arctan2(e[2, 1] / np.cos(beta), e[2, 2] / np.cos(beta)))            # reference keeps [0] = alpha (roll, not yaw) as "ez": kept
        sq = e[0, 3] ** 2 + e[1, 3] ** 2 + e[2, 3] ** 2
        t_abs.append(np.sqrt(sq)); t_sq.append(sq)
        r_abs.append(np.arccos(max(min(0.5 * (e[0, 0] + e[1, 1] + e[2, 2] - 1.0), 1.0), -1.0)))
    return {"rpe_trans": float(np.mean(t_abs)), "rpe_rot": float(np.mean(r_abs)), "rpe_trans_dev": float(np.std(t_abs)),
            "rpe_rot_dev": float(np.std(r_abs)), "bias_x": float(np.mean(ex)), "bias_y": float(np.mean(ey)),
            "bias_theta": float(np.mean(ez)), "rmse_trans": float(np.sqrt(np.mean(t_sq)))}


def evaluate(poses_gt: dict, poses_est: dict, alignment: str | None = "6dof", step_size: int = 10) -> dict:
    """One sequence of KittiEvalOdom.eval: alignment, drift over 100..800 m sub-sequences, ATE, RPE — the numbers of the reference's
    result.txt, in its units (`t_err_percent`, `r_err_deg_per_100m`, ATE in m, RPE rotation in rad as the reference stores it)."""
    gt, est = align_trajectories(poses_gt, poses_est, alignment)
    t, r, n = kitti_drift(gt, est, step_size)
    out = {"t_err_percent": 100.0 * t, "r_err_deg_per_100m": r / np.pi * 180 * 100, "n_segments": n, "ate": compute_ATE(gt, est)}
    out.update(compute_RPE(gt, est))
    return out
