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
