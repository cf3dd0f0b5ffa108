"""The caller on the near side of the hot path: the offline odometry loop that turns a stream of polar scans into a trajectory and the
simple graph (cfear_radarodometry/src/offline_odometry.cpp:57-141 `radarReader`; the same loop feeds the graph in
tbv_slam/src/tbv_slam_online.cpp:107-227).

Per scan: radarDriver::CallbackOffline (k-strongest filter, both clouds) -> OdometryKeyframeFuser::pointcloudCallback (compensation, surface
points, registration against the keyframe window, keyframe decision) -> EvalTrajectory::CallbackESTEigen; every keyframe becomes a graph
node carrying its compensated clouds and its cells, with one odometry constraint to the previous keyframe
(odometrykeyframefuser.cpp:169-175, 226-236, 428-446).  Save(): est / gt trajectories, ground truth into the graph, the graph itself
(offline_odometry.cpp:134-141).

The frame itself — filter, cells, registration — is ONE device step (tbv_odom_step: K1, K2, cells_fused, k_register, k_odom_update); the
clouds a keyframe stores are produced by the same filter + compensation entry points.  The device is a constructor argument so that this
bookkeeping can be exercised without a GPU (tests/test_offline_odometry_cpu.py plugs the CPU oracle in); the product device needs a GPU."""
from __future__ import annotations

import math
import os

import numpy as np

from . import graph_io as G
from . import trajectory_io as TIO


def _mul(a, b):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([a[0] + ca * b[0] - sa * b[1], a[1] + sa * b[0] + ca * b[1], a[2] + b[2]])


def _inv(a):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([-(ca * a[0] + sa * a[1]), -(-sa * a[0] + ca * a[1]), -a[2]])


class GpuOdometryDevice:
    """radarDriver + OdometryKeyframeFuser of ONE sequence on the GPU (api.OdometryKeyframeFuser with n_seq = 1)."""

    def __init__(self, ctx, n_az: int, n_range: int, params=None):
        from . import api
        self.api, self.ctx = api, ctx
        self.params = params or api.default_odom_params()
        self.fuser = api.OdometryKeyframeFuser(ctx, 1, n_az, n_range, self.params)

    def close(self):
        self.fuser.close()

    def step(self, scan: np.ndarray) -> dict:
        o = self.fuser.pointcloudCallback(np.ascontiguousarray(scan, np.uint8)[None])[0]
        if o.status != 0:
            raise self.api.TbvError(o.status, "a per-scan capacity was exceeded (tbv_odom_params.cell_capacity / sample_capacity)")
        return dict(pose=np.array([o.pose[0], o.pose[1], o.pose[2]]), is_keyframe=bool(o.is_keyframe), n_keyframes=int(o.n_keyframes),
                    n_cells=int(o.n_cells), score=float(o.score), itrs=int(o.itrs))

    def newest_keyframe_cells(self, n_keyframes: int) -> np.ndarray:
        cells, _ = self.fuser.cells(0, n_keyframes - 1)
        return cells

    def clouds(self, scan: np.ndarray, motion_xyt):
        """The two clouds a RadarScan stores: k-strongest filtered and its peaks, motion-compensated with the motion used for this frame."""
        f = self.params.filter
        filt, peaks = self.ctx.StructuredKStrongest(scan, f.z_min, f.k_strongest, f.min_distance, f.range_res, peaks=True)
        out = []
        for buf in (filt, peaks):
            _, _, I, x, y = buf.scan(0)
            if self.params.compensate and len(x):
                x, y = self.ctx.Compensate(x, y, motion_xyt, bool(self.params.radar_ccw))
            out.append(np.c_[x, y, np.zeros(len(x), np.float32), I.astype(np.float32)].astype(np.float32).reshape(-1, 4))
        return out[0], out[1]


class radarReader:
    """offline_odometry.cpp:57-141.  `stamps_ns` / `gt` (optional, (x, y, yaw) or 4x4 per scan) play the rosbag's /Navtech/Polar header stamps
    and /gt messages."""

    def __init__(self, device, radius: float = 3.0, weight_intensity: bool = True):
        self.dev, self.radius, self.weight_intensity = device, radius, weight_intensity
        self.graph = G.SimpleGraph()
        self.est, self.gt, self.stamps = [], [], []                   # EvalTrajectory: one pose per processed scan
        self.poses = []                                               # Tcurrent history
        self.keyframe_rows = []                                       # scan index of every graph node

    def process(self, scan: np.ndarray, stamp_ns: int = 0, gt=None) -> dict:
        k = len(self.poses)
        # the motion that compensates THIS frame is the previous frame-to-frame motion (TprevMot, odometrykeyframefuser.cpp:146-150)
        motion = _mul(_inv(self.poses[-2]), self.poses[-1]) if k >= 2 else np.zeros(3)
        out = self.dev.step(scan)
        pose = out["pose"]
        self.poses.append(pose)
        self.est.append(pose)
        self.stamps.append(int(stamp_ns))
        if gt is not None:
            self.gt.append(np.asarray(gt, np.float64))
        if out["is_keyframe"]:
            filtered, peaks = self.dev.clouds(scan, motion)
            cells = self.dev.newest_keyframe_cells(out["n_keyframes"])
            first = len(self.graph) == 0
            # first node: RadarScan(Identity, Identity, ...) with Identity covariance and no constraint (:169-175); later: cov_current (:245)
            self.graph.AddToGraph(pose, np.eye(6) if first else None, stamp_ns=stamp_ns, motion_xyt=np.zeros(3) if first else motion,
                                  cloud_peaks=peaks, cloud_nopeaks=filtered, cells=cells, radius=self.radius, weight_intensity=self.weight_intensity)
            self.keyframe_rows.append(k)
        return out

    def run(self, scans, stamps_ns=None, gt=None):
        for i, scan in enumerate(scans):
            # without stamps: the 4 Hz sensor clock (Tsensor = 0.25 s, odometrykeyframefuser.h:213), so that stamps identify frames
            self.process(scan, 250_000_000 * (len(self.poses) + 1) if stamps_ns is None else int(stamps_ns[i]), None if gt is None else gt[i])
        return self

    def Save(self, directory: str, sequence: str = "00") -> dict:
        """eval.Save + fuser.AddGroundTruth + fuser.SaveGraph: <dir>/est/<seq>.txt, <dir>/gt/<seq>.txt, <dir>/simple_graph.tbvg."""
        os.makedirs(os.path.join(directory, "est"), exist_ok=True)
        paths = {"est": os.path.join(directory, "est", sequence + ".txt"), "graph": os.path.join(directory, "simple_graph.tbvg")}
        TIO.write_kitti(paths["est"], self.est)
        if self.gt:
            os.makedirs(os.path.join(directory, "gt"), exist_ok=True)
            paths["gt"] = os.path.join(directory, "gt", sequence + ".txt")
            TIO.write_kitti(paths["gt"], self.gt)
            self.graph.AddGroundTruth(self.stamps, self.gt)
        G.save_simple_graph(paths["graph"], self.graph)
        return paths
