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


def _wrap(t: float) -> float:
    """Yaw in (-pi, pi], as Affine3dToVectorXYeZ's atan2 yields it (utils.cpp:115-122) and as the device stores poses (k_odom.cu)."""
    return math.atan2(math.sin(t), math.cos(t))


def _mul(a, b):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([a[0] + ca * b[0] - sa * b[1], a[1] + sa * b[0] + ca * b[1], _wrap(a[2] + b[2])])


def _inv(a):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([-(ca * a[0] + sa * a[1]), -(-sa * a[0] + ca * a[1]), _wrap(-a[2])])


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


# ---- the fuser with the reference's OWN interface: point clouds in ------------------------------------------------------------------------
# OdometryKeyframeFuser::pointcloudCallback(cloud_filtered, cloud_filtered_peaks, Tcurr, t) (odometrykeyframefuser.cpp:395-411) takes CLOUDS,
# whatever filter produced them (k-strongest, CA-CFAR: radar_driver.cpp:48-62, or a caller's own).  tbv_odom_step fuses the k-strongest
# filter into the frame and is the fast path; this class is the general one: the same processFrame (:143-259) as host bookkeeping over
# the primitive device calls tbv_compensate, tbv_build_cells and tbv_register — one launch each per frame.
class GpuPrimitiveDevice:
    def __init__(self, ctx):
        self.ctx = ctx

    def compensate(self, x, y, mot, ccw):
        return self.ctx.Compensate(x, y, mot, ccw)

    def build_cells(self, x, y, intensity, radius, downsample_factor, weight_intensity):
        cells, _ = self.ctx.MapPointNormal(x, y, intensity, radius, downsample_factor, weight_intensity)
        return cells

    def register(self, scans, T, reg_params):
        Tio, s = self.ctx.Register(scans, T, reg_params)
        return Tio, bool(s.success), int(s.itrs), float(s.score)


class _Affine:
    """Planar Eigen::Affine3d as the reference multiplies it (2x2 linear part + translation), operation order of Eigen's products."""
    __slots__ = ("r00", "r01", "r10", "r11", "tx", "ty")

    def __init__(self, r00=1.0, r01=0.0, r10=0.0, r11=1.0, tx=0.0, ty=0.0):
        self.r00, self.r01, self.r10, self.r11, self.tx, self.ty = r00, r01, r10, r11, tx, ty

    @staticmethod
    def from_xyt(v):                                                   # vectorToAffine3d (registration.cpp:129-135)
        c, s = math.cos(v[2]), math.sin(v[2])
        return _Affine(c, -s, s, c, float(v[0]), float(v[1]))

    def xyt(self):                                                     # Affine3dToVectorXYeZ: yaw from the rotation's second column pair
        return np.array([self.tx, self.ty, math.atan2(self.r10, self.r11)])

    def __matmul__(self, b):
        return _Affine(self.r00 * b.r00 + self.r01 * b.r10, self.r00 * b.r01 + self.r01 * b.r11,
                       self.r10 * b.r00 + self.r11 * b.r10, self.r10 * b.r01 + self.r11 * b.r11,
                       (self.r00 * b.tx + self.r01 * b.ty) + self.tx, (self.r10 * b.tx + self.r11 * b.ty) + self.ty)

    def inverse(self):                                                 # Eigen's general affine inverse: cofactors / determinant
        inv = 1.0 / (self.r00 * self.r11 - self.r01 * self.r10)
        a, b, c, d = self.r11 * inv, -self.r01 * inv, -self.r10 * inv, self.r00 * inv
        return _Affine(a, b, c, d, -(a * self.tx + b * self.ty), -(c * self.tx + d * self.ty))


class OdometryKeyframeFuser:
    """processFrame (odometrykeyframefuser.cpp:143-259) over clouds.  params: api.OdomParams (its filter member is unused here)."""

    def __init__(self, device, params):
        self.dev, self.par = device, params
        self.keyframes_ = []                                           # (pose _Affine, cells)
        self.Tcurrent, self.T_prev, self.Tmot = _Affine(), _Affine(), _Affine()
        self.updated, self.last_reg_ok, self.last_itrs, self.last_cells = False, True, 0, None

    @staticmethod
    def KeyFrameBasedFuse(diff: _Affine, use_keyframe, min_keyframe_dist, min_keyframe_rot_deg) -> bool:       # :62-73
        if not use_keyframe:
            return True
        return math.sqrt(diff.tx * diff.tx + diff.ty * diff.ty) > min_keyframe_dist or \
            abs(math.atan2(diff.r10, diff.r11)) > (min_keyframe_rot_deg * math.pi / 180.0)

    @staticmethod
    def AccelerationVelocitySanityCheck(prev: _Affine, cur: _Affine) -> bool:                                    # :76-94
        dt, vel_limit, acc_limit = 0.25, 200.0, 200.0
        vel = math.sqrt((cur.tx / dt) ** 2 + (cur.ty / dt) ** 2)
        acc = math.sqrt(((cur.tx - prev.tx) / (dt * dt)) ** 2 + ((cur.ty - prev.ty) / (dt * dt)) ** 2)
        return not (acc > acc_limit or vel > vel_limit)

    def pointcloudCallback(self, x, y, intensity, peaks=None):
        """-> (pose (x, y, yaw), compensated cloud (x, y), compensated peaks or None).  peaks: (x, y, intensity) of the peaks cloud."""
        p = self.par
        TprevMot = self.Tmot
        x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
        if p.compensate:
            mot = TprevMot.xyt()
            if len(x):
                x, y = self.dev.compensate(x, y, mot, bool(p.radar_ccw))
            if peaks is not None and len(peaks[0]):
                px, py = self.dev.compensate(peaks[0], peaks[1], mot, bool(p.radar_ccw))
                peaks = (px, py, peaks[2])
        cells = self.dev.build_cells(x, y, intensity, float(p.res), float(p.downsample_factor), bool(p.weight_intensity))
        self.last_cells = cells
        Tguess = (self.T_prev @ TprevMot) if p.use_guess else self.T_prev
        self.updated, self.last_reg_ok, self.last_itrs = False, True, 0
        if not self.keyframes_:
            self.keyframes_.append((_Affine(), cells))
            self.updated = True
            return self.Tcurrent.xyt(), (x, y), peaks
        scans = [c for _, c in self.keyframes_] + [cells]
        T = np.array([k.xyt() for k, _ in self.keyframes_] + [Tguess.xyt()])
        Tio, ok, itrs, _score = self.dev.register(scans, T, p.reg)     # the reference ignores `ok` (shadowed variable, :184-193)
        self.last_reg_ok, self.last_itrs = ok, itrs
        self.Tcurrent = _Affine.from_xyt(Tio[-1])
        Tmot_current = self.T_prev.inverse() @ self.Tcurrent
        if not self.AccelerationVelocitySanityCheck(self.Tmot, Tmot_current):
            self.Tcurrent = Tguess
        self.Tmot = self.T_prev.inverse() @ self.Tcurrent
        Tkeydiff = self.keyframes_[-1][0].inverse() @ self.Tcurrent
        if self.KeyFrameBasedFuse(Tkeydiff, bool(p.use_keyframe), p.min_keyframe_dist, p.min_keyframe_rot_deg):
            self.keyframes_.append((self.Tcurrent, cells))
            if len(self.keyframes_) > p.submap_scan_size:
                self.keyframes_.pop(0)
            self.updated = True
        self.T_prev = self.Tcurrent
        return self.Tcurrent.xyt(), (x, y), peaks
