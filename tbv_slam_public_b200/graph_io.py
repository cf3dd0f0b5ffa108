"""The odometry -> loop-closure hand-off ("simple graph") in a documented, Boost-free layout.

The reference hands the precomputed odometry to the TBV back end as `simple_graph.sgh`: a Boost *binary* archive of
`std::vector<std::pair<RadarScan, std::vector<Constraint3d>>>` (cfear_radarodometry/include/cfear_radarodometry/types.h:93-190, 192;
SaveSimpleGraph / LoadSimpleGraph, cfear_radarodometry/src/cfear_radarodometry/types.cpp:103-130).  That format depends on the Boost
version, on sizeof(long) and on PCL's own serialisers, so it is not reproduced; this module keeps the same *content*, field for field, in
a little-endian layout that any language can read (`.tbvg`), and builds the graph the way OdometryKeyframeFuser::AddToGraph does
(odometrykeyframefuser.cpp:428-446).  A maintainer converts with a 20-line loop over the loaded `simple_graph`.

Layout of a .tbvg file (all little endian, no padding):

    char[4] "TBVG" | u32 version (1) | u32 n_nodes
    per node, in graph order:
        f64[7]  T        p.x p.y p.z q.x q.y q.z q.w     (Pose3d: `ar & p; ar & q`, types.h:76-80)
        f64[7]  Tgt
        u8      has_Tgt_
        u32     idx_
        u64     stamp_   (nanoseconds, `typedef unsigned long stamp`)
        f64[16] motion_  (4x4 row-major)
        cloud   cloud_peaks_      u32 n | f32[n][4] x y z intensity
        cloud   cloud_nopeaks_
        cells   cloud_normal_     u32 n | f64[n][16] tbv_cell records (include/tbv_b200.h: u cov scale snormal orth_normal lambda_min
                                  lambda_max sum_intensity avg_intensity n_samples — the reference serialises the same fields minus
                                  orth_normal, plus valid_, which is true for every stored cell) |
                                  u32 n_down | f32[n_down][2] downsampled_ | f32 radius_ | u8 weight_intensity_
        u32     n_constraints
        per constraint:
            u64 id_begin | u64 id_end | f64[7] t_be | f64[36] information (row-major) | u32 type (0 odometry, 1 loop_appearance,
            2 mini_loop, 3 candidate) | u32 n_quality | n_quality x (u16 len, bytes key, f64 value) | u32 len, bytes info

MapPointNormal::input_ (a second copy of the filtered cloud) is not stored: it is cloud_nopeaks_ (odometrykeyframefuser.cpp:161, 232).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC, VERSION = b"TBVG", 1
ODOMETRY, LOOP_APPEARANCE, MINI_LOOP, CANDIDATE = 0, 1, 2, 3          # ConstraintType (types.h:149)


def Constraint2String(t: int) -> str:
    """types.cpp:144-151, spelling included: everything that is neither odometry nor loop_appearance prints as a candidate."""
    t = int(t)
    return "odometry" if t == ODOMETRY else ("loop_apperance" if t == LOOP_APPEARANCE else "loop_candidate")


def pose3d_from_xyt(xyt) -> np.ndarray:
    """PoseEigToCeres of a planar pose: (p, q) = (x, y, 0, 0, 0, sin(t/2), cos(t/2))."""
    x, y, t = (float(v) for v in xyt)
    return np.array([x, y, 0.0, 0.0, 0.0, math.sin(t / 2), math.cos(t / 2)])


def pose3d_to_matrix(pq) -> np.ndarray:
    """Pose3d::GetPose: 4x4 from (p, q xyzw)."""
    px, py, pz, x, y, z, w = (float(v) for v in pq)
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = [px, py, pz]
    return m


def pose3d_from_matrix(m) -> np.ndarray:
    """Pose3d(const Affine3d&): p = translation, q = Quaterniond(rotation) (Eigen's Shepperd branches; w >= 0 on the trace branch)."""
    m = np.asarray(m, np.float64)
    r = m[:3, :3]
    tr = r[0, 0] + r[1, 1] + r[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        x, y, z = (r[2, 1] - r[1, 2]) * s, (r[0, 2] - r[2, 0]) * s, (r[1, 0] - r[0, 1]) * s
    else:
        i = 0
        if r[1, 1] > r[0, 0]:
            i = 1
        if r[2, 2] > r[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(r[i, i] - r[j, j] - r[k, k] + 1.0)
        q = [0.0, 0.0, 0.0]
        q[i] = 0.5 * s
        s = 0.5 / s
        w = (r[k, j] - r[j, k]) * s
        q[j] = (r[j, i] + r[i, j]) * s
        q[k] = (r[k, i] + r[i, k]) * s
        x, y, z = q
    return np.array([m[0, 3], m[1, 3], m[2, 3], x, y, z, w])


def pose3d_to_xyt(pq) -> np.ndarray:
    m = pose3d_to_matrix(pq)
    return np.array([m[0, 3], m[1, 3], math.atan2(m[1, 0], m[0, 0])])


@dataclass
class Constraint3d:
    """types.h:155-190."""
    id_begin: int
    id_end: int
    t_be: np.ndarray                                  # [7]
    information: np.ndarray                           # [6, 6]
    type: int = ODOMETRY
    quality: dict = field(default_factory=dict)
    info: str = ""


@dataclass
class RadarScan:
    """types.h:93-142 (serialised members only)."""
    T: np.ndarray
    idx_: int = 0
    stamp_: int = 0
    Tgt: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 0, 0, 0, 1.0]))
    has_Tgt_: bool = False
    motion_: np.ndarray = field(default_factory=lambda: np.eye(4))
    cloud_peaks_: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    cloud_nopeaks_: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    cloud_normal_: np.ndarray = field(default_factory=lambda: np.zeros((0, 16)))          # tbv_cell records
    downsampled_: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    radius_: float = 3.0
    weight_intensity_: bool = True

    def GetPose(self) -> np.ndarray:
        return pose3d_to_matrix(self.T)

    def ToString(self) -> str:
        """RadarScan::ToString (types.cpp:93-102)."""
        m = self.GetPose()
        return " ".join("%.6f" % float(m[r, c]) for r in range(3) for c in range(4)) + " " + str(int(self.stamp_)) + "\n"


def _cloud4(x, y=None, intensity=None) -> np.ndarray:
    if y is None:
        a = np.asarray(x, np.float32)
        if a.ndim == 2 and a.shape[1] == 4:
            return np.ascontiguousarray(a)
        if a.ndim == 2 and a.shape[1] == 3:            # x y intensity
            return np.ascontiguousarray(np.c_[a[:, 0], a[:, 1], np.zeros(len(a), np.float32), a[:, 2]].astype(np.float32))
        if a.size == 0:
            return np.zeros((0, 4), np.float32)
        raise ValueError("cloud: expected [n,4] (x y z I) or [n,3] (x y I)")
    x = np.asarray(x, np.float32)
    return np.ascontiguousarray(np.c_[x, np.asarray(y, np.float32), np.zeros(len(x), np.float32), np.asarray(intensity, np.float32)].astype(np.float32))


class SimpleGraph:
    """`simple_graph` (types.h:192) with OdometryKeyframeFuser's construction rules."""

    def __init__(self):
        self.graph: list[tuple[RadarScan, list[Constraint3d]]] = []

    def __len__(self):
        return len(self.graph)

    def AddToGraph(self, pose_xyt, cov=None, stamp_ns=0, motion_xyt=(0.0, 0.0, 0.0), cloud_peaks=None, cloud_nopeaks=None, cells=None,
                   radius=3.0, weight_intensity=True) -> RadarScan:
        """One keyframe: RadarScan(Tcurrent, Tmot, peaks, filtered, normals, t) + AddToGraph (odometrykeyframefuser.cpp:228-236, 428-446).

        The first frame has no constraint; every later one gets ONE odometry constraint to the newest reference keyframe:
        id_begin = this scan, id_end = the previous keyframe, t_be = Tfrom^-1 Tto, information = C^-1 with the translational block of the
        registration covariance rotated into this scan's frame.  cov: the 6x6 cov_current (DEFAULT_REG_COV after a plain Register, or
        the sampled one), or a 3x3 over (x, y, theta), embedded by `cov6_from_xyt`."""
        cov = DEFAULT_REG_COV if cov is None else np.asarray(cov, np.float64)
        if cov.shape == (3, 3):
            cov = cov6_from_xyt(cov)
        if cov.shape != (6, 6):
            raise ValueError("cov must be 6x6 or 3x3")
        cells = np.zeros((0, 16)) if cells is None else np.ascontiguousarray(cells, np.float64).reshape(-1, 16)
        mx, my, mt = (float(v) for v in motion_xyt)
        motion = np.eye(4)
        motion[:2, :2] = [[math.cos(mt), -math.sin(mt)], [math.sin(mt), math.cos(mt)]]
        motion[0, 3], motion[1, 3] = mx, my
        scan = RadarScan(T=pose3d_from_xyt(pose_xyt), idx_=len(self.graph), stamp_=int(stamp_ns), motion_=motion,
                         cloud_peaks_=_cloud4(np.zeros((0, 4)) if cloud_peaks is None else cloud_peaks),
                         cloud_nopeaks_=_cloud4(np.zeros((0, 4)) if cloud_nopeaks is None else cloud_nopeaks),
                         cloud_normal_=cells, downsampled_=np.ascontiguousarray(cells[:, :2], np.float32), radius_=float(radius),
                         weight_intensity_=bool(weight_intensity))
        constraints = []
        if self.graph:
            prev = self.graph[-1][0]
            Tfrom, Tto = scan.GetPose(), prev.GetPose()
            Tinv = np.linalg.inv(Tfrom)
            Tdiff = Tinv @ Tto
            Cm = cov.copy()
            Cm[:3, :3] = Tinv[:3, :3] @ cov[:3, :3] @ Tinv[:3, :3].T
            constraints.append(Constraint3d(scan.idx_, prev.idx_, pose3d_from_matrix(Tdiff), _information(Cm), ODOMETRY))
        self.graph.append((scan, constraints))
        return scan

    def AddGroundTruth(self, stamps_ns, poses) -> int:
        """OdometryKeyframeFuser::AddGroundTruth (odometrykeyframefuser.cpp:447-463): exact stamp match. poses: 4x4 or (x, y, theta)."""
        table = {}
        for s, p in zip(stamps_ns, poses):
            p = np.asarray(p, np.float64)
            table[int(s)] = pose3d_from_matrix(p) if p.shape == (4, 4) else pose3d_from_xyt(p)
        hit = 0
        for scan, _ in self.graph:
            if scan.stamp_ in table:
                scan.Tgt, scan.has_Tgt_ = table[scan.stamp_], True
                hit += 1
        return hit

    def AddConstraint(self, c: Constraint3d):
        """Attach a (loop) constraint to its id_begin node, as PoseGraph keeps them per type (posegraph.cpp AddConstraint)."""
        for scan, cons in self.graph:
            if scan.idx_ == c.id_begin:
                cons.append(c)
                return
        raise KeyError("no node with idx_ %d" % c.id_begin)

    # ---- views for the device calls ------------------------------------------------------------------------------------------------
    def pgo_arrays(self):
        """-> (nodes [n,7], ids [m,3] int32 (row_begin, row_end, type: 0 odometry / 1 loop), meas [m,7], info [m,36], idx_ of every row):
        the arguments of tbv_pgo_assemble.  Constraint types other than odometry / loop_appearance are skipped, as
        CeresLeastSquares::BuildOptimizationProblem adds only those two (ceresoptimizer.cpp:34-35)."""
        row = {scan.idx_: i for i, (scan, _) in enumerate(self.graph)}
        nodes = np.array([scan.T for scan, _ in self.graph], np.float64).reshape(-1, 7)
        ids, meas, info = [], [], []
        for _, cons in self.graph:
            for c in cons:
                if c.type not in (ODOMETRY, LOOP_APPEARANCE):
                    continue
                ids.append((row[c.id_begin], row[c.id_end], 0 if c.type == ODOMETRY else 1))
                meas.append(c.t_be)
                info.append(np.asarray(c.information, np.float64).reshape(36))
        return (nodes, np.array(ids, np.int32).reshape(-1, 3), np.array(meas, np.float64).reshape(-1, 7),
                np.array(info, np.float64).reshape(-1, 36), np.array([scan.idx_ for scan, _ in self.graph], np.int64))

    def set_poses(self, nodes):
        """Write optimised parameter blocks back (the reference optimises RadarScan::T in place)."""
        nodes = np.asarray(nodes, np.float64).reshape(-1, 7)
        if len(nodes) != len(self.graph):
            raise ValueError("one pose per node expected")
        for (scan, _), pq in zip(self.graph, nodes):
            scan.T = pq.copy()

    def loopdb_sets(self):
        """Per-node cell arrays in graph order: the argument of LoopDB.add (tbv_loopdb_add)."""
        return [scan.cloud_normal_ for scan, _ in self.graph]

    def poses_xyt(self) -> np.ndarray:
        return np.array([pose3d_to_xyt(scan.T) for scan, _ in self.graph]).reshape(-1, 3)


DEFAULT_REG_COV = np.diag([0.1 * 0.1, 0.1 * 0.1, 0.0, 0.0, 0.0, 0.01 * 0.01])   # what Register leaves in reg_cov (n_scan_normal.cpp:171-175)


def cov6_from_xyt(c3) -> np.ndarray:
    """A planar (x, y, theta) covariance as the 6x6 the reference carries (order x y z ex ey ez), embedded the way
    approximateCovarianceBySampling does (odometrykeyframefuser.cpp:367-374): Identity with the xy block, (5,5) and the x/y-theta terms."""
    c3 = np.asarray(c3, np.float64)
    c = np.eye(6)
    c[:2, :2] = c3[:2, :2]
    c[5, 5] = c3[2, 2]
    c[0, 5], c[1, 5], c[5, 0], c[5, 1] = c3[0, 2], c3[1, 2], c3[2, 0], c3[2, 1]
    return c


def _information(cov6) -> np.ndarray:
    """C.inverse() (odometrykeyframefuser.cpp:440).  The registration's default covariance (DEFAULT_REG_COV) is singular: the reference then
    stores a non-finite matrix that only works because replace_cov_by_identity ignores it; here it becomes all-NaN, explicitly."""
    try:
        return np.linalg.inv(cov6)
    except np.linalg.LinAlgError:
        return np.full((6, 6), np.nan)


# ---- .tbvg reader / writer ---------------------------------------------------------------------------------------------------------
def _w_cloud(out, a, width, dtype):
    a = np.ascontiguousarray(a, dtype).reshape(-1, width)
    out.append(struct.pack("<I", len(a)))
    out.append(a.astype("<" + np.dtype(dtype).str[1:]).tobytes())


def save_simple_graph(path: str, graph: SimpleGraph) -> None:
    """SaveSimpleGraph (types.cpp:103-114) into the .tbvg layout of this module's docstring."""
    out = [MAGIC, struct.pack("<II", VERSION, len(graph.graph))]
    for scan, cons in graph.graph:
        out.append(np.asarray(scan.T, "<f8").reshape(7).tobytes())
        out.append(np.asarray(scan.Tgt, "<f8").reshape(7).tobytes())
        out.append(struct.pack("<BIQ", 1 if scan.has_Tgt_ else 0, int(scan.idx_), int(scan.stamp_)))
        out.append(np.asarray(scan.motion_, "<f8").reshape(16).tobytes())
        _w_cloud(out, scan.cloud_peaks_, 4, np.float32)
        _w_cloud(out, scan.cloud_nopeaks_, 4, np.float32)
        _w_cloud(out, scan.cloud_normal_, 16, np.float64)
        _w_cloud(out, scan.downsampled_, 2, np.float32)
        out.append(struct.pack("<fB", float(scan.radius_), 1 if scan.weight_intensity_ else 0))
        out.append(struct.pack("<I", len(cons)))
        for c in cons:
            out.append(struct.pack("<QQ", int(c.id_begin), int(c.id_end)))
            out.append(np.asarray(c.t_be, "<f8").reshape(7).tobytes())
            out.append(np.asarray(c.information, "<f8").reshape(36).tobytes())
            out.append(struct.pack("<II", int(c.type), len(c.quality)))
            for k in sorted(c.quality):                              # std::map order
                kb = k.encode("utf-8")
                out.append(struct.pack("<H", len(kb)) + kb + struct.pack("<d", float(c.quality[k])))
            ib = c.info.encode("utf-8")
            out.append(struct.pack("<I", len(ib)) + ib)
    with open(path, "wb") as f:
        f.write(b"".join(out))


class _Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.o + n > len(self.b):
            raise ValueError("truncated .tbvg file")
        v = struct.unpack_from(fmt, self.b, self.o)
        self.o += n
        return v

    def arr(self, count, dtype):
        n = count * np.dtype(dtype).itemsize
        if self.o + n > len(self.b):
            raise ValueError("truncated .tbvg file")
        a = np.frombuffer(self.b, dtype=np.dtype(dtype).newbyteorder("<"), count=count, offset=self.o).astype(dtype)
        self.o += n
        return a

    def raw(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated .tbvg file")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


def load_simple_graph(path: str) -> SimpleGraph:
    """LoadSimpleGraph (types.cpp:116-130).  Raises ValueError on a foreign or truncated file (the reference returns false)."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    if r.raw(4) != MAGIC:
        raise ValueError("not a .tbvg file")
    version, n_nodes = r.take("<II")
    if version != VERSION:
        raise ValueError("unsupported .tbvg version %d" % version)
    g = SimpleGraph()
    for _ in range(n_nodes):
        T, Tgt = r.arr(7, np.float64), r.arr(7, np.float64)
        has_gt, idx, stamp = r.take("<BIQ")
        motion = r.arr(16, np.float64).reshape(4, 4)
        (n,) = r.take("<I"); peaks = r.arr(4 * n, np.float32).reshape(n, 4)
        (n,) = r.take("<I"); nopeaks = r.arr(4 * n, np.float32).reshape(n, 4)
        (n,) = r.take("<I"); cells = r.arr(16 * n, np.float64).reshape(n, 16)
        (n,) = r.take("<I"); down = r.arr(2 * n, np.float32).reshape(n, 2)
        radius, wi = r.take("<fB")
        scan = RadarScan(T=T, idx_=idx, stamp_=stamp, Tgt=Tgt, has_Tgt_=bool(has_gt), motion_=motion, cloud_peaks_=peaks, cloud_nopeaks_=nopeaks,
                         cloud_normal_=cells, downsampled_=down, radius_=radius, weight_intensity_=bool(wi))
        (nc,) = r.take("<I")
        cons = []
        for _c in range(nc):
            a, b = r.take("<QQ")
            t_be, information = r.arr(7, np.float64), r.arr(36, np.float64).reshape(6, 6)
            ctype, nq = r.take("<II")
            quality = {}
            for _q in range(nq):
                (kl,) = r.take("<H")
                k = r.raw(kl).decode("utf-8")
                (quality[k],) = r.take("<d")
            (il,) = r.take("<I")
            cons.append(Constraint3d(a, b, t_be, information, ctype, quality, r.raw(il).decode("utf-8")))
        g.graph.append((scan, cons))
    if r.o != len(r.b):
        raise ValueError("trailing bytes in .tbvg file")
    return g
