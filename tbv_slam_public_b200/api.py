"""ctypes binding of libtbv_b200.so — the Python face of the C-ABI in include/tbv_b200.h.

Class and method names follow the reference's C++ seams (StructuredKStrongest, MapPointNormal, n_scan_normal_reg,
OdometryKeyframeFuser) so the parity tests read like tests of the reference classes.  There is no CPU fallback:
a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtbv_b200.so")
_LIB = None

P2P, P2L, P2D = 0, 1, 2
LOSS_NONE, HUBER, CAUCHY, SOFTLONE, COMBINED, TUKEY = 0, 1, 2, 3, 4, 5
W_UNIFORM, W_SIM_N, W_SIM_DIR, W_SIM_SCALE, W_COMBINED = 0, 1, 2, 3, 4
TBV_OK, TBV_ERR_INVALID, TBV_ERR_CUDA, TBV_ERR_CAPACITY, TBV_ERR_NO_GPU = 0, -1, -2, -3, -4


class TbvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tbv error {code}: {msg}")
        self.code = code


class FilterParams(C.Structure):
    _fields_ = [("z_min", C.c_float), ("k_strongest", C.c_int), ("min_distance", C.c_float), ("range_res", C.c_float)]


class CfarParams(C.Structure):
    _fields_ = [("window_size", C.c_int), ("false_alarm_rate", C.c_double), ("nb_guard_cells", C.c_int),
                ("range_resolution", C.c_double), ("static_threshold", C.c_double), ("min_distance", C.c_double),
                ("max_distance", C.c_double)]


class Points(C.Structure):
    _fields_ = [("capacity", C.c_int), ("count", C.POINTER(C.c_int)), ("azimuth", C.POINTER(C.c_uint16)),
                ("range", C.POINTER(C.c_uint16)), ("intensity", C.POINTER(C.c_uint8)), ("x", C.POINTER(C.c_float)),
                ("y", C.POINTER(C.c_float))]


class RegParams(C.Structure):
    _fields_ = [("cost", C.c_int), ("loss", C.c_int), ("weight_opt", C.c_int), ("loss_limit", C.c_double),
                ("cov_scale", C.c_double), ("regularization", C.c_double), ("max_itr_association", C.c_int),
                ("max_itr_solver", C.c_int)]


class RegSummary(C.Structure):
    _fields_ = [("success", C.c_int), ("itrs", C.c_int), ("lm_iterations", C.c_int), ("num_residuals", C.c_int),
                ("last_n_iterations", C.c_int), ("termination", C.c_int), ("score", C.c_double), ("final_cost", C.c_double),
                ("last_relative_decrease", C.c_double)]


class OdomParams(C.Structure):
    _fields_ = [("filter", FilterParams), ("reg", RegParams), ("submap_scan_size", C.c_int), ("weight_intensity", C.c_int),
                ("use_guess", C.c_int), ("compensate", C.c_int), ("radar_ccw", C.c_int), ("use_keyframe", C.c_int),
                ("res", C.c_double), ("min_keyframe_dist", C.c_double), ("min_keyframe_rot_deg", C.c_double),
                ("downsample_factor", C.c_double), ("cell_capacity", C.c_int), ("sample_capacity", C.c_int)]


class OdomOut(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("n_points", C.c_int), ("n_cells", C.c_int), ("itrs", C.c_int), ("reg_ok", C.c_int),
                ("is_keyframe", C.c_int), ("n_keyframes", C.c_int), ("lm_iterations", C.c_int), ("num_residuals", C.c_int),
                ("status", C.c_int), ("n_samples", C.c_int), ("score", C.c_double)]


def default_reg_params(**kw) -> RegParams:
    p = RegParams(P2L, HUBER, W_UNIFORM, 0.1, 1.0, 0.01, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_odom_params(**kw) -> OdomParams:
    """BASELINE config 2: CFEAR-3 filter (k=40, z_min=60, r=3), 4 keyframes, P2L, Huber 0.1, weight_opt 4, weight_intensity."""
    p = OdomParams(FilterParams(60.0, 40, 2.5, 0.0438), RegParams(P2L, HUBER, W_COMBINED, 0.1, 1.0, 1.0, 0, 0), 4, 1, 1, 1, 0, 1,
                   3.0, 1.5, 5.0, 1.0, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m tbv_slam_public_b200.build` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tbv_create.restype = C.c_void_p
        L.tbv_create.argtypes = [C.c_int]
        L.tbv_destroy.argtypes = [C.c_void_p]
        L.tbv_last_error.restype = C.c_char_p
        L.tbv_stream.restype = C.c_void_p
        L.tbv_stream.argtypes = [C.c_void_p]
        L.tbv_synchronize.argtypes = [C.c_void_p]
        L.tbv_launch_count.restype = C.c_longlong
        L.tbv_launch_count.argtypes = [C.c_void_p]
        L.tbv_profile_begin.argtypes = [C.c_void_p]
        L.tbv_profile_end.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tbv_host_alloc.restype = C.c_void_p
        L.tbv_host_alloc.argtypes = [C.c_size_t]
        L.tbv_host_free.argtypes = [C.c_void_p]
        L.tbv_filter_kstrongest.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(FilterParams),
                                            C.POINTER(Points), C.POINTER(Points)]
        L.tbv_filter_kstrongest_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                                C.POINTER(FilterParams), C.c_int]
        L.tbv_filter_fetch.argtypes = [C.c_void_p, C.POINTER(Points), C.POINTER(Points)]
        L.tbv_compensate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.tbv_rotate90ccw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        for name, argt, rest in _OPTIONAL:
            if hasattr(L, name):
                f = getattr(L, name)
                if argt is not None:
                    f.argtypes = argt
                if rest is not None:
                    f.restype = rest
        _LIB = L
    return _LIB


_OPTIONAL = [
    ("tbv_build_cells", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int, C.c_void_p,
                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
    ("tbv_pair_normal_eq", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(RegParams),
                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], None),
    ("tbv_register", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RegParams), C.POINTER(RegSummary)], None),
    ("tbv_get_cost", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RegParams), C.c_int, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_void_p, C.c_int], None),
    ("tbv_register_batch", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.POINTER(RegParams), C.c_void_p, C.c_void_p, C.c_void_p], None),
    ("tbv_odom_create", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(OdomParams)], C.c_void_p),
    ("tbv_odom_destroy", [C.c_void_p], None),
    ("tbv_odom_reset", [C.c_void_p], None),
    ("tbv_odom_set_wire_layout", [C.c_void_p, C.c_int], None),
    ("tbv_odom_set_graphs", [C.c_void_p, C.c_int], None),
    ("tbv_odom_set_overlap", [C.c_void_p, C.c_int], None),
    ("tbv_odom_step", [C.c_void_p, C.c_void_p, C.c_void_p], None),
    ("tbv_odom_step_dev", [C.c_void_p, C.c_void_p], None),
    ("tbv_odom_fetch", [C.c_void_p, C.c_void_p], None),
    ("tbv_odom_submit", [C.c_void_p, C.c_void_p], None),
    ("tbv_odom_collect", [C.c_void_p, C.c_void_p], None),
    ("tbv_odom_cells", [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
    ("tbv_filter_cacfar", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(CfarParams), C.POINTER(Points)], None),
    ("tbv_loopdb_create", [C.c_void_p, C.c_int, C.c_int], C.c_void_p),
    ("tbv_loopdb_destroy", [C.c_void_p], None),
    ("tbv_loopdb_size", [C.c_void_p], None),
    ("tbv_loopdb_add", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], None),
    ("tbv_loopdb_register", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(RegParams), C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
    ("tbv_loopdb_register_dev", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(RegParams), C.c_double, C.c_void_p, C.c_int, C.c_void_p], None),
    ("tbv_loopdb_register_sharded", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RegParams),
                                     C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
    ("tbv_loopdb_submit_sharded", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RegParams),
                                   C.c_double], None),
    ("tbv_loopdb_collect_sharded", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
    ("tbv_comm_unique_id", [C.c_void_p], None),
    ("tbv_comm_init_rank", [C.c_void_p, C.c_void_p, C.c_int, C.c_int], None),
    ("tbv_comm_init", [C.c_void_p, C.c_void_p], None),
    ("tbv_comm_world", [C.c_void_p, C.c_void_p, C.c_void_p], None),
    ("tbv_comm_destroy", [C.c_void_p], None),
    ("tbv_allgather_constraints", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p], None),
    ("tbv_allgather_constraints_dev", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], None),
]

COMM_ID_BYTES = 128


def _check(rc):
    if rc != 0:
        raise TbvError(rc, lib().tbv_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _PointBufs:
    """numpy storage behind a tbv_points."""

    def __init__(self, batch, cap):
        self.batch, self.cap = batch, cap
        self.count = np.zeros(batch, np.int32)
        self.az = np.zeros((batch, cap), np.uint16)
        self.rg = np.zeros((batch, cap), np.uint16)
        self.I = np.zeros((batch, cap), np.uint8)
        self.x = np.zeros((batch, cap), np.float32)
        self.y = np.zeros((batch, cap), np.float32)
        self.c = Points(cap, self.count.ctypes.data_as(C.POINTER(C.c_int)), self.az.ctypes.data_as(C.POINTER(C.c_uint16)),
                        self.rg.ctypes.data_as(C.POINTER(C.c_uint16)), self.I.ctypes.data_as(C.POINTER(C.c_uint8)),
                        self.x.ctypes.data_as(C.POINTER(C.c_float)), self.y.ctypes.data_as(C.POINTER(C.c_float)))

    def scan(self, b):
        n = int(self.count[b])
        return self.az[b, :n], self.rg[b, :n], self.I[b, :n], self.x[b, :n], self.y[b, :n]


class Context:
    """tbv_ctx: one CUDA device + one stream."""

    def __init__(self, device: int = 0):
        self.h = lib().tbv_create(device)
        if not self.h:
            raise TbvError(TBV_ERR_NO_GPU, lib().tbv_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().tbv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return lib().tbv_stream(self.h)

    def synchronize(self):
        _check(lib().tbv_synchronize(self.h))

    def launch_count(self) -> int:
        return lib().tbv_launch_count(self.h)

    def profile_begin(self):
        _check(lib().tbv_profile_begin(self.h))

    def profile_end(self, capacity=4096):
        """-> list of (kernel name, ms) in launch order since profile_begin."""
        names = (C.c_char_p * capacity)()
        ms = (C.c_float * capacity)()
        n = C.c_int(0)
        _check(lib().tbv_profile_end(self.h, capacity, names, ms, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(min(n.value, capacity))]

    # ---- multi-GPU: one NCCL communicator per context (tbv_comm_*) ----------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId through the library (rank 0 calls it and hands the 128 bytes to every rank)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().tbv_comm_unique_id(buf))
        return buf.raw

    def comm_init_rank(self, unique_id: bytes, world: int, rank: int):
        """Collective: every rank of the group calls it with the same id."""
        assert len(unique_id) == COMM_ID_BYTES
        _check(lib().tbv_comm_init_rank(self.h, unique_id, int(world), int(rank)))

    def comm_world(self):
        w, r = C.c_int(1), C.c_int(0)
        _check(lib().tbv_comm_world(self.h, C.byref(w), C.byref(r)))
        return w.value, r.value

    def comm_destroy(self):
        _check(lib().tbv_comm_destroy(self.h))

    def allgather_constraints(self, local_dev_ptr: int, n_local_dev_ptr: int, capacity: int) -> np.ndarray:
        """tbv_allgather_constraints: records of every rank, global candidate order (CONSTRAINT_DTYPE array, identical on every rank)."""
        world, _ = self.comm_world()
        out = np.zeros(max(world * capacity, 1), CONSTRAINT_DTYPE)
        n = C.c_int(0)
        _check(lib().tbv_allgather_constraints(self.h, C.c_void_p(local_dev_ptr), C.c_void_p(n_local_dev_ptr), int(capacity), _ptr(out), len(out),
                                               C.byref(n)))
        return out[:n.value].copy()

    # ---- radarDriver::Process (k-strongest branch) -----------------------------------------------------------
    def StructuredKStrongest(self, polar: np.ndarray, z_min=60.0, k_strongest=40, min_distance=2.5, range_res=0.0438,
                             peaks=True, n_range=None):
        """polar: [n_az, width] or [batch, n_az, width] u8 (host); n_range <= width bins are used (row stride = width).
        Returns (filtered, peaks) _PointBufs."""
        polar = np.ascontiguousarray(polar, np.uint8)
        if polar.ndim == 2:
            polar = polar[None]
        batch, n_az, width = polar.shape
        n_range = n_range or width
        stride = width
        par = FilterParams(z_min, k_strongest, min_distance, range_res)
        f = _PointBufs(batch, n_az * k_strongest)
        p = _PointBufs(batch, n_az * k_strongest) if peaks else None
        _check(lib().tbv_filter_kstrongest(self.h, _ptr(polar), n_az, n_range, stride, batch, C.byref(par), C.byref(f.c),
                                           C.byref(p.c) if peaks else None))
        return f, p

    def filter_dev(self, polar_dev_ptr: int, n_az, n_range, batch, par: FilterParams, want_peaks=True, row_stride=None):
        _check(lib().tbv_filter_kstrongest_dev(self.h, C.c_void_p(polar_dev_ptr), n_az, n_range, row_stride or n_range, batch,
                                               C.byref(par), int(want_peaks)))

    def filter_fetch(self, batch, n_az, k, peaks=True):
        f = _PointBufs(batch, n_az * k)
        p = _PointBufs(batch, n_az * k) if peaks else None
        _check(lib().tbv_filter_fetch(self.h, C.byref(f.c), C.byref(p.c) if peaks else None))
        return f, p

    def AzimuthCACFAR(self, polar, window_size=40, false_alarm_rate=0.01, nb_guard_cells=10, range_res=0.0438,
                      static_threshold=20.0, min_distance=2.5, max_distance=400.0, capacity=None):
        polar = np.ascontiguousarray(polar, np.uint8)
        if polar.ndim == 2:
            polar = polar[None]
        batch, n_az, n_range = polar.shape
        par = CfarParams(window_size, false_alarm_rate, nb_guard_cells, range_res, static_threshold, min_distance, max_distance)
        out = _PointBufs(batch, capacity or n_az * n_range)
        _check(lib().tbv_filter_cacfar(self.h, _ptr(polar), n_az, n_range, n_range, batch, C.byref(par), C.byref(out.c)))
        return out

    def rotate90ccw(self, src):
        src = np.ascontiguousarray(src, np.uint8)
        H, W = src.shape
        dst = np.zeros((W, H), np.uint8)
        _check(lib().tbv_rotate90ccw(self.h, _ptr(src), H, W, _ptr(dst)))
        return dst

    # ---- CFEAR_Radarodometry::Compensate -------------------------------------------------------------------------
    def Compensate(self, x, y, mot, ccw=False):
        x = np.array(x, np.float32, copy=True)
        y = np.array(y, np.float32, copy=True)
        m = np.ascontiguousarray(mot, np.float64)
        _check(lib().tbv_compensate(self.h, _ptr(x), _ptr(y), len(x), _ptr(m), int(ccw)))
        return x, y

    # ---- MapPointNormal ----------------------------------------------------------------------------------------------
    def MapPointNormal(self, x, y, intensity, radius=3.0, downsample_factor=1.0, weight_intensity=True, origin=(0.0, 0.0),
                       capacity=None):
        """Returns (cells [n,16] float64, n_samples)."""
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        intensity = np.ascontiguousarray(intensity, np.float32)
        cap = capacity or max(len(x), 1)
        cells = np.zeros((cap, 16), np.float64)
        o = np.ascontiguousarray(origin, np.float64)
        nc, ns = C.c_int(0), C.c_int(0)
        _check(lib().tbv_build_cells(self.h, _ptr(x), _ptr(y), _ptr(intensity), len(x), C.c_float(radius), C.c_double(downsample_factor),
                                     int(weight_intensity), _ptr(o), _ptr(cells), cap, C.byref(nc), C.byref(ns)))
        return cells[:nc.value].copy(), ns.value

    # ---- n_scan_normal_reg ---------------------------------------------------------------------------------------------
    @staticmethod
    def _scan_ptrs(scans):
        arrs = [np.ascontiguousarray(s, np.float64).reshape(-1, 16) for s in scans]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        ns = np.array([len(a) for a in arrs], np.int32)
        return arrs, ptrs, ns

    def pair_normal_eq(self, tgt, T_tgt, src, T_src, params: RegParams | None = None, itr=1):
        params = params or default_reg_params()
        tgt = np.ascontiguousarray(tgt, np.float64); src = np.ascontiguousarray(src, np.float64)
        Tt = np.ascontiguousarray(T_tgt, np.float64); Ts = np.ascontiguousarray(T_src, np.float64)
        H, g = np.zeros(9), np.zeros(3)
        cost, nres = C.c_double(0), C.c_int(0)
        assoc = np.zeros(max(len(src), 1), np.int32)
        _check(lib().tbv_pair_normal_eq(self.h, _ptr(tgt), len(tgt), _ptr(Tt), _ptr(src), len(src), _ptr(Ts), C.byref(params), itr,
                                        C.byref(cost), C.byref(nres), _ptr(H), _ptr(g), _ptr(assoc)))
        return dict(n_res=nres.value, cost=cost.value, H=H.reshape(3, 3), g=g, assoc=assoc[:len(src)])

    def Register(self, scans, T, params: RegParams | None = None):
        params = params or default_reg_params()
        arrs, ptrs, ns = self._scan_ptrs(scans)
        Tio = np.array(T, np.float64, copy=True).reshape(len(scans), 3)
        s = RegSummary()
        _check(lib().tbv_register(self.h, len(scans), ptrs, _ptr(ns), _ptr(Tio), C.byref(params), C.byref(s)))
        return Tio, s

    def GetCost(self, scans, T, params: RegParams | None = None, itr=0):
        params = params or default_reg_params()
        arrs, ptrs, ns = self._scan_ptrs(scans)
        Tio = np.ascontiguousarray(T, np.float64).reshape(len(scans), 3)
        cap = int(2 * ns[-1] * (len(scans) - 1)) + 8
        res = np.zeros(cap)
        score, cost, n = C.c_double(0), C.c_double(0), C.c_int(0)
        _check(lib().tbv_get_cost(self.h, len(scans), ptrs, _ptr(ns), _ptr(Tio), C.byref(params), itr, C.byref(score), C.byref(cost),
                                  C.byref(n), _ptr(res), cap))
        return n.value, score.value, cost.value, res[:max(n.value, 0)]

    def approximateCovarianceBySampling(self, scans, T, score_scale, params: RegParams | None = None, itr=2, xy_range=0.4,
                                        yaw_range=0.0043625, samples_per_axis=3, covariance_scaler=4.0):
        """OdometryKeyframeFuser::approximateCovarianceBySampling: returns (ok, cov6x6, samples[n^3, 4])."""
        params = params or default_reg_params()
        arrs, ptrs, ns = self._scan_ptrs(scans)
        Tio = np.ascontiguousarray(T, np.float64).reshape(len(scans), 3)
        samples = np.zeros((samples_per_axis ** 3, 4))
        _check(lib().tbv_cost_samples(self.h, len(scans), ptrs, _ptr(ns), _ptr(Tio), C.byref(params), itr, C.c_double(xy_range),
                                      C.c_double(yaw_range), samples_per_axis, _ptr(samples)))
        cov = np.zeros((6, 6))
        ok = C.c_int(0)
        _check(lib().tbv_cov_from_cost_samples(_ptr(samples), len(samples), C.c_double(score_scale), C.c_double(covariance_scaler), _ptr(cov),
                                               C.byref(ok)))
        return bool(ok.value), cov, samples

    def CFEARQualityBatch(self, sets, src_set, ref_set, T_src, T_ref, T_offset=None, params: RegParams | None = None):
        """CFEARQuality for many candidate pairs at once: [n, 3] = score, residual count, mean set size."""
        params = params or default_reg_params(cost=P2L, loss=HUBER, loss_limit=0.3, weight_opt=W_UNIFORM)
        arrs, ptrs, ns = self._scan_ptrs(sets)
        ss = np.ascontiguousarray(src_set, np.int32); rs = np.ascontiguousarray(ref_set, np.int32)
        n = len(ss)
        Ts = np.ascontiguousarray(T_src, np.float64).reshape(n, 3); Tr = np.ascontiguousarray(T_ref, np.float64).reshape(n, 3)
        To = np.ascontiguousarray(T_offset, np.float64).reshape(n, 3) if T_offset is not None else None
        q = np.zeros((n, 3))
        _check(lib().tbv_cfear_quality_batch(self.h, len(sets), ptrs, _ptr(ns), n, _ptr(ss), _ptr(rs), _ptr(Ts), _ptr(To), _ptr(Tr), C.byref(params),
                                             _ptr(q)))
        return q

    def RegisterBatch(self, sets, from_set, to_set, T_from, T_to, params: RegParams | None = None):
        """loopclosure::Register for many candidates at once. Returns (T_revised [n,3], T_align [n,3], summaries)."""
        params = params or default_reg_params(max_itr_association=4, max_itr_solver=10)
        arrs, ptrs, ns = self._scan_ptrs(sets)
        fs = np.ascontiguousarray(from_set, np.int32); ts = np.ascontiguousarray(to_set, np.int32)
        Tf = np.ascontiguousarray(T_from, np.float64).reshape(-1, 3); Tt = np.ascontiguousarray(T_to, np.float64).reshape(-1, 3)
        n = len(fs)
        Tr, Ta = np.zeros((n, 3)), np.zeros((n, 3))
        summ = (RegSummary * n)()
        _check(lib().tbv_register_batch(self.h, len(arrs), ptrs, _ptr(ns), n, _ptr(fs), _ptr(ts), _ptr(Tf), _ptr(Tt), C.byref(params),
                                        _ptr(Tr), _ptr(Ta), C.cast(summ, C.c_void_p)))
        return Tr, Ta, summ


# tbv_constraint (include/tbv_b200.h): Constraint3d for a planar pose, 128 bytes — the record all-gathered between GPUs
CONSTRAINT_DTYPE = np.dtype([("id_begin", np.int32), ("id_end", np.int32), ("type", np.int32), ("candidate", np.int32),
                             ("t_be", np.float64, 3), ("cov", np.float64, 4), ("score", np.float64), ("t_revised", np.float64, 3),
                             ("itrs", np.int32), ("num_residuals", np.int32), ("quality", np.float64, 2)])
assert CONSTRAINT_DTYPE.itemsize == 128


def loop_reg_params(**kw) -> RegParams:
    """loopclosure::Register: n_scan_normal_reg(P2L) with ctor defaults (Huber 0.1, uniform weights) + SetParameters(4, 10)."""
    return default_reg_params(max_itr_association=4, max_itr_solver=10, **kw)


class LoopDB:
    """Keyframe cell sets resident on the GPU + batched loopclosure::RegisterLoopCandidate (tbv_loopdb_*)."""

    def __init__(self, ctx: Context, max_keyframes: int, cell_capacity: int = 1024):
        self.ctx, self.cell_capacity = ctx, cell_capacity
        self.h = lib().tbv_loopdb_create(ctx.h, max_keyframes, cell_capacity)
        if not self.h:
            raise TbvError(TBV_ERR_CUDA, lib().tbv_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().tbv_loopdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().tbv_loopdb_size(self.h)

    def add(self, sets) -> int:
        """Appends keyframes (each an [n,16] array of tbv_cell records); returns the id of the first one."""
        arrs, ptrs, ns = Context._scan_ptrs(sets)
        first = C.c_int(0)
        _check(lib().tbv_loopdb_add(self.h, len(arrs), ptrs, _ptr(ns), C.byref(first)))
        return first.value

    @staticmethod
    def _cand_args(id_from, id_to, T_from, T_to, candidate_index, quality):
        fs = np.ascontiguousarray(id_from, np.int32); ts = np.ascontiguousarray(id_to, np.int32)
        Tf = np.ascontiguousarray(T_from, np.float64).reshape(-1, 3); Tt = np.ascontiguousarray(T_to, np.float64).reshape(-1, 3)
        ci = None if candidate_index is None else np.ascontiguousarray(candidate_index, np.int32)
        q = None if quality is None else np.ascontiguousarray(quality, np.float64).reshape(-1, 2)
        return fs, ts, Tf, Tt, ci, q

    def register_candidates(self, id_from, id_to, T_from, T_to, candidate_index=None, quality=None, params: RegParams | None = None,
                            max_score=0.0, want_summaries=False):
        """Returns the accepted constraints (CONSTRAINT_DTYPE array, candidate order) [and every candidate's RegSummary]."""
        params = params or loop_reg_params()
        fs, ts, Tf, Tt, ci, q = self._cand_args(id_from, id_to, T_from, T_to, candidate_index, quality)
        n = len(fs)
        out = np.zeros(max(n, 1), CONSTRAINT_DTYPE)
        n_out = C.c_int(0)
        summ = (RegSummary * max(n, 1))() if want_summaries else None
        _check(lib().tbv_loopdb_register(self.h, n, _ptr(fs), _ptr(ts), _ptr(Tf), _ptr(Tt), _ptr(ci), _ptr(q), C.byref(params),
                                         float(max_score), _ptr(out), n, C.byref(n_out), C.cast(summ, C.c_void_p) if summ else None))
        out = out[:n_out.value].copy()
        return (out, summ) if want_summaries else out

    def register_sharded(self, id_from, id_to, T_from, T_to, quality=None, params: RegParams | None = None, max_score=0.0, want_timing=False):
        """tbv_loopdb_register_sharded: every rank passes the SAME global candidate list; each registers the candidates with
        id_from mod world == rank and all ranks receive every accepted constraint in global candidate order (one NCCL all-gather
        inside the library).  Without a communicator on the context (world 1) it equals register_candidates."""
        params = params or loop_reg_params()
        fs, ts, Tf, Tt, _, q = self._cand_args(id_from, id_to, T_from, T_to, None, quality)
        n = len(fs)
        out = np.empty(max(n, 1), CONSTRAINT_DTYPE)          # the library writes records [0, n_out); the view below is all the caller sees
        n_out = C.c_int(0)
        tm = (C.c_float * 4)() if want_timing else None
        _check(lib().tbv_loopdb_register_sharded(self.h, n, _ptr(fs), _ptr(ts), _ptr(Tf), _ptr(Tt), _ptr(q), C.byref(params), float(max_score),
                                                 _ptr(out), n, C.byref(n_out), tm))
        out = out[:n_out.value]
        return (out, [float(v) for v in tm]) if want_timing else out

    def submit_sharded(self, id_from, id_to, T_from, T_to, quality=None, params: RegParams | None = None, max_score=0.0):
        """tbv_loopdb_submit_sharded: enqueue one sharded batch and return at once (at most two in flight; every rank submits the same
        batches in the same order).  The exchange of a batch overlaps the registration of the next one."""
        params = params or loop_reg_params()
        fs, ts, Tf, Tt, _, q = self._cand_args(id_from, id_to, T_from, T_to, None, quality)
        _check(lib().tbv_loopdb_submit_sharded(self.h, len(fs), _ptr(fs), _ptr(ts), _ptr(Tf), _ptr(Tt), _ptr(q), C.byref(params), float(max_score)))
        self._pending = getattr(self, "_pending", []) + [len(fs)]

    def collect_sharded(self, want_timing=False):
        """tbv_loopdb_collect_sharded: the accepted constraints of the OLDEST batch in flight, global candidate order (blocks until they
        are on the host)."""
        pending = getattr(self, "_pending", [])
        n = pending[0] if pending else 0
        self._pending = pending[1:]
        out = np.empty(max(n, 1), CONSTRAINT_DTYPE)
        n_out = C.c_int(0)
        tm = (C.c_float * 4)() if want_timing else None
        _check(lib().tbv_loopdb_collect_sharded(self.h, _ptr(out), n, C.byref(n_out), tm))
        out = out[:n_out.value]
        return (out, [float(v) for v in tm]) if want_timing else out

    def register_candidates_dev(self, id_from, id_to, T_from, T_to, out_dev_ptr: int, out_capacity: int, n_out_dev_ptr: int,
                                candidate_index=None, quality=None, params: RegParams | None = None, max_score=0.0):
        """Same, leaving the records on the device (for a collective); enqueued on the context's stream."""
        params = params or loop_reg_params()
        fs, ts, Tf, Tt, ci, q = self._cand_args(id_from, id_to, T_from, T_to, candidate_index, quality)
        _check(lib().tbv_loopdb_register_dev(self.h, len(fs), _ptr(fs), _ptr(ts), _ptr(Tf), _ptr(Tt), _ptr(ci), _ptr(q), C.byref(params),
                                             float(max_score), C.c_void_p(out_dev_ptr), out_capacity, C.c_void_p(n_out_dev_ptr)))


class OdometryKeyframeFuser:
    """n_seq independent radarDriver + OdometryKeyframeFuser pipelines advanced in lock-step on one GPU."""

    def __init__(self, ctx: Context, n_seq: int, n_az: int, n_range: int, params: OdomParams | None = None):
        self.ctx, self.n_seq, self.n_az, self.n_range = ctx, n_seq, n_az, n_range
        self.params = params or default_odom_params()
        self.h = lib().tbv_odom_create(ctx.h, n_seq, n_az, n_range, C.byref(self.params))
        if not self.h:
            raise TbvError(TBV_ERR_INVALID, lib().tbv_last_error().decode())
        self._out = (OdomOut * n_seq)()

    def close(self):
        if getattr(self, "h", None):
            lib().tbv_odom_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_graphs(self, enable: bool):
        """CUDA-graph replay of the step for input buffers seen before (default on)."""
        _check(lib().tbv_odom_set_graphs(self.h, int(bool(enable))))

    def set_overlap(self, enable: bool):
        """Overlapped steps for large batches: the filter of a step runs on a second stream under the registration of the previous step
        (identical results; scans given to step_dev must be complete in device memory when the call is made)."""
        _check(lib().tbv_odom_set_overlap(self.h, int(bool(enable))))

    def set_wire_layout(self, range_major: bool):
        """True: scans arrive [n_range][n_az] (MulRan wire layout) and are rotated 90 deg CCW on the device on receipt (radar_driver.cpp:80-84)."""
        _check(lib().tbv_odom_set_wire_layout(self.h, int(bool(range_major))))

    def reset(self):
        _check(lib().tbv_odom_reset(self.h))

    def pointcloudCallback(self, polar_host: np.ndarray):
        """One scan per sequence [n_seq, n_az, n_range] u8 from host memory -> list of OdomOut."""
        polar_host = np.ascontiguousarray(polar_host, np.uint8)
        assert polar_host.size == self.n_seq * self.n_az * self.n_range
        _check(lib().tbv_odom_step(self.h, _ptr(polar_host), C.cast(self._out, C.c_void_p)))
        return self._out

    def step_dev(self, polar_dev_ptr: int):
        _check(lib().tbv_odom_step_dev(self.h, C.c_void_p(polar_dev_ptr)))

    def fetch(self):
        _check(lib().tbv_odom_fetch(self.h, C.cast(self._out, C.c_void_p)))
        return self._out

    def submit(self, pinned_ptr: int):
        _check(lib().tbv_odom_submit(self.h, C.c_void_p(pinned_ptr)))

    def collect(self):
        _check(lib().tbv_odom_collect(self.h, C.cast(self._out, C.c_void_p)))
        return self._out

    def cells(self, seq: int, keyframe: int = -1, capacity: int = 8192):
        cells = np.zeros((capacity, 16))
        n = C.c_int(0)
        pose = np.zeros(3)
        _check(lib().tbv_odom_cells(self.h, seq, keyframe, _ptr(cells), capacity, C.byref(n), _ptr(pose)))
        return cells[:n.value].copy(), pose


def poses(outs) -> np.ndarray:
    return np.array([[o.pose[0], o.pose[1], o.pose[2]] for o in outs])


class PinnedBuffer:
    """cudaHostAlloc'ed byte buffer (tbv_host_alloc) exposed as a numpy u8 array: the source of pipelined uploads."""

    def __init__(self, nbytes: int):
        self.ptr = lib().tbv_host_alloc(nbytes)
        if not self.ptr:
            raise TbvError(TBV_ERR_CUDA, lib().tbv_last_error().decode())
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().tbv_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- Scan Context / pose graph ------------------------------------------------------------------------------------------
class SCParams(C.Structure):
    _fields_ = [("num_ring", C.c_int), ("num_sector", C.c_int), ("max_radius", C.c_double), ("search_ratio", C.c_double),
                ("num_candidates_from_tree", C.c_int), ("n_candidates", C.c_int), ("odom_sigma_error", C.c_double),
                ("odometry_coupled_closure", C.c_int), ("augment_sc", C.c_int), ("no_point", C.c_double),
                ("desc_function", C.c_int), ("desc_divider", C.c_double), ("distance_exclude_recent", C.c_double)]


class PGOParams(C.Structure):
    _fields_ = [("odom_vxx", C.c_double), ("odom_vyy", C.c_double), ("odom_vtt", C.c_double), ("loop_scaling", C.c_double),
                ("replace_cov_by_identity", C.c_int), ("loop_cauchy", C.c_double)]


class PGOOptions(C.Structure):   # tbv_pgo_options: fields <= 0 -> the reference's values
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_radius", C.c_double), ("max_cg_iterations", C.c_int), ("cg_rel_tol", C.c_double)]


class PGOSummaryC(C.Structure):  # tbv_pgo_summary
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("iterations", C.c_int), ("successful_steps", C.c_int),
                ("cg_iterations", C.c_int), ("termination", C.c_int), ("device_ms", C.c_float)]


PGO_TERMINATION = ("max_num_iterations", "gradient_tolerance", "parameter_tolerance", "function_tolerance", "min_trust_region_radius",
                   "failure: consecutive invalid steps")


def default_sc_params(**kw) -> SCParams:
    """TBV-8 offline settings (tbv_slam/src/tbv_slam_offline.cpp:81-101): 40 x 120, 80 m, sum / 1000, 3 candidates, augmentations."""
    p = SCParams(40, 120, 80.0, 0.1, 10, 3, 0.05, 1, 1, 0.0, 0, 1000.0, 10.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_pgo_params(**kw) -> PGOParams:
    p = PGOParams(0.01, 0.01, 0.001, 500000.0, 1, 0.1)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


AUGMENTS = ((0.0, 0.0), (0.0, -2.0), (0.0, 2.0), (0.0, -4.0), (0.0, 4.0))   # RadarScancontext.cpp:162-166 (+ identity)


def _sc_bind():
    L = lib()
    if getattr(L, "_sc_bound", False):
        return L
    L.tbv_sc_make.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SCParams), C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]
    L.tbv_sc_distance_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(SCParams), C.c_void_p, C.c_void_p]
    L.tbv_sc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(SCParams), C.c_void_p,
                                C.c_void_p, C.c_void_p]
    L.tbv_pgo_assemble.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PGOParams), C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tbv_pgo_solve_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                     C.c_double, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.tbv_pgo_solve_damped.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                       C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.tbv_pgo_optimize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PGOParams), C.c_int,
                                   C.POINTER(PGOOptions), C.POINTER(PGOSummaryC)]
    L._sc_bound = True
    return L


def sc_make(ctx: Context, x, y, intensity, params: SCParams | None = None, offsets=((0.0, 0.0),)):
    """RSCManager::MakeRadarCloudContext for each lateral offset. Returns (desc [n_off, S*R], ringkey [n_off, R] f32, sectorkey [n_off, S])."""
    params = params or default_sc_params()
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32); intensity = np.ascontiguousarray(intensity, np.float32)
    off = np.ascontiguousarray(offsets, np.float64).reshape(-1, 2)
    R, S = params.num_ring, params.num_sector
    desc = np.zeros((len(off), R * S)); rk = np.zeros((len(off), R), np.float32); sk = np.zeros((len(off), S))
    _check(_sc_bind().tbv_sc_make(ctx.h, _ptr(x), _ptr(y), _ptr(intensity), len(x), C.byref(params), len(off), _ptr(off), _ptr(desc), _ptr(rk), _ptr(sk)))
    return desc, rk, sk


def sc_distance_batch(ctx: Context, desc_q, desc_c, q_idx, c_idx, params: SCParams | None = None):
    params = params or default_sc_params()
    dq = np.ascontiguousarray(desc_q, np.float64).reshape(-1, params.num_ring * params.num_sector)
    dc = np.ascontiguousarray(desc_c, np.float64).reshape(-1, params.num_ring * params.num_sector)
    qi = np.ascontiguousarray(q_idx, np.int32); ci = np.ascontiguousarray(c_idx, np.int32)
    dist = np.zeros(len(qi)); shift = np.zeros(len(qi), np.int32)
    _check(_sc_bind().tbv_sc_distance_batch(ctx.h, _ptr(dq), len(dq), _ptr(dc), len(dc), len(qi), _ptr(qi), _ptr(ci), C.byref(params), _ptr(dist),
                                            _ptr(shift)))
    return dist, shift


def sc_search(ctx: Context, db_keys, odom_xyt, q_keys, q_current, params: SCParams | None = None):
    """Returns (cand_idx [n_q, k] (-1 padded), cand_odom_sim [n_q, k], n_exclude [n_q])."""
    params = params or default_sc_params()
    dk = np.ascontiguousarray(db_keys, np.float32).reshape(-1, params.num_ring)
    od = np.ascontiguousarray(odom_xyt, np.float64).reshape(-1, 3)
    qk = np.ascontiguousarray(q_keys, np.float32).reshape(-1, params.num_ring)
    qc = np.ascontiguousarray(q_current, np.int32)
    k = params.num_candidates_from_tree
    ci = np.zeros((len(qc), k), np.int32); cs = np.zeros((len(qc), k)); ne = np.zeros(len(qc), np.int32)
    _check(_sc_bind().tbv_sc_search(ctx.h, _ptr(dk), _ptr(od), len(dk), len(qc), _ptr(qk), _ptr(qc), C.byref(params), _ptr(ci), _ptr(cs), _ptr(ne)))
    return ci, cs, ne


def pgo_assemble(ctx: Context, nodes, ids, meas, params: PGOParams | None = None, info=None, fixed_node=0):
    """CeresLeastSquares problem build + one evaluation. Returns (cost, H_diag [n,6,6], H_off [m,6,6], g [n,6], residuals [m,6])."""
    params = params or default_pgo_params()
    nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, 7)
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    meas = np.ascontiguousarray(meas, np.float64).reshape(-1, 7)
    n, m = len(nodes), len(ids)
    Hd, Ho, g, res = np.zeros((n, 36)), np.zeros((max(m, 1), 36)), np.zeros((n, 6)), np.zeros((max(m, 1), 6))
    inf = np.ascontiguousarray(info, np.float64) if info is not None else None
    cost = C.c_double(0)
    _check(_sc_bind().tbv_pgo_assemble(ctx.h, n, _ptr(nodes), m, _ptr(ids), _ptr(meas), _ptr(inf), C.byref(params), fixed_node, C.byref(cost),
                                       _ptr(Hd), _ptr(Ho), _ptr(g), _ptr(res)))
    return cost.value, Hd.reshape(n, 6, 6), Ho[:m].reshape(m, 6, 6), g, res[:m]


def pgo_solve_step(ctx: Context, ids, H_diag, H_off, g, fixed_node=0, radius=1e4, max_iters=20000, rel_tol=1e-12):
    """(H + D) delta = -g, D = clamp(diag H, 1e-6, 1e32) / radius: the linear solve of one Ceres LM iteration (ceresoptimizer.cpp:50-62),
    block-Jacobi PCG on the device. Returns (delta [n, 6], cg_iterations, relative_residual)."""
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    Hd = np.ascontiguousarray(H_diag, np.float64).reshape(-1, 36)
    m = len(ids)
    Ho = np.ascontiguousarray(H_off, np.float64).reshape(-1, 36)[:m]
    if m == 0:
        Ho = np.zeros((1, 36))
    g = np.ascontiguousarray(g, np.float64).reshape(-1, 6)
    n = len(Hd)
    if len(g) != n or len(Ho) < m:
        raise ValueError("H_diag, H_off, g do not describe one graph")
    delta = np.zeros((n, 6)); it = C.c_int(0); rel = C.c_double(0)
    _check(_sc_bind().tbv_pgo_solve_step(ctx.h, n, m, _ptr(ids), _ptr(Hd), _ptr(Ho), _ptr(g), fixed_node, float(radius), int(max_iters), float(rel_tol),
                                         _ptr(delta), C.byref(it), C.byref(rel)))
    return delta, it.value, rel.value


def pgo_plus(nodes, delta):
    """x (+) delta of the graph's parameter blocks: p += dp; q = exp(dr) * q (ceres::EigenQuaternionParameterization::Plus, q = x y z w)."""
    nodes = np.array(nodes, np.float64).reshape(-1, 7)
    delta = np.asarray(delta, np.float64).reshape(-1, 6)
    out = nodes.copy()
    out[:, :3] += delta[:, :3]
    nrm = np.linalg.norm(delta[:, 3:], axis=1)
    k = np.where(nrm > 0, np.sin(nrm) / np.where(nrm > 0, nrm, 1.0), 1.0)
    dv, dw = delta[:, 3:] * k[:, None], np.cos(nrm)
    v, w = nodes[:, 3:6], nodes[:, 6]
    out[:, 6] = dw * w - np.einsum("ij,ij->i", dv, v)
    out[:, 3:6] = dw[:, None] * v + w[:, None] * dv + np.cross(dv, v)
    return out


def _pgo_hessian_times(ids, Hd, Ho, x):
    """H x for the block layout of pgo_assemble (host side: the model-cost term of the step-quality ratio)."""
    y = np.einsum("nij,nj->ni", Hd, x)
    if len(ids):
        a, b = ids[:, 0], ids[:, 1]
        np.add.at(y, a, np.einsum("cij,cj->ci", Ho, x[b]))
        np.add.at(y, b, np.einsum("cji,cj->ci", Ho, x[a]))
    return y


class PGOSummary:
    def __init__(self):
        self.initial_cost = self.final_cost = 0.0
        self.iterations, self.successful_steps, self.cg_iterations = 0, 0, 0
        self.termination = ""

    def __repr__(self):
        return (f"PGOSummary(initial_cost={self.initial_cost:.6e}, final_cost={self.final_cost:.6e}, iterations={self.iterations}, "
                f"successful_steps={self.successful_steps}, cg_iterations={self.cg_iterations}, termination={self.termination!r})")


def pgo_optimize(ctx: Context, nodes, ids, meas, params: PGOParams | None = None, info=None, fixed_node=0, max_num_iterations=200,
                 function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8, initial_radius=1e4, cg_rel_tol=1e-10,
                 cg_max_iters=20000):
    """CeresLeastSquares::Solve (ceresoptimizer.cpp:13-62): Levenberg-Marquardt over the pose graph, max_num_iterations = 200, every other
    option the Ceres 2.1.0 default (trust_region_minimizer.cc: initial radius 1e4, accept rho > 1e-3, radius /= max(1/3, 1 - (2 rho - 1)^3) on
    success, radius /= 2, 4, 8 ... on consecutive failures; stop on |dcost| <= function_tolerance * cost, max |g| <= gradient_tolerance,
    |step| <= parameter_tolerance * (|x| + parameter_tolerance)).  Evaluation (tbv_pgo_assemble) and the linear solve (tbv_pgo_solve_step)
    run on the device; this loop is the trust-region bookkeeping only.  Ceres' Jacobi column scaling is not applied, so iterates differ from
    Ceres' while the fixed point is the same: parity is stated on the optimum, not on the trajectory.  Returns (nodes [n, 7], PGOSummary)."""
    params = params or default_pgo_params()
    x = np.array(nodes, np.float64).reshape(-1, 7)
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    S = PGOSummary()
    cost, Hd, Ho, g, _ = pgo_assemble(ctx, x, ids, meas, params, info, fixed_node)
    S.initial_cost = S.final_cost = cost
    radius, decrease = float(initial_radius), 2.0
    S.termination = "max_num_iterations"
    if np.abs(g).max(initial=0.0) <= gradient_tolerance:
        S.termination = "gradient_tolerance"
        return x, S
    for _it in range(max_num_iterations):
        S.iterations += 1
        delta, cg_it, _rel = pgo_solve_step(ctx, ids, Hd, Ho, g, fixed_node, radius, cg_max_iters, cg_rel_tol)
        S.cg_iterations += cg_it
        model_change = -float(np.sum(delta * (g + 0.5 * _pgo_hessian_times(ids, Hd, Ho, delta))))
        step_norm, x_norm = float(np.linalg.norm(delta)), float(np.linalg.norm(x))
        if not np.all(np.isfinite(delta)) or model_change <= 0.0:
            radius /= decrease; decrease *= 2.0
            if radius < 1e-32:
                S.termination = "min_trust_region_radius"
                break
            continue
        if step_norm <= parameter_tolerance * (x_norm + parameter_tolerance):
            S.termination = "parameter_tolerance"
            break
        x_new = pgo_plus(x, delta)
        cost_new, Hd_n, Ho_n, g_n, _ = pgo_assemble(ctx, x_new, ids, meas, params, info, fixed_node)
        rho = (cost - cost_new) / model_change
        if rho > 1e-3:
            S.successful_steps += 1
            dcost = cost - cost_new
            x, cost, Hd, Ho, g = x_new, cost_new, Hd_n, Ho_n, g_n
            S.final_cost = cost
            radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16)
            decrease = 2.0
            if np.abs(g).max(initial=0.0) <= gradient_tolerance:
                S.termination = "gradient_tolerance"
                break
            if abs(dcost) <= function_tolerance * cost:
                S.termination = "function_tolerance"
                break
        else:
            radius /= decrease; decrease *= 2.0
            if radius < 1e-32:
                S.termination = "min_trust_region_radius"
                break
    return x, S


def pgo_solve_damped(ctx: Context, ids, H_diag, H_off, g, damping, fixed_node=0, max_iters=20000, rel_tol=1e-12):
    """(H + diag(damping)) delta = -g with a caller-given damping vector [n, 6] (> 0), on the device (tbv_pgo_solve_damped)."""
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    Hd = np.ascontiguousarray(H_diag, np.float64).reshape(-1, 36)
    m = len(ids)
    Ho = np.ascontiguousarray(H_off, np.float64).reshape(-1, 36)[:m]
    if m == 0:
        Ho = np.zeros((1, 36))
    g = np.ascontiguousarray(g, np.float64).reshape(-1, 6)
    n = len(Hd)
    damping = np.ascontiguousarray(damping, np.float64).reshape(-1, 6)
    if damping.shape != (n, 6) or not np.all(damping > 0):
        raise ValueError("damping: [n, 6], positive")
    delta = np.zeros((n, 6)); it = C.c_int(0); rel = C.c_double(0)
    _check(_sc_bind().tbv_pgo_solve_damped(ctx.h, n, m, _ptr(ids), _ptr(Hd), _ptr(Ho), _ptr(g), _ptr(damping), fixed_node, 1.0, int(max_iters),
                                           float(rel_tol), _ptr(delta), C.byref(it), C.byref(rel)))
    return delta, it.value, rel.value


def pgo_optimize_device(ctx: Context, nodes, ids, meas, params: PGOParams | None = None, info=None, fixed_node=0, max_num_iterations=200,
                        function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8, initial_radius=1e4, cg_rel_tol=1e-10,
                        cg_max_iters=20000):
    """CeresLeastSquares::Solve with the whole Levenberg-Marquardt loop on the device (tbv_pgo_optimize): the same iteration rules as
    pgo_optimize_ceres below, which drives them from the host one device call at a time and is this call's checker.
    Returns (nodes [n, 7], PGOSummary) — the summary also carries device_ms."""
    params = params or default_pgo_params()
    x = np.array(nodes, np.float64).reshape(-1, 7).copy()
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    meas = np.ascontiguousarray(meas, np.float64).reshape(-1, 7)
    info = None if info is None else np.ascontiguousarray(info, np.float64).reshape(-1, 36)
    opt = PGOOptions(int(max_num_iterations), float(function_tolerance), float(gradient_tolerance), float(parameter_tolerance), float(initial_radius),
                     int(cg_max_iters), float(cg_rel_tol))
    sc = PGOSummaryC()
    _check(_sc_bind().tbv_pgo_optimize(ctx.h, len(x), _ptr(x), len(ids), _ptr(ids), _ptr(meas), _ptr(info), C.byref(params), fixed_node, C.byref(opt),
                                       C.byref(sc)))
    S = PGOSummary()
    S.initial_cost, S.final_cost, S.iterations, S.successful_steps, S.cg_iterations = sc.initial_cost, sc.final_cost, sc.iterations, sc.successful_steps, sc.cg_iterations
    S.termination = PGO_TERMINATION[sc.termination]
    S.device_ms = float(sc.device_ms)
    return x, S


def pgo_optimize_ceres(ctx: Context, nodes, ids, meas, params: PGOParams | None = None, info=None, fixed_node=0, max_num_iterations=200,
                       function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8, initial_radius=1e4, cg_rel_tol=1e-10,
                       cg_max_iters=20000):
    """ceres::Solve as CeresLeastSquares::SolveOptimizationProblem configures it (ceresoptimizer.cpp:50-62), restated step by step the way
    oracle/tbv_oracle_reg.hpp restates it for the 3-parameter registration problem (TrustRegionMinimizer + LevenbergMarquardtStrategy of
    Ceres 2.1.0, monotonic steps):
      * Jacobi scaling s = 1 / (1 + sqrt(diag H)) fixed at iteration 0; LM diagonal clamp(diag(S H S), 1e-6, 1e32), recomputed after a
        successful step and reused after a rejected one; step from (S H S + diag / radius) y = -S g, delta = S y — on the device as
        (H + diag / (radius s^2)) delta = -g (pgo_solve_damped);
      * an invalid step (model change <= 0) shrinks the radius, five in a row fail the solve;
      * parameter and function tolerance are tested on the CANDIDATE before it is accepted (a converged run does not take its last step);
      * accepted when (cost - candidate) / model change > 1e-3; radius /= max(1/3, 1 - (2 rho - 1)^3), capped at 1e16; rejected: radius
        /= 2, 4, 8 ...; gradient tolerance on max |x - Plus(x, -g)| after every successful step; radius < 1e-32 ends the run.
    Ceres itself is not in the container, so the restatement is unpinned; pgo_optimize stays the default driver until this
    one has run on the GPU.  Returns (nodes [n, 7], PGOSummary)."""
    params = params or default_pgo_params()
    x = np.array(nodes, np.float64).reshape(-1, 7)
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    n = len(x)
    S = PGOSummary()
    free = np.ones((n, 1)); free[fixed_node] = 0.0
    idx = np.arange(6)

    def evaluate(xx):
        cost, Hd, Ho, g, _ = pgo_assemble(ctx, xx, ids, meas, params, info, fixed_node)
        return cost, Hd, Ho, g

    def gradient_max_norm(xx, g):
        return float(np.abs(xx - pgo_plus(xx, -g * free)).max(initial=0.0))

    x_cost, Hd, Ho, g = evaluate(x)
    S.initial_cost = S.final_cost = x_cost
    diag_h = Hd[:, idx, idx]
    scale = 1.0 / (1.0 + np.sqrt(np.maximum(diag_h, 0.0)))          # jacobi scaling, iteration 0 only
    gmax = gradient_max_norm(x, g)
    radius, decrease, lm_diag, reuse, invalid = float(initial_radius), 2.0, None, False, 0
    x_norm = float(np.linalg.norm(x))
    step_ok = True
    S.termination = "max_num_iterations"
    iteration = 0
    while True:
        # FinalizeIterationAndCheckIfMinimizerCanContinue
        if iteration >= max_num_iterations:
            S.termination = "max_num_iterations"; break
        if step_ok and gmax <= gradient_tolerance:
            S.termination = "gradient_tolerance"; break
        if radius < 1e-32:
            S.termination = "min_trust_region_radius"; break
        iteration += 1
        S.iterations = iteration
        # LevenbergMarquardtStrategy::ComputeStep
        if not reuse:
            lm_diag = np.clip(Hd[:, idx, idx] * scale * scale, 1e-6, 1e32)
        damping = lm_diag / (radius * scale * scale)
        damping[fixed_node] = 1.0                                    # the fixed block is not a variable; any positive value
        delta, cg_it, _ = pgo_solve_damped(ctx, ids, Hd, Ho, g, damping, fixed_node, cg_max_iters, cg_rel_tol)
        S.cg_iterations += cg_it
        reuse = True
        model_change = -float(np.sum(delta * (g + 0.5 * _pgo_hessian_times(ids, Hd, Ho, delta))))
        if not (np.all(np.isfinite(delta)) and model_change > 0.0):  # HandleInvalidStep
            invalid += 1
            step_ok = False
            if invalid >= 5:
                S.termination = "failure: consecutive invalid steps"; break
            radius /= decrease; decrease *= 2.0
            continue
        invalid = 0
        candidate = pgo_plus(x, delta)
        cand_cost, Hd_c, Ho_c, g_c = evaluate(candidate)
        if not math.isfinite(cand_cost):
            cand_cost = float(np.finfo(np.float64).max)
        if float(np.linalg.norm(x - candidate)) <= parameter_tolerance * (x_norm + parameter_tolerance):
            S.termination = "parameter_tolerance"; break
        if abs(x_cost - cand_cost) <= function_tolerance * x_cost:
            S.termination = "function_tolerance"; break
        rho = (x_cost - cand_cost) / model_change if cand_cost < np.finfo(np.float64).max else -np.inf
        if rho > 1e-3:                                               # HandleSuccessfulStep
            x, x_cost, Hd, Ho, g = candidate, cand_cost, Hd_c, Ho_c, g_c
            x_norm = float(np.linalg.norm(x))
            gmax = gradient_max_norm(x, g)
            step_ok = True
            S.successful_steps += 1
            S.final_cost = min(S.final_cost, x_cost)
            radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16)
            decrease, reuse = 2.0, False
        else:
            step_ok = False
            radius /= decrease; decrease *= 2.0
    return x, S


class RSCManager:
    """Host mirror of RSCManager (RadarScancontext.h:31-131): the database lives here, every computation runs on the GPU.

    makeAndSaveScancontextAndKeysRadarCloud -> sc_make (identity + 4 lateral augmentations in one launch);
    detectLoopClosureID -> sc_search (ring-key NN with the odometry likelihood) + sc_distance_batch over the <= 50 pairs,
    then the reference's own candidate bookkeeping (RadarScancontext.cpp:286-345: sort by distance, keep N_CANDIDATES).
    """

    def __init__(self, ctx: Context, params: SCParams | None = None):
        self.ctx, self.par = ctx, params or default_sc_params()
        self.polarcontexts, self.ringkeys, self.odom = [], [], []
        self.queries = None

    def makeAndSaveScancontextAndKeysRadarCloud(self, x, y, intensity, Todom):
        offs = AUGMENTS if self.par.augment_sc else AUGMENTS[:1]
        desc, rk, _ = sc_make(self.ctx, x, y, intensity, self.par, offs)
        self.polarcontexts.append(desc[0].copy()); self.ringkeys.append(rk[0].copy()); self.odom.append(np.asarray(Todom, np.float64))
        self.queries = (desc, rk, offs)

    def detectLoopClosureID(self):
        """-> list of dict(min_dist, min_dist_sc, min_dist_odom, yaw_diff_rad, nn_idx, argmin_shift, aug_idx, aug_xy)."""
        desc, rk, offs = self.queries
        cur = len(self.ringkeys) - 1
        nq = len(offs)
        ci, cs, ne = sc_search(self.ctx, np.stack(self.ringkeys), np.stack(self.odom), rk, np.full(nq, cur, np.int32), self.par)
        if len(self.ringkeys) < ne[0] + 1:
            return []
        pairs = [(q, int(ci[q, t]), float(cs[q, t])) for q in range(nq) for t in range(ci.shape[1]) if ci[q, t] >= 0]
        if not pairs:
            return []
        uniq = sorted({c for _, c, _ in pairs})
        pos = {c: i for i, c in enumerate(uniq)}
        dist, shift = sc_distance_batch(self.ctx, desc, np.stack([self.polarcontexts[c] for c in uniq]), [q for q, _, _ in pairs],
                                        [pos[c] for _, c, _ in pairs], self.par)
        similar = []
        unit = 360.0 / float(self.par.num_sector)
        for (q, c, sim), d_sc, sh in zip(pairs, dist, shift):
            d_odom = sim if self.par.odometry_coupled_closure else 0.0
            deg = np.float32(int(sh) * unit)
            similar.append(dict(min_dist=d_sc + d_odom if self.par.odometry_coupled_closure else d_sc, min_dist_sc=float(d_sc), min_dist_odom=d_odom,
                                yaw_diff_rad=float(np.float32(float(deg) * np.pi / 180.0)), nn_idx=c, argmin_shift=int(sh), aug_idx=q, aug_xy=offs[q]))
            similar.sort(key=lambda s: s["min_dist"])   # stable, like the oracle's restatement of the running std::sort
            if len(similar) > self.par.n_candidates:
                similar.pop()
        return similar


class CoralParams(C.Structure):
    _fields_ = [("radius", C.c_double), ("weight_res_intensity", C.c_int), ("overlap_req", C.c_int)]


class CoralResult(C.Structure):
    _fields_ = [("joint", C.c_double), ("sep", C.c_double), ("overlap", C.c_double), ("count_valid", C.c_int), ("merged_size", C.c_int),
                ("valid", C.c_int)]


def CorAlRadarQuality(ctx: Context, clouds, src_cloud, ref_cloud, T_src, T_ref, T_offset=None, radius=1.0, weight_res_intensity=False,
                      per_point=False):
    """CorAlRadarQuality for a batch of pairs (AlignmentQuality.cpp:99-229 via alignmentinterface.cpp:437-454).
    clouds: list of (x, y, intensity) float32 arrays in the scans' own frames; pair p = (src_cloud[p] at T_src[p] * T_offset[p],
    ref_cloud[p] at T_ref[p]).  Returns a list of CoralResult (and the per-point [sum merged_size, 3] array if asked)."""
    arrs = [[np.ascontiguousarray(a, np.float32) for a in c] for c in clouds]
    n_pts = np.array([len(c[0]) for c in arrs], np.int32)
    ptr = lambda k: (C.c_void_p * len(arrs))(*[c[k].ctypes.data_as(C.c_void_p).value for c in arrs])
    sc = np.ascontiguousarray(src_cloud, np.int32)
    rc_ = np.ascontiguousarray(ref_cloud, np.int32)
    n = len(sc)
    Ts = np.ascontiguousarray(T_src, np.float64).reshape(n, 3)
    Tr = np.ascontiguousarray(T_ref, np.float64).reshape(n, 3)
    To = np.ascontiguousarray(T_offset, np.float64).reshape(n, 3) if T_offset is not None else None
    res = (CoralResult * n)()
    par = CoralParams(radius, int(weight_res_intensity), 1)
    pp = np.zeros((int(sum(n_pts[s] + n_pts[r] for s, r in zip(sc, rc_))), 3)) if per_point else None
    _check(lib().tbv_coral_quality_batch(ctx.h, len(arrs), ptr(0), ptr(1), ptr(2), _ptr(n_pts), n, _ptr(sc), _ptr(rc_), _ptr(Ts), _ptr(To), _ptr(Tr),
                                         C.byref(par), res, _ptr(pp)))
    return (list(res), pp) if per_point else list(res)
