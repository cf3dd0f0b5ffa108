"""ctypes binding of the CPU oracle (oracle/liboracle.so).  *** TEST INFRASTRUCTURE ONLY ***

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (tbv_slam_public_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P2P, P2L, P2D = 0, 1, 2
LOSS_NONE, HUBER, CAUCHY, SOFTLONE, COMBINED, TUKEY = 0, 1, 2, 3, 4, 5
W_UNIFORM, W_SIM_N, W_SIM_DIR, W_SIM_SCALE, W_COMBINED = 0, 1, 2, 3, 4
VOXEL_ORDER_STABLE, VOXEL_ORDER_STD_SORT = 0, 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "tbv_oracle.hpp", "tbv_oracle_reg.hpp", "tbv_oracle_loop.hpp",
                                             "oracle_capi_loop.inc", "tbv_oracle_coral.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


class RegParams(C.Structure):
    _fields_ = [("cost", C.c_int), ("loss", C.c_int), ("weight_opt", C.c_int), ("loss_limit", C.c_double),
                ("cov_scale", C.c_double), ("regularization", C.c_double), ("max_itr_association", C.c_int),
                ("max_itr_solver", C.c_int)]


class RegSummary(C.Structure):
    _fields_ = [("success", C.c_int), ("itrs", C.c_int), ("lm_iterations", C.c_int), ("num_residuals", C.c_int),
                ("last_n_iterations", C.c_int), ("termination", C.c_int), ("score", C.c_double), ("final_cost", C.c_double),
                ("last_relative_decrease", C.c_double)]


class OdomParams(C.Structure):
    _fields_ = [("z_min", C.c_float), ("k_strongest", C.c_int), ("min_distance", C.c_float), ("range_res", C.c_float),
                ("cost_type", C.c_int), ("loss_type", C.c_int), ("weight_opt", C.c_int),
                ("loss_limit", C.c_double), ("covar_scale", C.c_double), ("regularization", C.c_double),
                ("submap_scan_size", C.c_int), ("weight_intensity", C.c_int), ("use_guess", C.c_int), ("compensate", C.c_int),
                ("radar_ccw", C.c_int), ("use_keyframe", C.c_int),
                ("res", C.c_double), ("min_keyframe_dist", C.c_double), ("min_keyframe_rot_deg", C.c_double),
                ("downsample_factor", C.c_double), ("voxel_order", C.c_int)]


class OdomOut(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("n_points", C.c_int), ("n_cells", C.c_int), ("itrs", C.c_int), ("reg_ok", C.c_int),
                ("is_keyframe", C.c_int), ("n_keyframes", C.c_int), ("ms_filter", C.c_double),
                ("ms_compensate_normals", C.c_double), ("ms_register", C.c_double)]


class SCParams(C.Structure):
    _fields_ = [("num_ring", C.c_int), ("num_sector", C.c_int), ("max_radius", C.c_double), ("search_ratio", C.c_double),
                ("num_candidates_from_tree", C.c_int), ("n_candidates", C.c_int), ("odom_sigma_error", C.c_double),
                ("odometry_coupled_closure", C.c_int), ("augment_sc", C.c_int), ("no_point", C.c_double),
                ("desc_function", C.c_int), ("desc_divider", C.c_double)]


class PGOParams(C.Structure):
    _fields_ = [("odom_vxx", C.c_double), ("odom_vyy", C.c_double), ("odom_vtt", C.c_double), ("loop_scaling", C.c_double),
                ("replace_cov_by_identity", C.c_int), ("loop_cauchy", C.c_double)]


def default_sc_params(**kw) -> SCParams:
    # TBV-8 offline settings: tbv_slam/src/tbv_slam_offline.cpp:81-101 (sum / 1000, 3 candidates, augmentations)
    p = SCParams(40, 120, 80.0, 0.1, 10, 3, 0.05, 1, 1, 0.0, 0, 1000.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_pgo_params(**kw) -> PGOParams:
    p = PGOParams(0.01, 0.01, 0.001, 500000.0, 1, 0.1)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_reg_params(**kw) -> RegParams:
    p = RegParams(P2L, HUBER, W_UNIFORM, 0.1, 1.0, 0.01, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_odom_params(**kw) -> OdomParams:
    # BASELINE config 2: CFEAR-3 filter (k=40, z_min=60, r=3), 4 keyframes, P2L, Huber 0.1, weight_opt 4, weight_intensity
    p = OdomParams(60.0, 40, 2.5, 0.0438, P2L, HUBER, W_COMBINED, 0.1, 1.0, 1.0, 4, 1, 1, 1, 0, 1, 3.0, 1.5, 5.0, 1.0,
                   VOXEL_ORDER_STABLE)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_kstrongest.restype = C.c_int
        L.orc_cacfar.restype = C.c_int
        L.orc_build_cells.restype = C.c_int
        L.orc_voxel_centroids.restype = C.c_int
        L.orc_closest_idx.restype = C.c_int
        L.orc_register.restype = C.c_int
        L.orc_get_cost.restype = C.c_int
        L.orc_pair_normal_eq.restype = C.c_int
        L.orc_loop_register.restype = C.c_int
        L.orc_odom_create.restype = C.c_void_p
        L.orc_odom_run.restype = C.c_double
        L.orc_odom_run_timed.restype = C.c_double
        L.orc_odom_keyframes.restype = C.c_int
        L.orc_odom_keyframe_cells.restype = C.c_int
        L.orc_rsc_create.restype = C.c_void_p
        L.orc_rsc_detect.restype = C.c_int
        L.orc_rsc_state.restype = C.c_int
        L.orc_rsc_search.restype = C.c_int
        L.orc_sc_dist_direct.restype = C.c_double
        L.orc_pgo_assemble.restype = C.c_double
        L.orc_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def kstrongest(img: np.ndarray, z_min=60.0, k=40, min_distance=2.5, range_res=0.0438, peaks=True):
    """Returns dict(filtered=(az,rg,I,x,y), peaks=(...)) for one scan [n_az, n_range] u8."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    n_az, n_range = img.shape
    cap = n_az * k

    def bufs():
        return (np.zeros(cap, np.uint16), np.zeros(cap, np.uint16), np.zeros(cap, np.uint8), np.zeros(cap, np.float32),
                np.zeros(cap, np.float32))
    f, pk = bufs(), bufs()
    npk = C.c_int(0)
    n = lib().orc_kstrongest(_p(img, C.c_uint8), n_az, n_range, C.c_long(n_range), C.c_float(z_min), k, C.c_float(min_distance),
                             C.c_float(range_res), cap, _p(f[0], C.c_uint16), _p(f[1], C.c_uint16), _p(f[2], C.c_uint8),
                             _p(f[3], C.c_float), _p(f[4], C.c_float), C.byref(npk) if peaks else None,
                             _p(pk[0], C.c_uint16), _p(pk[1], C.c_uint16), _p(pk[2], C.c_uint8), _p(pk[3], C.c_float),
                             _p(pk[4], C.c_float))
    out = {"filtered": tuple(a[:n] for a in f)}
    if peaks:
        out["peaks"] = tuple(a[:npk.value] for a in pk)
    return out


def cacfar(img, window_size=40, false_alarm_rate=0.01, nb_guard_cells=10, range_res=0.0438, static_threshold=20.0,
           min_distance=2.5, max_distance=400.0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    n_az, n_range = img.shape
    cap = n_az * n_range
    az, rg, I = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16), np.zeros(cap, np.uint8)
    x, y = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    n = lib().orc_cacfar(_p(img, C.c_uint8), n_az, n_range, C.c_long(n_range), window_size, C.c_double(false_alarm_rate),
                         nb_guard_cells, C.c_double(range_res), C.c_double(static_threshold), C.c_double(min_distance),
                         C.c_double(max_distance), cap, _p(az, C.c_uint16), _p(rg, C.c_uint16), _p(I, C.c_uint8),
                         _p(x, C.c_float), _p(y, C.c_float))
    return az[:n], rg[:n], I[:n], x[:n], y[:n]


def rotate90ccw(src):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape
    dst = np.zeros((W, H), np.uint8)
    lib().orc_rotate90ccw(_p(src, C.c_uint8), H, W, _p(dst, C.c_uint8))
    return dst


def compensate(x, y, mot, ccw=False):
    x = np.array(x, dtype=np.float32, copy=True)
    y = np.array(y, dtype=np.float32, copy=True)
    m = np.asarray(mot, dtype=np.float64)
    lib().orc_compensate(_p(x, C.c_float), _p(y, C.c_float), len(x), _p(m, C.c_double), int(ccw))
    return x, y


def build_cells(x, y, intensity, radius=3.0, downsample_factor=1.0, weight_intensity=True, origin=(0.0, 0.0),
                voxel_order=VOXEL_ORDER_STABLE, cap=None):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    intensity = np.ascontiguousarray(intensity, np.float32)
    cap = cap or max(len(x), 1)
    recs = np.zeros((cap, 16), np.float64)
    o = np.asarray(origin, np.float64)
    ns = C.c_int(0)
    n = lib().orc_build_cells(_p(x, C.c_float), _p(y, C.c_float), _p(intensity, C.c_float), len(x), C.c_float(radius),
                              C.c_double(downsample_factor), int(weight_intensity), _p(o, C.c_double), voxel_order, cap,
                              _p(recs, C.c_double), C.byref(ns))
    return recs[:n].copy(), ns.value


def voxel_centroids(x, y, intensity, leaf=3.0, voxel_order=VOXEL_ORDER_STABLE):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    intensity = np.ascontiguousarray(intensity, np.float32)
    cap = max(len(x), 1)
    cx, cy, ci = np.zeros(cap, np.float32), np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    n = lib().orc_voxel_centroids(_p(x, C.c_float), _p(y, C.c_float), _p(intensity, C.c_float), len(x), C.c_float(leaf),
                                  voxel_order, cap, _p(cx, C.c_float), _p(cy, C.c_float), _p(ci, C.c_float))
    return cx[:n], cy[:n], ci[:n]


def eig2(m00, m10, m11):
    ev, evec = np.zeros(2), np.zeros(4)
    lib().orc_eig2(C.c_double(m00), C.c_double(m10), C.c_double(m11), _p(ev, C.c_double), _p(evec, C.c_double))
    return ev, evec.reshape(2, 2)


def closest_idx(recs, px, py, d, brute=False):
    recs = np.ascontiguousarray(recs, np.float64)
    return lib().orc_closest_idx(_p(recs, C.c_double), len(recs), C.c_float(0), C.c_double(px), C.c_double(py), C.c_double(d),
                                 int(brute))


def _scan_ptrs(scans):
    arrs = [np.ascontiguousarray(s, np.float64).reshape(-1, 16) for s in scans]
    ptrs = (C.POINTER(C.c_double) * len(arrs))(*[_p(a, C.c_double) for a in arrs])
    ns = np.array([len(a) for a in arrs], np.int32)
    return arrs, ptrs, ns


def register(scans, T, params: RegParams | None = None):
    """scans: list of [n,16] cell records (last = moving); T: [n_scans,3]. Returns (T_out, summary)."""
    params = params or default_reg_params()
    arrs, ptrs, ns = _scan_ptrs(scans)
    Tio = np.array(T, dtype=np.float64, copy=True).reshape(len(scans), 3)
    s = RegSummary()
    lib().orc_register(len(scans), ptrs, _p(ns, C.c_int), _p(Tio, C.c_double), C.byref(params), C.byref(s))
    return Tio, s


def get_cost(scans, T, params: RegParams | None = None, itr=0):
    params = params or default_reg_params()
    arrs, ptrs, ns = _scan_ptrs(scans)
    Tio = np.ascontiguousarray(T, np.float64).reshape(len(scans), 3)
    cap = int(2 * ns[-1] * (len(scans) - 1)) + 8
    res = np.zeros(cap)
    score, cost = C.c_double(0), C.c_double(0)
    n = lib().orc_get_cost(len(scans), ptrs, _p(ns, C.c_int), _p(Tio, C.c_double), C.byref(params), itr, C.byref(score),
                           C.byref(cost), cap, _p(res, C.c_double))
    return n, score.value, cost.value, res[:max(n, 0)]


def cost_samples(scans, T, params: RegParams | None = None, itr=2, xy_range=0.4, yaw_range=0.0043625, n_per_axis=3):
    """approximateCovarianceBySampling's sampling half: [n^3, 4] = (dx, dy, dyaw, cost)."""
    params = params or default_reg_params()
    arrs, ptrs, ns = _scan_ptrs(scans)
    Tio = np.ascontiguousarray(T, np.float64).reshape(len(scans), 3)
    out = np.zeros((n_per_axis ** 3, 4))
    lib().orc_cost_samples.restype = C.c_int
    n = lib().orc_cost_samples(len(scans), ptrs, _p(ns, C.c_int), _p(Tio, C.c_double), C.byref(params), itr, C.c_double(xy_range),
                               C.c_double(yaw_range), n_per_axis, _p(out, C.c_double))
    assert n == len(out)
    return out


def pair_normal_eq(tgt, T_tgt, src, T_src, params: RegParams | None = None, itr=1):
    params = params or default_reg_params()
    tgt = np.ascontiguousarray(tgt, np.float64); src = np.ascontiguousarray(src, np.float64)
    Tt = np.asarray(T_tgt, np.float64); Ts = np.asarray(T_src, np.float64)
    H, g = np.zeros(9), np.zeros(3)
    cost, nres = C.c_double(0), C.c_int(0)
    assoc = np.zeros(len(src), np.int32)
    nb = lib().orc_pair_normal_eq(_p(tgt, C.c_double), len(tgt), _p(Tt, C.c_double), _p(src, C.c_double), len(src),
                                  _p(Ts, C.c_double), C.byref(params), itr, C.byref(cost), C.byref(nres), _p(H, C.c_double),
                                  _p(g, C.c_double), _p(assoc, C.c_int))
    return dict(n_blocks=nb, n_res=nres.value, cost=cost.value, H=H.reshape(3, 3), g=g, assoc=assoc)


def loop_register(cells_from, cells_to, Tfrom, Tto):
    f = np.ascontiguousarray(cells_from, np.float64); t = np.ascontiguousarray(cells_to, np.float64)
    Tf = np.asarray(Tfrom, np.float64); Tt = np.asarray(Tto, np.float64)
    Ta, Tr = np.zeros(3), np.zeros(3)
    itrs, score = C.c_int(0), C.c_double(0)
    ok = lib().orc_loop_register(_p(f, C.c_double), len(f), _p(t, C.c_double), len(t), _p(Tf, C.c_double), _p(Tt, C.c_double),
                                 _p(Ta, C.c_double), _p(Tr, C.c_double), C.byref(itrs), C.byref(score))
    return bool(ok), Ta, Tr, itrs.value, score.value


class Odometry:
    def __init__(self, params: OdomParams | None = None):
        self.params = params or default_odom_params()
        self.h = C.c_void_p(lib().orc_odom_create(C.byref(self.params)))

    def step(self, img: np.ndarray) -> OdomOut:
        img = np.ascontiguousarray(img, np.uint8)
        o = OdomOut()
        lib().orc_odom_step(self.h, _p(img, C.c_uint8), img.shape[0], img.shape[1], C.c_long(img.shape[1]), C.byref(o))
        return o

    def keyframes(self):
        poses = np.zeros((64, 3)); nc = np.zeros(64, np.int32)
        n = lib().orc_odom_keyframes(self.h, 64, _p(poses, C.c_double), _p(nc, C.c_int))
        return poses[:n], nc[:n]

    def keyframe_cells(self, kf: int):
        n = lib().orc_odom_keyframe_cells(self.h, kf, 0, None)
        recs = np.zeros((max(n, 1), 16))
        lib().orc_odom_keyframe_cells(self.h, kf, n, _p(recs, C.c_double))
        return recs[:n]

    def __del__(self):
        try:
            lib().orc_odom_destroy(self.h)
        except Exception:
            pass


def odom_run(params: OdomParams, pool: np.ndarray, first, n_frames: int, n_threads: int):
    """Run len(first) independent sequences over a pool of scans on n_threads host threads. Returns (seconds, poses)."""
    pool = np.ascontiguousarray(pool, np.uint8)
    first = np.ascontiguousarray(first, np.int32)
    poses = np.zeros((len(first), n_frames, 3))
    sec = lib().orc_odom_run(C.byref(params), _p(pool, C.c_uint8), pool.shape[0], pool.shape[1], pool.shape[2], len(first),
                             _p(first, C.c_int), n_frames, n_threads, _p(poses, C.c_double))
    return sec, poses


def odom_run_timed(params: OdomParams, pool: np.ndarray, first, n_warm: int, n_frames: int, n_threads: int):
    """Like odom_run but times only frames >= n_warm of every sequence. Returns (max-over-threads timed seconds, poses)."""
    pool = np.ascontiguousarray(pool, np.uint8)
    first = np.ascontiguousarray(first, np.int32)
    poses = np.zeros((len(first), n_frames, 3))
    sec = lib().orc_odom_run_timed(C.byref(params), _p(pool, C.c_uint8), pool.shape[0], pool.shape[1], pool.shape[2], len(first),
                                   _p(first, C.c_int), n_warm, n_frames, n_threads, _p(poses, C.c_double), None)
    return sec, poses


def hardware_threads() -> int:
    return lib().orc_hardware_threads()


def loss(kind, loss_limit, weight, s):
    rho = np.zeros(3)
    lib().orc_loss(kind, C.c_double(loss_limit), C.c_double(weight), C.c_double(s), _p(rho, C.c_double))
    return rho


# ---- scan context -------------------------------------------------------------------------------------------
def sc_make(x, y, intensity, params: SCParams | None = None, off=(0.0, 0.0)):
    params = params or default_sc_params()
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    intensity = np.ascontiguousarray(intensity, np.float32)
    desc = np.zeros(params.num_ring * params.num_sector)
    rk = np.zeros(params.num_ring, np.float32); sk = np.zeros(params.num_sector)
    lib().orc_sc_make(_p(x, C.c_float), _p(y, C.c_float), _p(intensity, C.c_float), len(x), C.byref(params), C.c_double(off[0]),
                      C.c_double(off[1]), _p(desc, C.c_double), _p(rk, C.c_float), _p(sk, C.c_double))
    return desc, rk, sk


def sc_distance(sc1, sc2, R=40, S=120, search_ratio=0.1):
    sc1 = np.ascontiguousarray(sc1, np.float64); sc2 = np.ascontiguousarray(sc2, np.float64)
    d, sh = C.c_double(0), C.c_int(0)
    lib().orc_sc_distance(_p(sc1, C.c_double), _p(sc2, C.c_double), R, S, C.c_double(search_ratio), C.byref(d), C.byref(sh))
    return d.value, sh.value


class RSC:
    def __init__(self, params: SCParams | None = None):
        self.params = params or default_sc_params()
        self.h = C.c_void_p(lib().orc_rsc_create(C.byref(self.params)))

    def add(self, x, y, intensity, Todom):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        intensity = np.ascontiguousarray(intensity, np.float32)
        T = np.asarray(Todom, np.float64)
        lib().orc_rsc_add(self.h, _p(x, C.c_float), _p(y, C.c_float), _p(intensity, C.c_float), len(x), _p(T, C.c_double))

    def detect(self):
        out = np.zeros((64, 8))
        n = lib().orc_rsc_detect(self.h, 64, _p(out, C.c_double))
        return out[:n]

    def state(self, cap=100000):
        sim = np.zeros(cap)
        ne = C.c_int(0)
        n = lib().orc_rsc_state(self.h, C.byref(ne), cap, _p(sim, C.c_double))
        return ne.value, sim[:n]

    def search(self, key):
        key = np.ascontiguousarray(key, np.float32)
        idx = np.zeros(64, np.int32)
        n = lib().orc_rsc_search(self.h, _p(key, C.c_float), 64, _p(idx, C.c_int))
        return idx[:n]

    def __del__(self):
        try:
            lib().orc_rsc_destroy(self.h)
        except Exception:
            pass


# ---- pose graph ----------------------------------------------------------------------------------------------
def coral_quality(src, ref, Tsrc, Tref, Toffset=(0.0, 0.0, 0.0), radius=1.0, weight_res_intensity=False, per_point=False):
    """CorAlRadarQuality (AlignmentQuality.cpp:99-229): src / ref = (x, y, intensity) float32 peaks clouds in their own frames.
    Returns dict(joint, sep, overlap, count_valid, merged_size, valid[, per_point [n, 3] = sep, joint, valid])."""
    sx, sy, si = (np.ascontiguousarray(a, np.float32) for a in src)
    rx, ry, ri = (np.ascontiguousarray(a, np.float32) for a in ref)
    out = np.zeros(6)
    pp = np.zeros((len(sx) + len(rx), 3)) if per_point else None
    a3 = lambda t: np.ascontiguousarray(t, np.float64)
    lib().orc_coral_quality(_p(sx, C.c_float), _p(sy, C.c_float), _p(si, C.c_float), len(sx), _p(rx, C.c_float), _p(ry, C.c_float),
                            _p(ri, C.c_float), len(rx), _p(a3(Tsrc), C.c_double), _p(a3(Toffset), C.c_double), _p(a3(Tref), C.c_double),
                            C.c_double(radius), int(weight_res_intensity), _p(out, C.c_double), _p(pp, C.c_double))
    r = dict(joint=out[0], sep=out[1], overlap=out[2], count_valid=int(out[3]), merged_size=int(out[4]), valid=bool(out[5]))
    if per_point:
        r["per_point"] = pp
    return r


def pgo_assemble(nodes, ids, meas, params: PGOParams | None = None, info=None, fixed_node=0):
    params = params or default_pgo_params()
    nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, 7)
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    meas = np.ascontiguousarray(meas, np.float64).reshape(-1, 7)
    n, m = len(nodes), len(ids)
    Hd, Ho, g, res = np.zeros((n, 36)), np.zeros((m, 36)), np.zeros((n, 6)), np.zeros((m, 6))
    inf = np.ascontiguousarray(info, np.float64) if info is not None else None
    cost = lib().orc_pgo_assemble(n, _p(nodes, C.c_double), m, _p(ids, C.c_int), _p(meas, C.c_double), _p(inf, C.c_double),
                                  C.byref(params), fixed_node, _p(Hd, C.c_double), _p(Ho, C.c_double), _p(g, C.c_double),
                                  _p(res, C.c_double))
    return cost, Hd.reshape(n, 6, 6), Ho.reshape(m, 6, 6), g, res


def pgo_residual(a, b, meas, ctype=0, params: PGOParams | None = None):
    params = params or default_pgo_params()
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); meas = np.ascontiguousarray(meas, np.float64)
    r, Ja, Jb = np.zeros(6), np.zeros(36), np.zeros(36)
    lib().orc_pgo_residual(_p(a, C.c_double), _p(b, C.c_double), _p(meas, C.c_double), ctype, C.byref(params), _p(r, C.c_double),
                           _p(Ja, C.c_double), _p(Jb, C.c_double))
    return r, Ja.reshape(6, 6), Jb.reshape(6, 6)
