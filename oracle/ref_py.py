"""ctypes binding of oracle/_ref/libtbv_ref_filters.so — the reference's OWN filter sources (radar_filters.cpp, cfar.cpp),
compiled unmodified from /root/reference against the container stand-ins of oracle/ref_shim/.  *** TEST INFRASTRUCTURE ONLY ***

Used to pin the oracle (tests/test_reference_pin_cpu.py), to generate tests/golden/ (tests/golden/make_golden.py) and as the
`kind: "reference"` Filtering-stage timing of bench.py --impl reference.  The library exists only where it was built
(`make -C oracle ref`, needs /root/reference) or where the built .so travelled to; `available()` says which.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libtbv_ref_filters.so")
REFERENCE_ROOT = "/root/reference"
_LIB = None


def build() -> str | None:
    """Builds oracle/_ref when the reference sources are present (build container); otherwise leaves a prebuilt one alone."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "cfear_radarodometry")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return SO if os.path.exists(SO) else None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.tbv_ref_kstrongest.restype = C.c_int
        L.tbv_ref_kstrongest_both.restype = C.c_int
        L.tbv_ref_cacfar.restype = C.c_int
        L.tbv_ref_kstrongest_many.restype = C.c_long
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def kstrongest(img: np.ndarray, z_min=60.0, k=40, min_distance=2.5, range_res=0.0438):
    """radarDriver::Process's k-strongest branch on one scan: returns dict(filtered=(x, y, I), peaks=(x, y, I)) float32."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    n_az, n_range = img.shape
    cap = n_az * max(k, 1)
    f = [np.zeros(cap, np.float32) for _ in range(3)]
    p = [np.zeros(cap, np.float32) for _ in range(3)]
    nf, npk = C.c_int(0), C.c_int(0)
    lib().tbv_ref_kstrongest_both(_p(img, C.c_uint8), n_az, n_range, C.c_long(n_range), C.c_float(z_min), int(k), C.c_float(min_distance),
                                  C.c_float(range_res), _p(f[0], C.c_float), _p(f[1], C.c_float), _p(f[2], C.c_float), cap, C.byref(nf),
                                  _p(p[0], C.c_float), _p(p[1], C.c_float), _p(p[2], C.c_float), cap, C.byref(npk))
    assert nf.value <= cap and npk.value <= cap
    return {"filtered": tuple(a[:nf.value] for a in f), "peaks": tuple(a[:npk.value] for a in p)}


def cacfar(img, window_size=40, false_alarm_rate=0.01, nb_guard_cells=10, range_res=0.0438, static_threshold=20.0, min_distance=2.5,
           max_distance=400.0):
    """radarDriver::Process's CA-CFAR branch (float parameters widened at the call, as radarDriver::Parameters does)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    n_az, n_range = img.shape
    cap = n_az * n_range
    x, y, I = (np.zeros(cap, np.float32) for _ in range(3))
    n = lib().tbv_ref_cacfar(_p(img, C.c_uint8), n_az, n_range, C.c_long(n_range), int(window_size), C.c_float(false_alarm_rate), int(nb_guard_cells),
                             C.c_float(range_res), C.c_float(static_threshold), C.c_float(min_distance), C.c_double(max_distance),
                             _p(x, C.c_float), _p(y, C.c_float), _p(I, C.c_float), cap)
    return x[:n], y[:n], I[:n]


def kstrongest_many(imgs: np.ndarray, z_min=60.0, k=40, min_distance=2.5, range_res=0.0438) -> int:
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n, n_az, n_range = imgs.shape
    return int(lib().tbv_ref_kstrongest_many(_p(imgs, C.c_uint8), n, n_az, n_range, C.c_float(z_min), int(k), C.c_float(min_distance), C.c_float(range_res)))


def statistics(pairs) -> str:
    """CFEAR_Radarodometry::statistics: Document every (name, value) of `pairs`, return GetStatistics() (statistics.cpp:40-51)."""
    names = b"".join(n.encode() + b"\0" for n, _ in pairs)
    vals = np.ascontiguousarray([v for _, v in pairs], np.float64)
    out = C.create_string_buffer(1 << 16)
    lib().tbv_ref_statistics.restype = C.c_int
    lib().tbv_ref_statistics(names, _p(vals, C.c_double), len(pairs), out, len(out))
    return out.value.decode()
