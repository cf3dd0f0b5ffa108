"""Shared by tests/test_ref_downstream_{cpu,gpu}.py: recomputes, with a given backend (the oracle, or the CUDA path through the C-ABI), every
record tools/ref_fixtures/dump_ref_downstream.cpp writes with the REFERENCE's own classes, and compares the two sets.

Tolerances (north_star): filtered point indices bit-exact -> compensated float clouds <= 1 float ulp; cells: same count and order, mean / cov
1e-9; Register: same success and association rounds, pose 1e-5 m / 1e-6 rad; GetCost: same residual count, cost 1e-9 relative; Scan Context:
descriptor 1e-9, candidate index and shift exact."""
import os

import numpy as np

from tbv_slam_public_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "ref_downstream.npz")
N_SCANS = 6


def _mul(a, b):
    c, s = np.cos(a[2]), np.sin(a[2])
    return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], a[2] + b[2]])


def _inv(a):
    c, s = np.cos(a[2]), np.sin(a[2])
    return np.array([-(c * a[0] + s * a[1]), -(-s * a[0] + c * a[1]), -a[2]])


def compute_records(B):
    """B: backend with kstrongest(scan) -> {'filtered': (x, y, I), 'peaks': (x, y, I)}, compensate(x, y, mot), build_cells(x, y, I) -> [n, 16]
    tbv_cell records, register(scans, T, loop) -> (ok, pose, score, itrs), get_cost(scans, T) -> (ok, cost, residuals),
    sc(x, y, I, T) -> (descriptor [40, 120], candidates [(nn_idx, shift, dist, dist_sc, dist_odom, yaw)])."""
    st = synth.make_stream(N_SCANS)
    rec, cells, T = {}, [], [np.array(p, np.float64) for p in st.gt]
    for s in range(N_SCANS):
        r = B.kstrongest(st.scans[s])
        mot = _mul(_inv(T[s - 1]), T[s]) if s > 0 else np.zeros(3)
        clouds = {}
        for key, tag in (("filtered", "cloud"), ("peaks", "peaks")):
            x, y, I = r[key]
            x, y = B.compensate(x, y, mot)
            clouds[key] = (x, y, I)
            rec[f"{tag}_{s}"] = np.stack([x, y, I], axis=1).astype(np.float64).reshape(-1)
        c = B.build_cells(*clouds["filtered"])
        cells.append(c)
        # dumper layout: u(2) cov(4: 00 01 10 11) scale snormal(2) lambda_min lambda_max sum_intensity avg_intensity Nsamples  (tbv_cell: see include/tbv_b200.h)
        rec[f"cells_{s}"] = B.cells_as_reference_layout(c).reshape(-1)
        rec[f"_peaks_{s}"] = clouds["peaks"]
    for s in range(1, N_SCANS):
        lo = max(0, s - 4)
        scans = cells[lo:s] + [cells[s]]
        Ts = [T[t] for t in range(lo, s)] + [_mul(T[s], np.array([0.3, -0.2, 0.02]))]
        for loop, tag in ((False, "register"), (True, "register_loop")):
            ok, pose, score, itrs = B.register(scans, np.array(Ts), loop)
            rec[f"{tag}_{s}"] = np.array([float(ok), pose[0], pose[1], pose[2], score, float(itrs)])
        ok, cost, res = B.get_cost([cells[s - 1], cells[s]], np.array([T[s - 1], T[s]]))
        rec[f"get_cost_{s}"] = np.r_[float(ok), cost, float(len(res)), res]
    for s, (desc, cand) in enumerate(B.scan_context([rec[f"_peaks_{s}"] for s in range(N_SCANS)], T)):
        rec[f"sc_desc_{s}"] = np.asarray(desc, np.float64).reshape(40, 120).T.reshape(-1)     # column-major, as Eigen stores it
        rec[f"sc_cand_{s}"] = np.array([v for k in cand for v in k], np.float64)
    return {k: v for k, v in rec.items() if not k.startswith("_")}


def _ang(d):
    return np.abs(np.arctan2(np.sin(d), np.cos(d)))


def compare(got, ref):
    """Raises AssertionError with the first record that leaves its tolerance."""
    for s in range(N_SCANS):
        for tag in ("cloud", "peaks"):
            g, r = got[f"{tag}_{s}"].reshape(-1, 3), ref[f"{tag}_{s}"].reshape(-1, 3)
            assert g.shape == r.shape, (tag, s, g.shape, r.shape)
            assert np.array_equal(g[:, 2], r[:, 2]), (tag, s, "intensities / order")
            ulp = np.spacing(np.abs(r[:, :2]).astype(np.float32)).astype(np.float64)
            assert np.all(np.abs(g[:, :2] - r[:, :2]) <= ulp), (tag, s, "more than one float ulp")
        g, r = got[f"cells_{s}"].reshape(-1, 14), ref[f"cells_{s}"].reshape(-1, 14)
        assert g.shape == r.shape, ("cells", s, g.shape, r.shape)
        assert np.array_equal(g[:, 13], r[:, 13]) and np.allclose(g[:, :6], r[:, :6], rtol=0, atol=1e-9), ("cells", s)
        assert np.allclose(g[:, 6], r[:, 6], rtol=1e-7) and np.allclose(g[:, 7:9], r[:, 7:9], atol=1e-7) and np.allclose(g[:, 9:13], r[:, 9:13], rtol=1e-7, atol=1e-9), ("cells", s)
    for s in range(1, N_SCANS):
        for tag in ("register", "register_loop"):
            g, r = got[f"{tag}_{s}"], ref[f"{tag}_{s}"]
            assert g[0] == r[0] and g[5] == r[5], (tag, s, "success / association rounds", g, r)
            assert np.abs(g[1:3] - r[1:3]).max() < 1e-5 and _ang(g[3] - r[3]) < 1e-6, (tag, s, g, r)
            assert abs(g[4] - r[4]) <= 1e-7 * abs(r[4]) + 1e-12, (tag, s, "score")
        g, r = got[f"get_cost_{s}"], ref[f"get_cost_{s}"]
        assert g[0] == r[0] and g[2] == r[2] and abs(g[1] - r[1]) <= 1e-9 * abs(r[1]), ("get_cost", s, g[:3], r[:3])
        assert np.allclose(g[3:], r[3:], rtol=0, atol=1e-9), ("get_cost residuals", s)
    for s in range(N_SCANS):
        assert np.allclose(got[f"sc_desc_{s}"], ref[f"sc_desc_{s}"], rtol=0, atol=1e-9), ("sc_desc", s)
        g, r = got[f"sc_cand_{s}"].reshape(-1, 6), ref[f"sc_cand_{s}"].reshape(-1, 6)
        assert g.shape == r.shape and np.array_equal(g[:, :2], r[:, :2]) and np.allclose(g[:, 2:], r[:, 2:], atol=1e-9), ("sc_cand", s)


class OracleBackend:
    def __init__(self, O):
        self.O = O

    def kstrongest(self, scan):
        r = self.O.kstrongest(scan, peaks=True)
        return {k: (r[k][3], r[k][4], r[k][2].astype(np.float32)) for k in ("filtered", "peaks")}

    def compensate(self, x, y, mot):
        return self.O.compensate(x, y, mot, False) if len(x) else (x, y)

    def build_cells(self, x, y, I):
        return self.O.build_cells(x, y, I, radius=3.0, weight_intensity=True)[0]

    @staticmethod
    def cells_as_reference_layout(c):
        # tbv_cell (16 doubles): u0 u1 c00 c01 c10 c11 planarity n0 n1 o0 o1 lambda_min lambda_max sum_w avg_w N
        c = np.asarray(c, np.float64).reshape(-1, 16)
        return np.c_[c[:, 0:6], c[:, 6], c[:, 7:9], c[:, 11], c[:, 12], c[:, 13], c[:, 14], c[:, 15]]

    def register(self, scans, T, loop):
        O = self.O
        P = O.default_reg_params(weight_opt=O.W_UNIFORM, max_itr_association=4, max_itr_solver=10) if loop else O.default_reg_params(weight_opt=O.W_COMBINED)
        To, s = O.register(scans, T, P)
        return bool(s.success), To[-1], s.score, s.itrs

    def get_cost(self, scans, T):
        n, score, cost, res = self.O.get_cost(scans, T, self.O.default_reg_params(loss_limit=0.3), itr=0)
        return n > 1, cost, res

    def scan_context(self, peaks, T):
        rsc, out = self.O.RSC(), []
        for (x, y, I), t in zip(peaks, T):
            rsc.add(x, y, I, t)
            desc = self.O.sc_make(x, y, I)[0]
            out.append((desc, [(c["nn_idx"], c["argmin_shift"], c["min_dist"], c["min_dist_sc"], c["min_dist_odom"], c["yaw_diff_rad"]) for c in rsc.detect()]))
        return out
