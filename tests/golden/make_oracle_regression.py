"""Generates tests/golden/oracle_regression.json: outputs of the CPU oracle (oracle/) on seeded synthetic inputs, downstream of the filter
stage.  These are NOT reference outputs (the reference's downstream code cannot be built here — DESIGN.md §2); they freeze the oracle
itself, so that an accidental change to the checker shows up as a diff instead of silently moving the parity target.
    python tests/golden/make_oracle_regression.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def compute():
    from oracle import oracle_py as o
    from tbv_slam_public_b200 import synth
    st = synth.make_stream(6)
    out = {}
    sets, peaks = [], []
    for i in range(4):
        r = o.kstrongest(st.scans[i], z_min=60.0, k=40)
        az, rg, I, x, y = r["filtered"]
        c, ns = o.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)
        sets.append(c)
        paz, prg, pI, px, py = r["peaks"]
        peaks.append((px, py, pI.astype(np.float32)))
        out[f"scan{i}"] = {"n_points": int(len(x)), "n_peaks": int(len(px)), "n_samples": int(ns), "n_cells": int(len(c)),
                           "cells_field_sums": [float(v) for v in c.sum(axis=0)]}
    T = np.array([(0, 0, 0), (2.5, 0, 0), (5.0, 0, 0), (7.4, 0.05, 0.003)], float)
    res = o.register(sets, T.copy(), o.default_reg_params())
    out["register_p2l_huber_combined"] = {"result": json.loads(json.dumps(res, default=lambda v: v.tolist() if hasattr(v, "tolist") else
                                                                           {f[0]: getattr(v, f[0]) for f in v._fields_}))}
    n, score, cost_v, resid = o.get_cost(sets, T, o.default_reg_params(), itr=2)
    out["get_cost"] = {"n": int(n), "score": float(score), "cost": float(cost_v), "residual_sum": float(np.sum(resid)), "residual_abs_sum": float(np.abs(resid).sum())}
    S = o.cost_samples(sets, T, itr=2)
    out["cost_samples_sum"] = float(S[:, 3].sum())
    d = o.sc_make(*peaks[0])
    out["sc_make"] = {"desc_sum": float(np.sum(d[0])), "ringkey_sum": float(np.sum(d[1])), "sectorkey_sum": float(np.sum(d[2]))}
    dist, shift = o.sc_distance(o.sc_make(*peaks[0])[0], o.sc_make(*peaks[1])[0])
    out["sc_distance_0_1"] = {"dist": float(dist), "shift": int(shift)}
    q = o.coral_quality(peaks[1], peaks[0], st.gt[1], st.gt[0])
    out["coral_1_0"] = {k: (float(v) if not isinstance(v, (bool, int)) else int(v)) for k, v in q.items()}
    od = o.Odometry(o.default_odom_params())
    poses = []
    for i in range(6):
        r = od.step(st.scans[i])
        poses.append([float(r.pose[0]), float(r.pose[1]), float(r.pose[2]), int(r.n_cells), int(r.itrs), int(r.is_keyframe)])
    out["odometry_6_frames"] = poses
    return out


if __name__ == "__main__":
    data = compute()
    path = os.path.join(HERE, "oracle_regression.json")
    json.dump(data, open(path, "w"), indent=1)
    print(path, os.path.getsize(path), "bytes")
