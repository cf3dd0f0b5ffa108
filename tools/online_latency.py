"""Online latency of ONE sequence (the reference's online node: one scan in, one pose out): wall time per frame of tbv_odom_step — upload,
7 kernels, pose record back — with the step replayed from a CUDA graph and with direct launches."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api, synth  # noqa: E402


def run(ctx, scans, graphs: bool):
    fuser = api.OdometryKeyframeFuser(ctx, 1, 400, 3768)
    fuser.set_graphs(graphs)
    pin = api.PinnedBuffer(scans[0].nbytes)
    ts = []
    for f in range(len(scans)):
        pin.array[:] = scans[f].reshape(-1)
        t0 = time.perf_counter()
        fuser.pointcloudCallback(pin.array.reshape(1, 400, 3768))
        ts.append(time.perf_counter() - t0)
    poses = api.poses(fuser._out).copy()
    fuser.close(); pin.free()
    return np.array(ts[8:]) * 1e3, poses


def main():
    ctx = api.Context(0)
    st = synth.make_stream(72)
    out = {}
    for graphs in (True, False, True, False):
        ms, poses = run(ctx, st.scans, graphs)
        key = "graph_replay" if graphs else "direct_launches"
        out.setdefault(key, []).append({"median_ms": round(float(np.median(ms)), 4), "p10_ms": round(float(np.percentile(ms, 10)), 4), "p90_ms": round(float(np.percentile(ms, 90)), 4)})
        out["final_pose_" + key] = poses.tolist()
    out["same_result"] = out["final_pose_graph_replay"] == out["final_pose_direct_launches"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
