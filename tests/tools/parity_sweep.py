"""Frame-by-frame parity of the batched GPU odometry against the oracle over the bench's 8 stretches of the synthetic world:
counts (points, cells, association rounds, keyframe decisions) must be identical, poses within 1e-5 m / 1e-6 rad.
Prints the worst deviations and every count mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from tbv_slam_public_b200 import api  # noqa: E402
from oracle import oracle_py as o  # noqa: E402


def main(n_seq=64, n_frames=25):
    pool = bench.make_pool(n_frames + bench.POOL_EXTRA, 0)
    first = bench.first_offsets(n_seq, n_frames + bench.POOL_EXTRA, 0)
    ctx = api.Context(0)
    fuser = api.OdometryKeyframeFuser(ctx, n_seq, bench.N_AZ, bench.N_RANGE, api.default_odom_params())
    refs = [o.Odometry(o.default_odom_params()) for _ in range(n_seq)]
    worst_xy = worst_yaw = 0.0
    mismatches = []
    for f in range(n_frames):
        batch = pool[first + f]
        outs = fuser.pointcloudCallback(batch)
        for s in range(n_seq):
            r = refs[s].step(batch[s])
            g = outs[s]
            a = (g.n_points, g.n_cells, g.itrs, g.is_keyframe, g.n_keyframes, g.reg_ok)
            b = (r.n_points, r.n_cells, r.itrs, r.is_keyframe, r.n_keyframes, r.reg_ok)
            if a != b:
                mismatches.append({"frame": f, "seq": s, "first": int(first[s]), "gpu": a, "oracle": b})
            d = np.array(g.pose[:]) - np.array(r.pose[:])
            worst_xy = max(worst_xy, float(np.abs(d[:2]).max()))
            worst_yaw = max(worst_yaw, float(abs(np.arctan2(np.sin(d[2]), np.cos(d[2])))))
    print(json.dumps({"sequences": n_seq, "frames": n_frames, "max_abs_xy_m": worst_xy, "max_abs_yaw_rad": worst_yaw, "count_mismatches": len(mismatches),
                      "first_mismatches": mismatches[:8]}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
