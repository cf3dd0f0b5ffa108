"""CorAl quality over a batch of candidate pairs: kernel time (run under ncu for the exact figure) and the oracle's per-pair time."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tbv_slam_public_b200 import api, synth
from oracle import oracle_py as o


def main(n=256):
    st = synth.make_stream(8)
    cl = []
    for i in range(8):
        az, rg, I, x, y = o.kstrongest(st.scans[i], z_min=60.0, k=40)["peaks"]
        cl.append((x, y, I.astype(np.float32)))
    ctx = api.Context(0)
    src = [(i % 7) + 1 for i in range(n)]
    ref = [i % 7 for i in range(n)]
    Ts = np.array([st.gt[s] for s in src]); Tr = np.array([st.gt[r] for r in ref])
    api.CorAlRadarQuality(ctx, cl, src, ref, Ts, Tr)
    t = time.perf_counter()
    for _ in range(5):
        res = api.CorAlRadarQuality(ctx, cl, src, ref, Ts, Tr)
    wall = (time.perf_counter() - t) / 5
    t = time.perf_counter()
    m = min(n, 64)
    for k in range(m):
        o.coral_quality(cl[src[k]], cl[ref[k]], Ts[k], Tr[k])
    cpu = (time.perf_counter() - t) / m
    print(json.dumps({"pairs": n, "points_per_cloud": float(np.mean([len(c[0]) for c in cl])), "gpu_call_ms (upload + kernel + download)": wall * 1e3,
                      "gpu_pairs_per_s": n / wall, "oracle_ms_per_pair_1_thread": cpu * 1e3, "valid": int(sum(r.valid for r in res))}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
