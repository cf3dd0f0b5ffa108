"""Experiment: G independent fusers (contexts/streams) of n_seq/G sequences each, stepped alternately — do the latency-bound
kernels of different groups overlap?  Prints scans/s for G = 1, 2, 4."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api, synth

def run(n_seq, G, n_frames=14, timed=6, stagger=False):
    st = synth.make_stream(n_frames + 16)
    first = (np.arange(n_seq) * 5) % 16
    per = n_seq // G
    ctxs = [api.Context(0) for _ in range(G)]
    fus = [api.OdometryKeyframeFuser(ctxs[g], per, 400, 3768, api.default_odom_params()) for g in range(G)]
    dev = [[torch.from_numpy(st.scans[first[g * per:(g + 1) * per] + t]).cuda() for g in range(G)] for t in range(n_frames)]
    torch.cuda.synchronize()
    for t in range(n_frames - timed):
        for g in range(G):
            fus[g].step_dev(dev[t][g].data_ptr())
    for c in ctxs: c.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(n_frames - timed, n_frames):
        for g in range(G):
            fus[g].step_dev(dev[t][g].data_ptr())
    for c in ctxs: c.synchronize()
    dt = time.perf_counter() - t0
    outs = [f.fetch() for f in fus]
    pose_sum = float(sum(o.pose[0] for oo in outs for o in oo))
    for f in fus: f.close()
    for c in ctxs: c.close()
    return {"n_seq": n_seq, "groups": G, "ms_per_step": dt / timed * 1e3, "scans_per_s": n_seq * timed / dt, "pose_checksum": pose_sum}

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    for G in (1, 2, 4):
        print(json.dumps(run(n, G)))
