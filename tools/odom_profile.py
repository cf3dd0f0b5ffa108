"""Per-kernel device time of the odometry pipeline (tbv_profile_begin/_end) for a given number of lock-step sequences."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api, synth


def main(n_seq=512, n_frames=10, prof_frames=3):
    ctx = api.Context(0)
    st = synth.make_stream(n_frames + 16)
    first = (np.arange(n_seq) * 5) % 16
    fuser = api.OdometryKeyframeFuser(ctx, n_seq, 400, 3768, api.default_odom_params())
    dev = [torch.from_numpy(st.scans[first + t]).cuda() for t in range(n_frames)]
    torch.cuda.synchronize()
    s = torch.cuda.ExternalStream(ctx.stream)
    for t in range(n_frames - prof_frames):
        fuser.step_dev(dev[t].data_ptr())
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # timed: last prof_frames without profiling
    fuser2 = api.OdometryKeyframeFuser(ctx, n_seq, 400, 3768, api.default_odom_params())
    for t in range(n_frames - prof_frames):
        fuser2.step_dev(dev[t].data_ptr())
    ctx.synchronize()
    e0.record(s)
    for t in range(n_frames - prof_frames, n_frames):
        fuser2.step_dev(dev[t].data_ptr())
    e1.record(s)
    e1.synchronize()
    ms_step = e0.elapsed_time(e1) / prof_frames
    ctx.profile_begin()
    for t in range(n_frames - prof_frames, n_frames):
        fuser.step_dev(dev[t].data_ptr())
    recs = ctx.profile_end()
    agg = {}
    for name, ms in recs:
        agg[name] = agg.get(name, 0.0) + ms / prof_frames
    outs = fuser.fetch()
    itrs = np.mean([o.itrs for o in outs]); lm = np.mean([o.lm_iterations for o in outs]); nres = np.mean([o.num_residuals for o in outs])
    ncell = np.mean([o.n_cells for o in outs])
    print(json.dumps({"n_seq": n_seq, "ms_per_step": ms_step, "scans_per_s": n_seq / ms_step * 1e3, "kernels_ms": {k: round(v, 4) for k, v in agg.items()},
                      "sum_ms": sum(agg.values()), "itrs": itrs, "lm": lm, "nres": nres, "ncell": ncell,
                      "status": int(max(o.status for o in outs))}))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
