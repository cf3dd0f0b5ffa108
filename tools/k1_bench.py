"""Micro-benchmark of K1/K2 alone (device-resident scans): ms per launch and achieved HBM GB/s. Run on the GPU box."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api, synth


def main(batch=256, iters=20, kind="radar"):
    """kind: 'radar' (synthetic Oxford-shape scans) or a synth.stress_image kind: 'uniform' (A), 'equal' (B), 'ramp', 'sparse', 'zeros'."""
    ctx = api.Context(0)
    if kind == "radar":
        src = synth.make_stream(16).scans
    else:
        src = np.stack([synth.stress_image(kind, seed=s + 1) for s in range(16)])
    idx = np.arange(batch) % 16
    dev = torch.from_numpy(src[idx]).cuda()
    par = api.FilterParams(60.0, 40, 2.5, 0.0438)
    s = torch.cuda.ExternalStream(ctx.stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.filter_dev(dev.data_ptr(), 400, 3768, batch, par)
    ctx.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        ctx.filter_dev(dev.data_ptr(), 400, 3768, batch, par)
        e1.record(s)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    bytes_alg = batch * (400 * 3768 + 400 * 40 * 13)
    print(json.dumps({"kind": kind, "batch": batch, "ms_k1_k2": ms, "scans_per_s": batch / ms * 1e3, "algorithmic_GBps": bytes_alg / ms / 1e6,
                      "min_ms": float(np.min(ts))}))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]), *sys.argv[3:4])
