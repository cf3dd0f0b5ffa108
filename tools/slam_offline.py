"""The reference's offline flow on a B200, end to end (precompute odometry -> simple graph -> tbv_slam_offline -> evaluation):

    python tools/slam_offline.py --scans scans.npy [--gt gt.npy] [--alignment-coefficients trained_alignment_classifier.txt] --out OUT

scans.npy: [n, 400, 3768] u8 polar scans (Oxford layout); gt.npy: [n, 3] (x, y, yaw), optional.  Without --scans a synthetic stream is
rendered (tbv_slam_public_b200/synth.py).  Without coefficients the alignment classifier is trained from the odometry itself, as
odometry_training_node does.  Writes OUT/est/00.txt, OUT/gt/00.txt, OUT/simple_graph.tbvg, OUT/loop/loop.csv, OUT/est_slam/00.txt and
prints the drift / ATE before and after loop closure.  GPU only: there is no CPU fallback."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tbv_slam_public_b200 import api, graph_io as G, offline_odometry as OO, synth, tbv_slam as TS, trajectory_io as TIO, verification as V  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans"); ap.add_argument("--gt"); ap.add_argument("--alignment-coefficients"); ap.add_argument("--out", required=True)
    ap.add_argument("--frames", type=int, default=400, help="synthetic stream length when --scans is not given")
    ap.add_argument("--loop-scaling", type=float, default=500000.0, help="ceresoptimizer.cpp:18-27 default")
    args = ap.parse_args()
    if args.scans:
        scans = np.load(args.scans, mmap_mode="r")
        gt = np.load(args.gt) if args.gt else None
    else:
        st = synth.make_stream(args.frames)
        scans, gt = st.scans, st.gt
    ctx = api.Context(0)
    odo = OO.GpuOdometryDevice(ctx, scans.shape[1], scans.shape[2])
    rd = OO.radarReader(odo).run(scans, gt=gt)
    paths = rd.Save(args.out)
    g = rd.graph
    dev = TS.GpuLoopDevice(ctx, max_keyframes=len(g) + 1)
    if args.alignment_coefficients:
        clf = V.LogisticRegression().LoadCoefficients(args.alignment_coefficients)
    else:
        sli = TS.ScanLearningInterface(dev)
        for s, _ in g.graph:
            sli.AddTrainingData(G.pose3d_to_xyt(s.T), s.cloud_peaks_, s.cloud_normal_)
        sli.FitModels()
        clf = sli.combined_class
    slam = TS.TBVSLAM(g, dev, clf, TS.LoopClosureParams(), api.default_pgo_params(loop_scaling=args.loop_scaling))
    res = slam.Run(batched=True)          # one registration / CorAl / CFEAR launch for the whole sequence
    os.makedirs(os.path.join(args.out, "loop"), exist_ok=True)
    os.makedirs(os.path.join(args.out, "est_slam"), exist_ok=True)
    n_rows = TS.write_loop_csv(os.path.join(args.out, "loop", "loop.csv"), g, slam.loop.statistics, "dataset,sequence", "synthetic,00")
    TIO.write_kitti(os.path.join(args.out, "est_slam", "00.txt"), res.poses_after)
    G.save_simple_graph(os.path.join(args.out, "optimised_graph.tbvg"), g)
    out = {"frames": len(rd.est), "keyframes": len(g), "candidates_evaluated": n_rows, "loops_applied": res.n_loop_constraints,
           "pgo": repr(res.summary), "kernel_launches": ctx.launch_count(), "files": paths}
    if gt is not None:
        kf_gt = {i: TIO.pose_matrix(gt[r]) for i, r in enumerate(rd.keyframe_rows)}
        for name, poses in (("odometry", res.poses_before), ("slam", res.poses_after)):
            e = TIO.evaluate(kf_gt, {i: TIO.pose_matrix(p) for i, p in enumerate(poses)}, "6dof", step_size=1)
            out[name] = {k: e[k] for k in ("t_err_percent", "r_err_deg_per_100m", "ate", "rpe_trans")}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
