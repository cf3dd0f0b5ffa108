"""Extracts the one input-less anchor the reference holds for the downstream stages: the value ranges of
tbv_slam/model_parameters/combined.txt — 58 071 rows [aligned, coral_joint, coral_sep, overlap, cfear_cost, n_residuals, mean_cells] written by
the reference's own ScanLearningInterface on real Oxford data — into tests/golden/ref_envelope.json (min / 1 % / median / 99 % / max per
column, over all rows and over the aligned rows).  Run in the build container (reads /root/reference); the JSON travels."""
import json
import os
import sys

import numpy as np

SRC = "/root/reference/tbv_slam/model_parameters/combined.txt"
COLS = ["aligned", "coral_joint", "coral_sep", "overlap", "cfear_cost", "n_residuals", "mean_cells"]


def main():
    d = np.loadtxt(SRC, delimiter=",")
    out = {"source": "tbv_slam/model_parameters/combined.txt", "rows": int(len(d)), "columns": COLS}
    for name, rows in (("all", d), ("aligned", d[d[:, 0] == 1])):
        out[name] = {"rows": int(len(rows))}
        for c, col in enumerate(COLS):
            v = rows[:, c]
            out[name][col] = {"min": float(v.min()), "p01": float(np.percentile(v, 1)), "median": float(np.median(v)), "p99": float(np.percentile(v, 99)),
                              "max": float(v.max()), "mean": float(v.mean())}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_envelope.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path)


if __name__ == "__main__":
    sys.exit(main())
