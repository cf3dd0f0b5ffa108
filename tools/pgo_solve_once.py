"""One damped solve of a 4 500-node graph (radius 1e8: ~22 CG iterations) — the launch ncu captures for the pose-graph solver."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from tbv_slam_public_b200 import api  # noqa: E402
from pgo_bench import graph  # noqa: E402

ctx = api.Context(0)
nodes, ids, meas = graph(4500, np.random.default_rng(0))
_, Hd, Ho, g, _ = api.pgo_assemble(ctx, nodes, ids, meas)
for _ in range(3):
    d, it, rel = api.pgo_solve_step(ctx, ids, Hd, Ho, g, radius=1e8, rel_tol=1e-10)
print(it, rel)
