"""The .tbvg layout across languages: graph_io.py writes, the C++ mirror (include/tbv_b200.hpp: ParseSimpleGraph / SerializeSimpleGraph /
GraphToOptimizerInput / CeresLeastSquaresT) parses, re-serialises byte-identically, optimises (oracle backend, CPU) and writes back;
graph_io.py reads the result."""
import os
import subprocess

import numpy as np

from tbv_slam_public_b200 import graph_io as G
from test_graph_io_cpu import _drive

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_graph_io")


def test_cpp_reads_writes_and_optimises_a_python_graph(tmp_path):
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", BIN, os.path.join(ROOT, "tests", "cpp", "test_graph_io.cpp")])
    rng = np.random.default_rng(12)
    g = _drive(25, rng)
    g.AddGroundTruth([s.stamp_ for s, _ in g.graph][::3], [np.array([i, 0.5 * i, 0.01 * i]) for i in range(9)])
    # a loop constraint that disagrees with the dead-reckoned poses by a few centimetres (inside the Cauchy kernel's quadratic range):
    # the optimiser has something to do
    Tb, Te = g.graph[22][0].GetPose(), g.graph[2][0].GetPose()
    off = G.pose3d_to_matrix(G.pose3d_from_xyt((0.04, -0.03, 0.002)))
    g.AddConstraint(G.Constraint3d(22, 2, G.pose3d_from_matrix(np.linalg.inv(Tb) @ Te @ off), np.eye(6), G.LOOP_APPEARANCE,
                                   {"sc-sim": 0.07, "odom-bounds": 0.0, "alignment_quality": 4.5}, "verified"))
    g.AddConstraint(G.Constraint3d(9, 4, G.pose3d_from_xyt((0, 0, 0)), np.eye(6), G.CANDIDATE, {}, "Trusted candidate"))
    src, copy, opt = (str(tmp_path / n) for n in ("in.tbvg", "copy.tbvg", "opt.tbvg"))
    G.save_simple_graph(src, g)
    r = subprocess.run([BIN, src, copy, opt], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    line, rejected = r.stdout.strip().splitlines()
    f = line.split()
    v = dict(zip(f[0::2], f[1::2]))
    n_cells = sum(len(s.cloud_normal_) for s, _ in g.graph)
    n_pts = sum(len(s.cloud_peaks_) + len(s.cloud_nopeaks_) for s, _ in g.graph)
    assert (int(v["nodes"]), int(v["constraints"]), int(v["optimised"]), int(v["cells"]), int(v["points"]), int(v["gt"]), int(v["quality"])) == \
        (25, 26, 25, n_cells, n_pts, 9, 3)
    assert v["identical"] == "1" and open(copy, "rb").read() == open(src, "rb").read()
    assert rejected == "rejected 3"
    # the C++ optimiser saw the same problem the Python view describes, and improved it
    from oracle import oracle_py as O
    O.lib()
    nodes, ids, meas, info, _ = g.pgo_arrays()
    P = O.default_pgo_params(loop_scaling=1.0)
    c0 = O.pgo_assemble(nodes, ids, meas, P)[0]
    assert abs(float(f[f.index("cost") + 1]) - c0) <= 1e-12 * c0
    h = G.load_simple_graph(opt)
    nodes2 = h.pgo_arrays()[0]
    c1 = O.pgo_assemble(nodes2, ids, meas, P)[0]
    assert abs(float(f[f.index("cost") + 2]) - c1) <= 1e-9 * max(c1, 1e-12) and c1 < 0.5 * c0
    assert np.array_equal(nodes2[0], nodes[0]) and not np.allclose(nodes2[20], nodes[20])
    for (a, ca), (b, cb) in zip(g.graph, h.graph):                    # everything but the poses travels unchanged
        assert np.array_equal(a.cloud_normal_, b.cloud_normal_) and a.stamp_ == b.stamp_ and len(ca) == len(cb)
        for c, d in zip(ca, cb):
            assert c.quality == d.quality and c.info == d.info and np.array_equal(c.information, d.information)
