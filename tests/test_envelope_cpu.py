"""Sanity envelope against the one input-less anchor the reference holds for the stages downstream of the filter (VERDICT r1
item 6b): tbv_slam/model_parameters/combined.txt, 58 071 rows written by the reference's own ScanLearningInterface on real Oxford data —
per keyframe pair [aligned, CorAl joint, CorAl separate, overlap, CFEAR cost, residual count, mean cells per scan].  The ranges are committed as
tests/golden/ref_envelope.json (tests/golden/make_envelope.py).

What this can and cannot show.  The synthetic world is not Oxford, so this is NOT a pin of the restatement: it checks that the quantities the
reference itself logged for ALIGNED consecutive keyframes — cells per scan (cell::cell / ComputeNormals validity rules at r = 3 m), the number
of P2L residuals of CFEARQuality (radius 2.0, 30 degree normal gate) and its Huber(0.3) cost — come out in the reference's own observed ranges when
the oracle (and, in tests/test_envelope_gpu.py, the CUDA path) processes Oxford-shape scans.  A restatement that associated twice as many or half
as many cells, kept degenerate cells, or summed the cost differently would leave these bands.  CorAl's entropies and overlap depend on
the peak density of the scene and are reported, not asserted."""
import json
import os

import numpy as np
import pytest

from tbv_slam_public_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ENV = json.load(open(os.path.join(HERE, "golden", "ref_envelope.json")))


def keyframe_pairs(process):
    """(cells_ref, cells_src, T_ref, T_src) of consecutive keyframes (2.5 m apart) along three stretches of the synthetic drive, clouds
    compensated with the frame-to-frame motion as processFrame does."""
    out = []
    for s0 in (0.0, 411.0, 822.0):
        st = synth.make_stream(8, s0=s0)
        cells = []
        for i in range(8):
            mot = synth.se2_mul(synth.se2_inv(st.gt[i - 1]), st.gt[i]) if i > 0 else np.zeros(3)
            cells.append(process(st.scans[i], mot))
        out += [(cells[i - 1], cells[i], st.gt[i - 1], st.gt[i]) for i in range(1, 8)]
    return out


def check_envelope(rows):
    """rows: [n, 3] = CFEAR cost, residual count, mean cells per scan of aligned pairs."""
    A = ENV["aligned"]
    cost, nres, cells = rows[:, 0], rows[:, 1], rows[:, 2]
    assert np.all((cells >= A["mean_cells"]["min"]) & (cells <= A["mean_cells"]["max"])), (cells.min(), cells.max())
    assert np.all((nres >= A["n_residuals"]["min"]) & (nres <= A["n_residuals"]["max"])), (nres.min(), nres.max())
    assert A["n_residuals"]["p01"] <= np.median(nres) <= A["n_residuals"]["p99"]
    inside = (cost >= A["cfear_cost"]["min"]) & (cost <= A["cfear_cost"]["max"])
    assert inside.mean() >= 0.8 and A["cfear_cost"]["p01"] <= np.median(cost) <= A["cfear_cost"]["p99"], (cost.min(), np.median(cost), cost.max())
    # the cost of an aligned pair per residual is what separates aligned from misaligned pairs in the reference's own data
    ref_per_res = A["cfear_cost"]["median"] / A["n_residuals"]["median"]
    assert 0.25 * ref_per_res <= np.median(cost / nres) <= 4.0 * ref_per_res


def test_envelope_file_is_the_reference_table():
    assert ENV["rows"] == 58071 and ENV["aligned"]["rows"] == 4467 and ENV["columns"][4:] == ["cfear_cost", "n_residuals", "mean_cells"]
    src = "/root/reference/tbv_slam/model_parameters/combined.txt"
    if os.path.exists(src):     # in the build container: the committed ranges are the file's
        d = np.loadtxt(src, delimiter=",")
        al = d[d[:, 0] == 1]
        assert abs(ENV["aligned"]["cfear_cost"]["mean"] - al[:, 4].mean()) < 1e-9 and ENV["all"]["mean_cells"]["max"] == d[:, 6].max()


def test_oracle_quantities_lie_in_the_reference_envelope(oracle):
    O = oracle

    def process(scan, mot):
        az, rg, I, x, y = O.kstrongest(scan, peaks=False)["filtered"]
        x, y = O.compensate(x, y, mot, False)
        return O.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)[0]

    P = O.default_reg_params(loss_limit=0.3)          # CFEARQuality: P2L, Huber 0.3, uniform weights (AlignmentQuality.cpp:336-344)
    rows = []
    for c_ref, c_src, T_ref, T_src in keyframe_pairs(process):
        n, score, cost, res = O.get_cost([c_ref, c_src], [T_ref, T_src], P, itr=0)
        rows.append((cost, n, (len(c_ref) + len(c_src)) / 2))
    check_envelope(np.array(rows))
