"""The CUDA path against the reference's own logged value ranges (tests/test_envelope_cpu.py explains what this shows): k-strongest +
compensation + cells + CFEARQuality through the C-ABI on Oxford-shape keyframe pairs."""
import numpy as np
import pytest

from tbv_slam_public_b200 import api
from test_envelope_cpu import check_envelope, keyframe_pairs

pytestmark = pytest.mark.gpu


def test_gpu_quantities_lie_in_the_reference_envelope(ctx):
    def process(scan, mot):
        f, _ = ctx.StructuredKStrongest(scan, peaks=False)
        az, rg, I, x, y = f.scan(0)
        x, y = ctx.Compensate(x.copy(), y.copy(), mot, False)
        return ctx.MapPointNormal(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)[0]

    pairs = keyframe_pairs(process)
    sets, src, ref, Ts, Tr = [], [], [], [], []
    for c_ref, c_src, T_ref, T_src in pairs:
        ref.append(len(sets)); sets.append(c_ref)
        src.append(len(sets)); sets.append(c_src)
        Ts.append(T_src); Tr.append(T_ref)
    q = ctx.CFEARQualityBatch(sets, src, ref, np.array(Ts), np.array(Tr))        # [n, 3] = cost, residual count, mean cells
    check_envelope(q)
