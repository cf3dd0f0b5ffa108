"""tbv_slam.py on the GPU (SURVEY §8d config C4 at test size): the same 1.4-lap drive as tests/test_tbv_slam_cpu.py through
tbv_slam.GpuLoopDevice — Scan-Context, batched candidate registration against the resident keyframe database, CorAl + CFEAR quality and
the pose-graph solve all through the C-ABI — checked against ground truth and against the oracle-backed run of the same driver.
Tolerances: the oracle comparison is on the SET of applied loops (>= 80 % common: a point on a sector edge may reorder two Scan-Context
candidates, tests/test_loop_gpu.py docstring), registered transforms within 0.3 m / 0.02 rad of ground truth, drift more than halved."""
import copy
import math

import numpy as np
import pytest

from tbv_slam_public_b200 import api, synth, tbv_slam as TS
from test_tbv_slam_cpu import N_KF, N_LAP, OracleLoopDevice, _classifier, drive  # noqa: F401  (drive is a fixture)

pytestmark = pytest.mark.gpu


def test_offline_slam_closes_the_loop_on_the_gpu(ctx, drive):
    g, gt, est = drive
    dev = TS.GpuLoopDevice(ctx, max_keyframes=64)
    launches0 = ctx.launch_count()
    slam = TS.TBVSLAM(copy.deepcopy(g), dev, _classifier(), TS.LoopClosureParams(), api.default_pgo_params(loop_scaling=1.0))
    res = slam.Run()
    assert ctx.launch_count() > launches0 + 4 * N_KF                   # descriptors, search, registration, quality, solver: all device work
    applied = [r for r in slam.loop.statistics if r.applied]
    assert len(applied) >= 6
    for r in applied:
        assert r.id_from >= N_LAP - 2 and abs((r.id_from - r.id_to) - N_LAP) <= 2 and r.probability > 0.9 and r.reg_ok
        Tgt = synth.se2_mul(synth.se2_inv(gt[r.id_from]), gt[r.id_to])
        assert np.hypot(*(r.t_be[:2] - Tgt[:2])) < 0.3 and abs(math.remainder(r.t_be[2] - Tgt[2], 2 * math.pi)) < 0.02
    # the oracle-backed run of the same driver finds (nearly) the same loops
    ref = TS.TBVSLAM(copy.deepcopy(g), OracleLoopDevice(), _classifier(), TS.LoopClosureParams(), api.default_pgo_params(loop_scaling=1.0))
    ref.ProcessFrame(False, True)
    a = {(r.id_from, r.id_to) for r in applied}
    b = {(r.id_from, r.id_to) for r in ref.loop.statistics if r.applied}
    assert len(a & b) >= 0.8 * max(len(a), len(b))
    assert {r.id_from for r in slam.loop.statistics} == set(range(N_KF))
    # optimisation with the verified loops pulls the dead-reckoned second lap back onto the first
    rel = lambda P: np.array([synth.se2_mul(synth.se2_inv(P[0]), p) for p in P])
    want, before, after = rel(gt), rel(res.poses_before), rel(res.poses_after)
    e_before = np.hypot(*(before[N_LAP:, :2] - want[N_LAP:, :2]).T).max()
    e_after = np.hypot(*(after[N_LAP:, :2] - want[N_LAP:, :2]).T).max()
    assert res.n_loop_constraints == len(slam.loop.loop_constraints) and res.summary.final_cost < res.summary.initial_cost
    assert e_before > 1.0 and e_after < 0.5 * e_before
    dev.close()


def test_batched_and_sharded_search_equal_the_per_keyframe_search_on_the_gpu(ctx, drive):
    """One registration / CorAl / CFEAR launch for the whole graph (SearchAndAddConstraintBatched), directly and through the sharded front end
    (parallel.ShardedLoopClosure, world size 1 here; 2 GPUs: tests/tools/loop_bench.py), against one launch per keyframe: every kernel works on
    one candidate per CTA, so the records must be the same whatever the batch they arrived in."""
    g, gt, est = drive
    runs = []
    for kw, batched in ((dict(), False), (dict(), True), (dict(sharded=True), True)):
        dev = TS.GpuLoopDevice(ctx, max_keyframes=64, **kw)
        loop = TS.ScanContextClosure(copy.deepcopy(g), dev, _classifier(), TS.LoopClosureParams())
        n0 = ctx.launch_count()
        assert (loop.SearchAndAddConstraintBatched() if batched else loop.SearchAndAddConstraint()) is False
        runs.append((loop, ctx.launch_count() - n0))
        dev.close()
    ref, n_ref = runs[0]
    assert len(ref.loop_constraints) >= 6
    for loop, n_launch in runs[1:]:
        assert n_launch < n_ref                                       # fewer, larger launches
        assert len(loop.statistics) == len(ref.statistics)
        for a, b in zip(ref.statistics, loop.statistics):
            assert (a.id_from, a.id_to, a.guess_nr, a.reg_ok, a.applied) == (b.id_from, b.id_to, b.guess_nr, b.reg_ok, b.applied)
            assert np.allclose(a.t_be, b.t_be, rtol=0, atol=1e-12) and abs(a.probability - b.probability) <= 1e-12
        assert loop.loop_constraints.keys() == ref.loop_constraints.keys()
