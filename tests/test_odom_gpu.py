"""End-to-end odometry parity: radarDriver::Process + OdometryKeyframeFuser::processFrame on the GPU (batched over
sequences, state on the device) vs the oracle run sequence by sequence on the same scans.

Bar (north_star): poses within 1e-5 m / 1e-6 rad; point counts, cell counts, association iterations and the keyframe
decisions identical.
"""
import numpy as np
import pytest

from tbv_slam_public_b200 import api, synth

pytestmark = pytest.mark.gpu

POS_TOL, ANG_TOL = 1e-5, 1e-6


def _ang(d):
    return np.abs(np.arctan2(np.sin(d), np.cos(d)))


def _run_pair(ctx, oracle, scans_per_seq, gkw, okw):
    n_seq, n_frames = len(scans_per_seq), len(scans_per_seq[0])
    n_az, n_range = scans_per_seq[0][0].shape
    fuser = api.OdometryKeyframeFuser(ctx, n_seq, n_az, n_range, api.default_odom_params(**gkw))
    refs = [oracle.Odometry(oracle.default_odom_params(**okw)) for _ in range(n_seq)]
    for f in range(n_frames):
        batch = np.stack([scans_per_seq[s][f] for s in range(n_seq)])
        outs = fuser.pointcloudCallback(batch)
        for s in range(n_seq):
            r = refs[s].step(scans_per_seq[s][f])
            g = outs[s]
            assert g.status == 0
            assert (g.n_points, g.n_cells, g.itrs, g.is_keyframe, g.n_keyframes, g.reg_ok) == \
                   (r.n_points, r.n_cells, r.itrs, r.is_keyframe, r.n_keyframes, r.reg_ok), f"frame {f} seq {s}"
            assert abs(g.pose[0] - r.pose[0]) < POS_TOL and abs(g.pose[1] - r.pose[1]) < POS_TOL, f"frame {f} seq {s}"
            assert _ang(g.pose[2] - r.pose[2]) < ANG_TOL, f"frame {f} seq {s}"
    # keyframe window parity at the end
    for s in range(n_seq):
        poses, nc = refs[s].keyframes()
        for i in range(len(poses)):
            cells, pose = fuser.cells(s, i)
            assert len(cells) == nc[i]
            assert np.abs(pose[:2] - poses[i][:2]).max() < POS_TOL and _ang(pose[2] - poses[i][2]) < ANG_TOL
            ref_cells = refs[s].keyframe_cells(i)
            assert np.allclose(cells[:, :2], ref_cells[:, :2], rtol=0, atol=1e-9)
    fuser.close()


def test_odometry_baseline_config(ctx, oracle):
    """BASELINE config 2: CFEAR-3 filter, 4 keyframes, P2L, Huber 0.1, combined weights — two sequences in lock-step."""
    a = synth.make_stream(14, s0=0.0)
    b = synth.make_stream(14, s0=300.0, seed=99)
    _run_pair(ctx, oracle, [list(a.scans), list(b.scans)], {}, {})


def test_odometry_cfear1_p2l_one_keyframe(ctx, oracle, stream8):
    g = dict(submap_scan_size=1, weight_intensity=0, res=3.5)
    gp = api.default_odom_params(**g)
    gp.filter.k_strongest = 12
    gp.filter.z_min = 70.0
    fuser_kw = dict(submap_scan_size=1, weight_intensity=0, res=3.5, filter=gp.filter)
    _run_pair(ctx, oracle, [list(stream8.scans)], fuser_kw, dict(submap_scan_size=1, weight_intensity=0, res=3.5, k_strongest=12, z_min=70.0))


def test_odometry_p2p_and_slow_motion_keyframes(ctx, oracle):
    """P2P cost (the CFEAR-3 preset) on a slow stream: most frames are NOT keyframes, the window is reused."""
    st = synth.make_stream(10, speed=2.0)
    rp = api.default_reg_params(cost=api.P2P, weight_opt=api.W_COMBINED, regularization=1.0)
    _run_pair(ctx, oracle, [list(st.scans)], dict(reg=rp), dict(cost_type=oracle.P2P))


def test_odometry_mulran_shape_rotate_on_receipt_ccw(ctx, oracle):
    """BASELINE config 5 shape: MulRan scans arrive range-major (3360 x 400) and are rotated 90 degrees CCW on receipt
    (radar_driver.cpp:80-84), 400 x 3360 bins of 0.0595238 m, counter-clockwise sweep (compensation sign flips), CFEAR-1 (MulRan) preset:
    z_min 70, k 12, r 3.5, P2L, one keyframe, uniform cell weights."""
    st = synth.make_stream(8, dataset=synth.MULRAN)
    scans = []
    for img in st.scans:
        wire = np.ascontiguousarray(np.rot90(img, -1))          # what the driver receives: range-major
        assert wire.shape == (3360, 400)
        rot = ctx.rotate90ccw(wire)
        assert np.array_equal(rot, img) and np.array_equal(oracle.rotate90ccw(wire), img)
        scans.append(rot)
    gp = api.default_odom_params(submap_scan_size=1, weight_intensity=0, res=3.5, radar_ccw=1)
    gp.filter.k_strongest = 12
    gp.filter.z_min = 70.0
    gp.filter.range_res = 0.0595238
    _run_pair(ctx, oracle, [scans], dict(submap_scan_size=1, weight_intensity=0, res=3.5, radar_ccw=1, filter=gp.filter),
              dict(submap_scan_size=1, weight_intensity=0, res=3.5, radar_ccw=1, k_strongest=12, z_min=70.0, range_res=0.0595238))
    # rotate-on-receipt inside the device step (tbv_odom_set_wire_layout): the wire-layout scans give, bit for bit, the poses of the
    # pre-rotated ones — two sequences so that the batched rotate kernel sees more than one image
    kw = dict(submap_scan_size=1, weight_intensity=0, res=3.5, radar_ccw=1, filter=gp.filter)
    pre = api.OdometryKeyframeFuser(ctx, 2, 400, 3360, api.default_odom_params(**kw))
    wired = api.OdometryKeyframeFuser(ctx, 2, 400, 3360, api.default_odom_params(**kw))
    wired.set_wire_layout(True)
    for f in range(len(scans) - 1):
        a = pre.pointcloudCallback(np.stack([scans[f], scans[f + 1]]))
        w = np.stack([np.ascontiguousarray(np.rot90(st.scans[f], -1)), np.ascontiguousarray(np.rot90(st.scans[f + 1], -1))])
        b = wired.pointcloudCallback(w.reshape(2, 400, 3360))      # same bytes, wire layout [3360][400] per scan
        assert np.array_equal(api.poses(a), api.poses(b)) and [o.n_points for o in a] == [o.n_points for o in b]
    pre.close(); wired.close()


def test_dense_short_range_clutter_uses_the_global_point_arrays(ctx, oracle):
    """16 000 points per scan (every azimuth keeps k = 40) inside 39 m: more points than the fused cells kernel holds in shared
    memory (8 192), so its point arrays live in global scratch — same results as the oracle, cell for cell."""
    rng = np.random.default_rng(5)
    img = np.full((400, 3768), 20, np.uint8)
    hit = rng.random((400, 840)) < 0.3
    img[:, 60:900][hit] = rng.integers(100, 256, size=int(hit.sum()), dtype=np.uint8)
    scans = [img, img, img]
    fuser = api.OdometryKeyframeFuser(ctx, 1, 400, 3768, api.default_odom_params())
    ref = oracle.Odometry(oracle.default_odom_params())
    for f in range(3):
        g = fuser.pointcloudCallback(scans[f][None])[0]
        r = ref.step(scans[f])
        assert g.status == 0 and g.n_points > 8192
        assert (g.n_points, g.n_cells, g.is_keyframe, g.n_keyframes) == (r.n_points, r.n_cells, r.is_keyframe, r.n_keyframes), f"frame {f}"
        assert abs(g.pose[0] - r.pose[0]) < POS_TOL and abs(g.pose[1] - r.pose[1]) < POS_TOL and _ang(g.pose[2] - r.pose[2]) < ANG_TOL
    cells, _ = fuser.cells(0, 0)
    ref_cells = ref.keyframe_cells(0)
    assert len(cells) == len(ref_cells) > 100
    assert np.allclose(cells[:, :2], ref_cells[:, :2], rtol=0, atol=1e-9)
    fuser.close()


def test_empty_and_nearly_empty_scans_do_not_break_the_pipeline(ctx):
    """A blank scan (no bin >= z_min) has no points, no cells and nothing to register: the reference exits on an empty cloud
    (pointnormal.cpp:72-75); the library reports zero counts, keeps the previous pose and carries on with the next frame."""
    st = synth.make_stream(3)
    blank = np.zeros((400, 3768), np.uint8)
    sparse = blank.copy()
    sparse[10, 500] = 200; sparse[200, 900] = 210          # two points: fewer than any cell needs
    fuser = api.OdometryKeyframeFuser(ctx, 2, 400, 3768, api.default_odom_params())
    seq = [np.stack([st.scans[0], st.scans[0]]), np.stack([blank, sparse]), np.stack([st.scans[1], st.scans[1]]), np.stack([st.scans[2], st.scans[2]])]
    outs = []
    for b in seq:   # the returned records live in a buffer the next call reuses: copy what is checked
        outs.append([(o.status, o.n_points, o.n_cells, tuple(o.pose)) for o in fuser.pointcloudCallback(b)])
    assert all(o[0] == 0 for f in outs for o in f)
    assert outs[1][0][1:3] == (0, 0) and outs[1][1][1:3] == (2, 0)
    for s in range(2):
        assert outs[2][s][2] > 100 and outs[3][s][2] > 100 and np.all(np.isfinite(outs[3][s][3]))
    fuser.close()


def test_pipelined_submit_collect_matches_sync(ctx):
    st = synth.make_stream(6)
    n_seq = 3
    p = api.default_odom_params()
    f1 = api.OdometryKeyframeFuser(ctx, n_seq, 400, 3768, p)
    sync = []
    for f in range(6):
        batch = np.stack([st.scans[(f + s) % 6] for s in range(n_seq)])
        sync.append(api.poses(f1.pointcloudCallback(batch)).copy())
    f1.close()
    f2 = api.OdometryKeyframeFuser(ctx, n_seq, 400, 3768, p)
    bufs = [api.PinnedBuffer(n_seq * 400 * 3768) for _ in range(2)]
    got = []
    for f in range(6):
        batch = np.stack([st.scans[(f + s) % 6] for s in range(n_seq)])
        bufs[f & 1].array[:] = batch.reshape(-1)
        f2.submit(bufs[f & 1].ptr)
        if f >= 1:
            got.append(api.poses(f2.collect()).copy())
    got.append(api.poses(f2.collect()).copy())
    f2.close()
    for a, b in zip(sync, got):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("wire", [False, True])
def test_overlapped_steps_change_nothing(ctx, wire):
    """tbv_odom_set_overlap: the filter of a step on a second stream (into a second set of clouds) under the previous step's registration,
    compensation as its own launch — poses, counts and keyframe decisions bit for bit those of the fused single-stream step, through the
    device-input call, the host-input pipeline and with the rotate-on-receipt wire layout; the step's clouds stay fetchable."""
    import torch
    st = synth.make_stream(8)
    n_seq, n_az, n_range = 3, 400, 3768
    p = api.default_odom_params()
    frames = [np.stack([st.scans[(f + s) % 8] for s in range(n_seq)]) for f in range(8)]
    if wire:
        frames = [np.ascontiguousarray(np.rot90(b, -1, axes=(1, 2))) for b in frames]     # [n_range][n_az] per scan, as the driver receives them
    dev = [torch.from_numpy(b.reshape(-1)).cuda() for b in frames]
    torch.cuda.synchronize()

    def run(overlap, host):
        fu = api.OdometryKeyframeFuser(ctx, n_seq, n_az, n_range, p)
        fu.set_wire_layout(wire)
        fu.set_overlap(overlap)
        rec, poses = [], []

        def take(out):   # the fuser hands out the same record array every time: copy what is compared
            rec.append([(o.n_points, o.n_cells, o.itrs, o.is_keyframe) for o in out])
            poses.append(api.poses(out).copy())

        if host:
            bufs = [api.PinnedBuffer(n_seq * n_az * n_range) for _ in range(2)]
            for f in range(8):
                bufs[f & 1].array[:] = frames[f].reshape(-1)
                fu.submit(bufs[f & 1].ptr)
                if f >= 1:
                    take(fu.collect())
                if f == 3:
                    ctx.StructuredKStrongest(st.scans[:2])        # another user of the context's clouds between two overlapped steps
            take(fu.collect())
        else:
            for f in range(8):
                fu.step_dev(dev[f].data_ptr())
                if f == 3:
                    ctx.StructuredKStrongest(st.scans[:2])        # (enqueued behind the step, before its results are fetched)
                take(fu.fetch())
        filt, _ = ctx.filter_fetch(n_seq, n_az, p.filter.k_strongest)      # the last step's (compensated) clouds
        fu.close()
        return rec, poses, filt

    ref_rec, ref_poses, ref_filt = run(False, False)
    for overlap, host in ((True, False), (True, True), (False, True)):
        rec, poses, filt = run(overlap, host)
        assert rec == ref_rec, (overlap, host)
        for a, b in zip(poses, ref_poses):
            assert np.array_equal(a, b), (overlap, host)
        for b in range(n_seq):
            for x, y in zip(filt.scan(b), ref_filt.scan(b)):
                assert np.array_equal(x, y), (overlap, host, b)


def test_cuda_graph_replay_of_the_step_changes_nothing(ctx, stream8):
    """The step is replayed from a CUDA graph once an input buffer has been seen twice (tbv_odom_set_graphs, default on): same poses, counts
    and keyframe decisions, bit for bit, as direct launches — also after another fuser on the same context has grown the context's scratch
    buffers (the captured graphs are then stale and must be re-captured, not replayed)."""
    a = api.OdometryKeyframeFuser(ctx, 1, 400, 3768)            # graphs on: host path alternates between its two upload buffers
    b = api.OdometryKeyframeFuser(ctx, 1, 400, 3768)
    b.set_graphs(False)
    l0 = ctx.launch_count()
    for f in range(6):
        oa = a.pointcloudCallback(stream8.scans[f][None])
        ob = b.pointcloudCallback(stream8.scans[f][None])
        assert np.array_equal(api.poses(oa), api.poses(ob))
        assert (oa[0].n_points, oa[0].n_cells, oa[0].itrs, oa[0].is_keyframe) == (ob[0].n_points, ob[0].n_cells, ob[0].itrs, ob[0].is_keyframe)
    per_step = (ctx.launch_count() - l0) / 12
    assert 6 <= per_step <= 8                                   # replayed steps are counted like launched ones
    big = api.OdometryKeyframeFuser(ctx, 3, 400, 3768)          # grows the context's filter / cells / registration scratch
    big.pointcloudCallback(stream8.scans[:3])
    for f in range(6, 8):
        oa = a.pointcloudCallback(stream8.scans[f][None])
        ob = b.pointcloudCallback(stream8.scans[f][None])
        assert np.array_equal(api.poses(oa), api.poses(ob)) and oa[0].n_cells == ob[0].n_cells
    a.close(); b.close(); big.close()
