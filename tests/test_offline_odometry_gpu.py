"""offline_odometry.radarReader over offline_odometry.GpuOdometryDevice on the B200 against the oracle-backed reader of
tests/test_offline_odometry_cpu.py on the same scans: same keyframes, poses within 1e-5 m / 1e-6 rad, node cells within 2e-6 (count exact; the kernel-level bar of
1e-9 is tests/test_odom_gpu.py's — this test is about the reader's plumbing), stored clouds within a few float ulps of the oracle's (device libm vs glibc in the compensation's atan2 / sincos)."""
import numpy as np
import pytest

from tbv_slam_public_b200 import graph_io as G, offline_odometry as OO, synth
from test_offline_odometry_cpu import OracleOdometryDevice

pytestmark = pytest.mark.gpu


def test_reader_on_the_gpu_matches_the_oracle_backed_reader(ctx, tmp_path):
    st = synth.make_stream(14, speed=4.0)
    dev = OO.GpuOdometryDevice(ctx, st.scans.shape[1], st.scans.shape[2])
    launches0 = ctx.launch_count()
    gpu = OO.radarReader(dev).run(st.scans, gt=st.gt)
    assert ctx.launch_count() >= launches0 + 4 * len(st.scans)
    ref = OO.radarReader(OracleOdometryDevice()).run(st.scans, gt=st.gt)
    assert gpu.keyframe_rows == ref.keyframe_rows and len(gpu.graph) == len(ref.graph) >= 5
    for a, b in zip(gpu.est, ref.est):
        assert np.abs(a[:2] - b[:2]).max() < 1e-5 and abs(np.remainder(a[2] - b[2] + np.pi, 2 * np.pi) - np.pi) < 1e-6
    for (sa, ca), (sb, cb) in zip(gpu.graph.graph, ref.graph.graph):
        assert sa.cloud_normal_.shape == sb.cloud_normal_.shape and np.allclose(sa.cloud_normal_[:, :2], sb.cloud_normal_[:, :2], rtol=0, atol=2e-6)
        assert sa.cloud_nopeaks_.shape == sb.cloud_nopeaks_.shape and sa.cloud_peaks_.shape == sb.cloud_peaks_.shape
        assert np.abs(sa.cloud_nopeaks_ - sb.cloud_nopeaks_).max() <= 5e-5 and np.array_equal(sa.cloud_nopeaks_[:, 3], sb.cloud_nopeaks_[:, 3])
        assert np.abs(sa.cloud_peaks_ - sb.cloud_peaks_).max() <= 5e-5
        assert len(ca) == len(cb) and all(np.allclose(x.t_be, y.t_be, atol=1e-5) for x, y in zip(ca, cb))
    paths = gpu.Save(str(tmp_path))
    assert len(G.load_simple_graph(paths["graph"])) == len(gpu.graph)
    dev.close()


def test_cloud_interface_fuser_on_the_gpu_equals_the_fused_frame(ctx):
    """OdometryKeyframeFuser.pointcloudCallback(cloud) over tbv_compensate / tbv_build_cells / tbv_register against tbv_odom_step on the same
    scans (both on the GPU): the same kernels see the same points, so poses agree to 1e-9 (host vs device pose algebra), keyframe decisions
    and iteration counts exactly; then the same fuser on CA-CFAR clouds (tbv_filter_cacfar) follows the ground truth."""
    from tbv_slam_public_b200 import api
    st = synth.make_stream(12, speed=4.0)
    fused = api.OdometryKeyframeFuser(ctx, 1, st.scans.shape[1], st.scans.shape[2])
    fuser = OO.OdometryKeyframeFuser(OO.GpuPrimitiveDevice(ctx), api.default_odom_params())
    for i in range(len(st.scans)):
        o = fused.pointcloudCallback(st.scans[i][None])[0]
        f, _ = ctx.StructuredKStrongest(st.scans[i], 60.0, 40, 2.5, 0.0438, peaks=False)
        _, _, I, x, y = f.scan(0)
        pose, _, _ = fuser.pointcloudCallback(x, y, I.astype(np.float32))
        assert np.abs(pose[:2] - np.array(o.pose[:2])).max() < 1e-9 and abs(pose[2] - o.pose[2]) < 1e-9, i
        assert (fuser.updated, fuser.last_itrs, len(fuser.keyframes_), len(fuser.last_cells)) == (bool(o.is_keyframe), o.itrs, o.n_keyframes, o.n_cells), i
    fused.close()
    st = synth.make_stream(10)
    par = api.default_odom_params(weight_intensity=0)
    par.reg = api.default_reg_params(cost=api.P2P, weight_opt=api.W_UNIFORM, regularization=1.0)
    fuser = OO.OdometryKeyframeFuser(OO.GpuPrimitiveDevice(ctx), par)
    est = []
    for i in range(len(st.scans)):
        out = ctx.AzimuthCACFAR(st.scans[i], window_size=40, false_alarm_rate=0.01, nb_guard_cells=10, capacity=65536)
        _, _, I, x, y = out.scan(0)
        est.append(fuser.pointcloudCallback(x, y, I.astype(np.float32))[0])
    est = np.array(est)
    rel_gt = np.array([synth.se2_mul(synth.se2_inv(st.gt[0]), p) for p in st.gt])
    assert np.hypot(*(est[:, :2] - rel_gt[:, :2]).T).max() < 0.5 and np.abs(est[:, 2] - rel_gt[:, 2]).max() < 0.02
