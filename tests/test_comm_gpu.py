"""The multi-GPU exports of the C-ABI (tbv_comm_*, tbv_allgather_constraints, tbv_loopdb_register_sharded) and the
"no global state" contract of include/tbv_b200.h: several contexts in one process, on one device from two threads and on two devices.

Bar: the sharded call returns, on every rank, byte for byte the records of the single-GPU tbv_loopdb_register (same kernel, same data,
global candidate order)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from tbv_slam_public_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def cellsets(oracle, stream8):
    sets = []
    for i in range(8):
        az, rg, I, x, y = oracle.kstrongest(stream8.scans[i])["filtered"]
        c, _ = oracle.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)
        sets.append(c)
    return sets


def _cands(gt, n, seed):
    rng = np.random.default_rng(seed)
    fr = rng.integers(0, 8, n).astype(np.int32)
    to = ((fr + rng.integers(1, 4, n)) % 8).astype(np.int32)
    err = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-0.1, 0.1, n)], axis=1)
    return fr, to, gt[fr] + err, gt[to]


def test_register_sharded_without_and_with_a_one_rank_communicator(stream8, cellsets):
    import torch
    ctx = api.Context(0)
    db = api.LoopDB(ctx, 8, 1024)
    db.add(cellsets)
    fr, to, Tf, Tt = _cands(stream8.gt, 70, 21)
    ref = db.register_candidates(fr, to, Tf, Tt)
    assert ctx.comm_world() == (1, 0)
    assert db.register_sharded(fr, to, Tf, Tt).tobytes() == ref.tobytes()            # no communicator: world 1, NCCL never loaded
    ctx.comm_init_rank(api.Context.comm_unique_id(), 1, 0)                             # a real one-rank NCCL communicator
    assert ctx.comm_world() == (1, 0)
    got, timing = db.register_sharded(fr, to, Tf, Tt, want_timing=True)
    assert got.tobytes() == ref.tobytes() and len(timing) == 4 and timing[3] >= timing[0] > 0
    assert len(db.register_sharded([], [], np.zeros((0, 3)), np.zeros((0, 3)))) == 0
    # pipelined form (tbv_loopdb_submit_sharded / _collect_sharded): two batches in flight, collected in order, a third is refused
    ref5 = db.register_candidates(fr[:5], to[:5], Tf[:5], Tt[:5])
    for _ in range(3):
        db.submit_sharded(fr, to, Tf, Tt)
        db.submit_sharded(fr[:5], to[:5], Tf[:5], Tt[:5])
        with pytest.raises(api.TbvError):
            db.submit_sharded(fr, to, Tf, Tt)
        assert db.collect_sharded().tobytes() == ref.tobytes()
        db.submit_sharded(fr, to, Tf, Tt)
        assert db.collect_sharded().tobytes() == ref5.tobytes()
        got, timing = db.collect_sharded(want_timing=True)
        assert got.tobytes() == ref.tobytes() and timing[3] >= timing[0] > 0
    with pytest.raises(api.TbvError):
        db.collect_sharded()                                                   # nothing in flight
    # low-level export on caller-owned device buffers
    buf = torch.zeros((len(fr), 128), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    db.register_candidates_dev(fr, to, Tf, Tt, buf.data_ptr(), len(fr), cnt.data_ptr())
    assert ctx.allgather_constraints(buf.data_ptr(), cnt.data_ptr(), len(fr)).tobytes() == ref.tobytes()
    ctx.comm_destroy()
    assert db.register_sharded(fr, to, Tf, Tt).tobytes() == ref.tobytes()
    db.close(); ctx.close()


def test_two_contexts_on_one_device_from_two_threads(stream8):
    """Thread compatibility: distinct contexts driven from distinct host threads at the same time give the single-threaded results."""
    ref_ctx = api.Context(0)
    ref_f, ref_p = ref_ctx.StructuredKStrongest(stream8.scans[:4])
    ref = [tuple(a.copy() for a in ref_f.scan(b)) for b in range(4)]
    ref_ctx.close()
    errs = []

    def work(idx):
        try:
            c = api.Context(0)
            for _ in range(6):
                f, p = c.StructuredKStrongest(stream8.scans[:4])
                for b in range(4):
                    for a, r in zip(f.scan(b), ref[b]):
                        assert np.array_equal(a, r)
                az, rg, I, x, y = f.scan(idx)
                cells, _ = c.MapPointNormal(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)
                assert len(cells) > 50
            c.close()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_contexts_on_two_devices_in_one_process(stream8, oracle):
    """A second context on another device gets its own kernel attributes (the large dynamic shared-memory opt-ins are per device)."""
    c0, c1 = api.Context(0), api.Context(1)
    outs = []
    for c in (c0, c1, c0, c1):                                   # alternate: every entry point must select its context's device
        f, p = c.StructuredKStrongest(stream8.scans[:2])
        az, rg, I, x, y = f.scan(1)
        cells, _ = c.MapPointNormal(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)
        outs.append((x.copy(), y.copy(), np.array(cells)))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o, outs[0]))
    fu0 = api.OdometryKeyframeFuser(c0, 2, 400, 3768)
    fu1 = api.OdometryKeyframeFuser(c1, 2, 400, 3768)
    for t in range(3):
        a = fu0.pointcloudCallback(stream8.scans[t:t + 2])
        b = fu1.pointcloudCallback(stream8.scans[t:t + 2])
        assert np.array_equal(api.poses(a), api.poses(b))
    fu0.close(); fu1.close(); c0.close(); c1.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_register_sharded_world2_nccl(tmp_path):
    """Two ranks, two GPUs, NCCL: every rank receives exactly the single-GPU records; the low-level export agrees."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "workers", "sharded_loop_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    d = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    ref = d[0]["single"].tobytes()
    assert len(ref) > 40 * 128
    for k in range(2):
        assert d[k]["single"].tobytes() == ref
        for key in ("got", "again", "low", "pipe0", "pipe2"):
            assert d[k][key].tobytes() == ref, (k, key)
        assert d[k]["pipe1"].tobytes() == d[k]["small_ref"].tobytes()
        assert d[k]["small"].tobytes() == d[k]["small_ref"].tobytes() and d[k]["lop"].tobytes() == d[k]["lop_ref"].tobytes()
        assert d[k]["timing"][3] > 0
