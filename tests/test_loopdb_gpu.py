"""Loop-closure keyframe database (tbv_loopdb_*): batched RegisterLoopCandidate vs the oracle's loopclosure::Register, the
constraint records, and the sharded front end (world size 1 here; world size 2 is covered on CPU with gloo in
test_parallel_cpu.py and on 2 GPUs with NCCL by tests/tools/loop_bench.py).

Bar: accept/reject decisions and iteration counts identical; Talign / Trevised within 1e-5 m / 1e-6 rad (north_star).
"""
import numpy as np
import pytest

from tbv_slam_public_b200 import api

pytestmark = pytest.mark.gpu

POS_TOL, ANG_TOL = 1e-5, 1e-6


def _ang(a, b):
    d = a - b
    return np.abs(np.arctan2(np.sin(d), np.cos(d)))


@pytest.fixture(scope="module")
def cellsets(oracle, stream8):
    sets = []
    for i in range(8):
        az, rg, I, x, y = oracle.kstrongest(stream8.scans[i])["filtered"]
        c, _ = oracle.build_cells(x, y, I.astype(np.float32), radius=3.0, weight_intensity=True)
        sets.append(c)
    return sets


def _candidates(gt, n, seed, n_kf):
    rng = np.random.default_rng(seed)
    fs, ts, Tf, Tt = [], [], [], []
    for _ in range(n):
        a, b = rng.choice(n_kf, 2, replace=False)
        err = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(-0.1, 0.1)])
        fs.append(a); ts.append(b); Tf.append(gt[a] + err); Tt.append(gt[b])
    return np.array(fs, np.int32), np.array(ts, np.int32), np.array(Tf), np.array(Tt)


def test_loopdb_matches_oracle_loop_register(ctx, oracle, stream8, cellsets):
    db = api.LoopDB(ctx, 16, 1024)
    assert db.add(cellsets[:5]) == 0 and db.add(cellsets[5:]) == 5 and len(db) == 8
    fs, ts, Tf, Tt = _candidates(stream8.gt, 40, 11, 8)
    # a hopeless candidate (no overlap at all): must be rejected like the reference rejects it
    fs = np.append(fs, 0).astype(np.int32); ts = np.append(ts, 1).astype(np.int32)
    Tf = np.vstack([Tf, [500.0, 500.0, 0.0]]); Tt = np.vstack([Tt, [0.0, 0.0, 0.0]])
    quality = np.stack([np.linspace(0, 1, len(fs)), np.linspace(1, 2, len(fs))], axis=1)
    out, summ = db.register_candidates(fs, ts, Tf, Tt, quality=quality, want_summaries=True)
    accepted = {int(c["candidate"]): c for c in out}
    assert list(out["candidate"]) == sorted(out["candidate"])          # candidate order
    n_ok = 0
    for p in range(len(fs)):
        ok, Ta_ref, Tr_ref, itrs, score = oracle.loop_register(cellsets[fs[p]], cellsets[ts[p]], Tf[p], Tt[p])
        assert bool(summ[p].success) == ok and (p in accepted) == ok
        if not ok:
            continue
        n_ok += 1
        c = accepted[p]
        assert (c["id_begin"], c["id_end"], c["type"]) == (fs[p], ts[p], 1)
        assert c["itrs"] == itrs == summ[p].itrs
        assert np.abs(Ta_ref[:2] - c["t_be"][:2]).max() < POS_TOL and _ang(Ta_ref[2], c["t_be"][2]) < ANG_TOL
        assert np.abs(Tr_ref[:2] - c["t_revised"][:2]).max() < POS_TOL and _ang(Tr_ref[2], c["t_revised"][2]) < ANG_TOL
        assert abs(score - c["score"]) <= 1e-9 * abs(score)
        # reg_cov: diag(0.1^2, 0.1^2, 0.01^2) with the xy block rotated into the revised frame (stays 0.01 * I)
        assert np.allclose(c["cov"], [0.01, 0.0, 0.01, 1e-4], atol=1e-15)
        assert np.array_equal(c["quality"], quality[p])
    assert n_ok >= 30 and len(fs) - 1 not in accepted
    db.close()


def test_loopdb_equals_register_batch_and_score_gate(ctx, stream8, cellsets):
    db = api.LoopDB(ctx, 8, 1024)
    db.add(cellsets)
    fs, ts, Tf, Tt = _candidates(stream8.gt, 64, 5, 8)
    Tr, Ta, summ = ctx.RegisterBatch(cellsets, fs, ts, Tf, Tt)
    out = db.register_candidates(fs, ts, Tf, Tt)
    ok = np.array([s.success for s in summ], bool)
    assert np.array_equal(out["candidate"], np.nonzero(ok)[0])
    assert np.array_equal(out["t_be"], Ta[ok])              # same kernel, same data: bit-identical
    scores = np.array([s.score for s in summ])[ok]
    gate = float(np.median(scores))
    gated = db.register_candidates(fs, ts, Tf, Tt, max_score=gate)
    assert np.array_equal(gated["candidate"], np.nonzero(ok)[0][scores <= gate])
    db.close()


def test_loopdb_errors_and_empty(ctx, cellsets):
    db = api.LoopDB(ctx, 2, 1024)
    db.add(cellsets[:2])
    with pytest.raises(api.TbvError):
        db.add(cellsets[2:3])                                # full
    with pytest.raises(api.TbvError):
        db.register_candidates([0], [5], [(0, 0, 0)], [(0, 0, 0)])   # unknown keyframe
    assert len(db.register_candidates([], [], np.zeros((0, 3)), np.zeros((0, 3)))) == 0
    db.close()
    small = api.LoopDB(ctx, 2, 16)
    with pytest.raises(api.TbvError):
        small.add(cellsets[:1])                              # set larger than the cell capacity
    small.close()


def test_sharded_front_end_world1(ctx, stream8, cellsets):
    from tbv_slam_public_b200 import parallel
    db = api.LoopDB(ctx, 8, 1024)
    db.add(cellsets)
    fs, ts, Tf, Tt = _candidates(stream8.gt, 48, 9, 8)
    ref = db.register_candidates(fs, ts, Tf, Tt)
    got = parallel.ShardedLoopClosure(db).register_candidates(fs, ts, Tf, Tt)
    assert got.tobytes() == ref.tobytes()
    # emulate two ranks on one GPU: the union of the two shares, merged by candidate, is the serial result
    parts = []
    for r in range(2):
        mine = parallel.shard_candidates(fs, 2, r)
        parts.append(db.register_candidates(fs[mine], ts[mine], Tf[mine], Tt[mine], candidate_index=mine.astype(np.int32)))
    merged = np.concatenate(parts)
    merged = merged[np.argsort(merged["candidate"], kind="stable")]
    assert merged.tobytes() == ref.tobytes()
    db.close()
