"""The sharded loop-closure search at world size 2 over gloo, on CPU: every rank runs the same batched search on the
same graph; the registration stage is sharded by `id_from mod world` (parallel.shard_candidates), each rank registers its share — with the
oracle here, tbv_loopdb_register_dev on the GPUs — packs the accepted candidates into 128-byte tbv_constraint records and ONE all-gather
(parallel.all_gather_constraints) gives every rank the full list.  Both ranks must end with the records of the single-process search."""
import os
import pickle
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tbv_slam_public_b200 import parallel, tbv_slam as TS  # noqa: E402
from tbv_slam_public_b200.api import CONSTRAINT_DTYPE  # noqa: E402
import test_tbv_slam_cpu as T  # noqa: E402


class ShardedOracleLoopDevice(T.OracleLoopDevice):
    """GpuLoopDevice(sharded=True) with the oracle in place of the kernels: same sharding rule, same record layout, same collective."""

    def __init__(self, world, rank):
        super().__init__()
        self.world, self.rank, self.shares = world, rank, []

    def register(self, id_from, id_to, T_from, T_to):
        self.calls["register"] += 1
        id_from, id_to = np.asarray(id_from, np.int32), np.asarray(id_to, np.int32)
        mine = parallel.shard_candidates(id_from, self.world, self.rank)
        self.shares.append(len(mine))
        cap = max(parallel.shard_capacity(id_from, self.world), 1)
        buf = np.zeros(cap, CONSTRAINT_DTYPE)
        n = 0
        for p in mine:
            ok, Ta, Tr, itrs, score = self.O.loop_register(self.cells[id_from[p]], self.cells[id_to[p]], T_from[p], T_to[p])
            if ok:
                buf[n]["id_begin"], buf[n]["id_end"], buf[n]["type"], buf[n]["candidate"] = id_from[p], id_to[p], 1, p
                buf[n]["t_be"], buf[n]["cov"], buf[n]["score"], buf[n]["t_revised"], buf[n]["itrs"] = Ta, [0.01, 0.0, 0.01, 1e-4], score, Tr, itrs
                n += 1
        local = torch.from_numpy(buf.view(np.uint8).reshape(cap, parallel.RECORD_BYTES))
        out = parallel.all_gather_constraints(local, torch.tensor([n], dtype=torch.int32))
        acc = {int(c["candidate"]): c for c in out}
        return [(True, np.array(acc[p]["t_be"]), np.array(acc[p]["cov"]), float(acc[p]["score"])) if p in acc
                else (False, np.zeros(3), np.array([1.0, 0.0, 1.0, 1.0]), 0.0) for p in range(len(id_from))]


def _summary(loop):
    return [(r.id_from, r.id_to, r.guess_nr, r.reg_ok, r.applied, r.probability, r.t_be.tobytes(), sorted(r.quality.items())) for r in loop.statistics]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, graph_bytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = pickle.loads(graph_bytes)
        dev = ShardedOracleLoopDevice(world, rank)
        loop = TS.ScanContextClosure(g, dev, T._classifier(), TS.LoopClosureParams())
        loop.SearchAndAddConstraintBatched()
        q.put((rank, pickle.dumps((_summary(loop), sorted(loop.loop_constraints), dev.shares, dev.calls))))
    finally:
        dist.destroy_process_group()


def test_sharded_batched_search_world2_gloo():
    g, gt, est = T.drive.__wrapped__()
    single = TS.ScanContextClosure(T._copy(g), T.OracleLoopDevice(), T._classifier(), TS.LoopClosureParams())
    single.SearchAndAddConstraintBatched()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    blob = pickle.dumps(g)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blob, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = pickle.loads(res[0]), pickle.loads(res[1])
    assert r0[0] == r1[0] == _summary(single)                         # every rank: the single-process records, in order
    assert r0[1] == r1[1] == sorted(single.loop_constraints) and len(r0[1]) >= 6
    n_cand = sum(1 for r in single.statistics if r.guess_nr >= 0)
    assert r0[2][0] + r1[2][0] == n_cand and min(r0[2][0], r1[2][0]) > 0.3 * n_cand      # one registration call, split about evenly
    assert r0[3]["register"] == r1[3]["register"] == 1
