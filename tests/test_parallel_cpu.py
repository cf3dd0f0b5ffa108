"""Host-side multi-GPU logic on CPU: sharding rules and the constraint all-gather with world size 2 over gloo.
The records are synthetic here (no GPU): what is checked is the partition, the padding / count protocol and the global order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tbv_slam_public_b200 import parallel  # noqa: E402
from tbv_slam_public_b200.api import CONSTRAINT_DTYPE  # noqa: E402


def test_shard_sequences_is_a_balanced_partition():
    for n, w in [(512, 8), (513, 8), (7, 8), (0, 4), (1000, 3)]:
        parts = [parallel.shard_sequences(n, w, r) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1


def test_shard_candidates_partition_and_capacity():
    rng = np.random.default_rng(0)
    id_from = rng.integers(0, 4500, 1000)
    for w in (1, 2, 4, 8):
        shares = [parallel.shard_candidates(id_from, w, r) for r in range(w)]
        assert sorted(np.concatenate(shares).tolist()) == list(range(1000))
        for r, s in enumerate(shares):
            assert np.all(id_from[s] % w == r) and np.all(np.diff(s) > 0)
        assert parallel.shard_capacity(id_from, w) == max(len(s) for s in shares)
    assert parallel.shard_capacity([], 4) == 0


def _records(idx):
    """Deterministic synthetic constraint for global candidate i."""
    r = np.zeros(len(idx), CONSTRAINT_DTYPE)
    r["candidate"] = idx
    r["id_begin"] = idx * 3 + 1
    r["id_end"] = idx // 2
    r["type"] = 1
    r["t_be"] = np.stack([idx * 0.5, -idx * 0.25, np.sin(idx)], axis=1)
    r["score"] = 1.0 / (1 + idx)
    return r


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cand, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        id_from = (np.arange(n_cand) * 7 + 3) % 11
        mine = parallel.shard_candidates(id_from, world, rank)
        accepted = mine[(mine % 3) != 0]                     # "registration failed" for every third candidate
        cap = max(parallel.shard_capacity(id_from, world), 1)
        buf = np.zeros(cap, CONSTRAINT_DTYPE)
        buf[:len(accepted)] = _records(accepted)
        local = torch.from_numpy(buf.view(np.uint8).reshape(cap, parallel.RECORD_BYTES))
        count = torch.tensor([len(accepted)], dtype=torch.int32)
        out = parallel.all_gather_constraints(local, count)
        q.put((rank, out.tobytes()))
    finally:
        dist.destroy_process_group()


def _run(world, n_cand):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cand, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_all_gather_constraints_world2_gloo():
    n_cand = 257
    res = _run(2, n_cand)
    all_idx = np.arange(n_cand)
    expect = _records(all_idx[(all_idx % 3) != 0]).tobytes()
    assert res[0] == expect and res[1] == expect            # every rank: the serial list, in global candidate order


def test_all_gather_constraints_empty_share_gloo():
    # 2 candidates: rank 1 holds candidate 0 and rejects it (an empty share), rank 0 accepts candidate 1; then 0 candidates at all
    res = _run(2, 2)
    assert res[0] == res[1]
    assert np.frombuffer(res[0], CONSTRAINT_DTYPE)["candidate"].tolist() == [1]
    res = _run(2, 0)
    assert res[0] == res[1] == b""
